"""CPU oracle for the MQuant quantized-linear hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's algorithm (StiphyJay/MQuant, ``fake_quant/``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product (``mquant_b200``)
never does, and fails loudly when its CUDA library is missing.

Parity pinning: the reference has NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8(c)), so the oracle is pinned against outputs of the reference's
own Python executed in the build container: ``oracle/gen_golden.py`` imports
``/root/reference/fake_quant`` (two stubs, see ``oracle/_ref_import.py``), runs the
functions cited below on seeded inputs and freezes inputs+outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them through this file
bit-for-bit.

Third-party arithmetic absent from /root/reference: ``fast_hadamard_transform``
(Dao-AILab, unpinned; call sites hadamard_utils.py:118,124).  Its published algorithm
(Sylvester-ordered FWHT along the last dim, fp32 internally, output in input dtype,
times ``scale``) is restated in :func:`fht`; the in-tree pure-torch twin
``matmul_hadU`` (hadamard_utils.py:79-100) anchors it.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np

F16 = np.float16
F32 = np.float32

_HERE = os.path.dirname(os.path.abspath(__file__))
_HADK_PATH = os.path.join(os.path.dirname(_HERE), "tests", "golden", "hadk.bin")

# ----------------------------------------------------------------------------------
# bit types (fake_quant/bit_type.py:42-50).  ActQuantizer.configure looks up
# BIT_TYPE_DICT[f"int{bits}"] (quant_utils.py:165) so only the signed ones are reachable.
# ----------------------------------------------------------------------------------


def int_bounds(bits: int, signed: bool = True):
    """fake_quant/bit_type.py:18-28."""
    if signed:
        return -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
    return 0, 2 ** bits - 1


# ----------------------------------------------------------------------------------
# static activation quantizer: UniformQuantizer (fake_quant/quantizer/uniform.py:20-43)
# driven by BaseQuantizer.forward (fake_quant/quantizer/base.py:44-50)
# ----------------------------------------------------------------------------------


def act_quant_codes(x, scale, zero_point=0, bits=8):
    """``outputs = inputs / scale + zero_point; round; clamp`` (uniform.py:28-31).

    ``inputs`` is ``x.float()`` (base.py:46); ``scale`` keeps the observer's dtype (fp16 for
    fp16 activations) and is promoted to fp32 by the division -> true IEEE fp32 division,
    round-half-even (torch.round), clamp to the signed range.  Returns fp32 codes exactly
    as the reference holds them (they are integers in [-2^(b-1), 2^(b-1)-1]).
    ``scale`` may be a scalar (layer_wise) or ``[C]`` (channel_wise, reshaped to
    (1,1,-1) by get_reshape_range, base.py:14-33).
    """
    lo, hi = int_bounds(bits)
    xf = np.asarray(x).astype(F32)
    s = np.asarray(scale).astype(F32)
    zp = np.asarray(zero_point).astype(F32)
    out = xf / s + zp
    return np.clip(np.rint(out), lo, hi).astype(F32)


def act_dequant(codes, scale, zero_point=0, out_dtype=F16):
    """``(q - zero_point) * scale`` in fp32 then ``.to(dtype)`` (uniform.py:42, base.py:49)."""
    s = np.asarray(scale).astype(F32)
    zp = np.asarray(zero_point).astype(F32)
    return ((np.asarray(codes, dtype=F32) - zp) * s).astype(out_dtype)


def act_fake_quant(x, scale, zero_point=0, bits=8):
    """BaseQuantizer.forward (base.py:44-50): float -> quant -> dequantize -> input dtype."""
    x = np.asarray(x)
    return act_dequant(act_quant_codes(x, scale, zero_point, bits), scale, zero_point, x.dtype)


def msq_token_scales(modality, s_text, s_visual):
    """MSQ extension (no reference counterpart): per-token scale select.
    ``modality[m] != 0`` -> visual token.  Degenerates to the reference when the two scales
    are equal."""
    m = np.asarray(modality).astype(bool)
    return np.where(m, F32(s_visual), F32(s_text)).astype(F32)


def msq_quant_codes(x, modality, s_text, s_visual, bits=8):
    """a-3 arithmetic applied per token with the selected scale.  x: [M, C]."""
    s = msq_token_scales(modality, s_text, s_visual)[:, None]
    return act_quant_codes(x, s, 0, bits)


# ----------------------------------------------------------------------------------
# MinmaxObserver (fake_quant/observer/minmax.py:7-52, base.py:13-28)
# ----------------------------------------------------------------------------------


class MinmaxObserver:
    """Running per-channel max/min with 0 folded in on the first update (minmax.py:13-24),
    collapsed to a scalar in layer_wise mode (minmax.py:26-28).  Values keep the input dtype
    (fp16 activations -> fp16 statistics -> fp16 scale)."""

    def __init__(self, bits=8, calibration_mode="layer_wise"):
        self.bits = bits
        self.calibration_mode = calibration_mode
        self.max_val = None
        self.min_val = None
        self.eps = np.finfo(np.float32).eps  # base.py:13

    def update(self, v):
        v = np.asarray(v)
        v = v.reshape(-1, v.shape[-1]).T  # base.py:25-27  -> [C, tokens]
        cur_max = v.max(axis=-1)
        cur_min = v.min(axis=-1)
        if self.max_val is None:
            self.max_val = np.maximum(cur_max, np.zeros_like(cur_max))
        else:
            self.max_val = np.maximum(cur_max, self.max_val)
        if self.min_val is None:
            self.min_val = np.minimum(cur_min, np.zeros_like(cur_min))
        else:
            self.min_val = np.minimum(cur_min, self.min_val)
        if self.calibration_mode == "layer_wise":
            self.max_val = self.max_val.max()
            self.min_val = self.min_val.min()

    def get_quantization_params(self):
        return minmax_params(self.min_val, self.max_val, self.bits)


def minmax_params(min_val, max_val, bits=8):
    """Symmetric branch of MinmaxObserver.get_quantization_params (minmax.py:41-46).

    ``scale = max(|min/qmin|, |max/qmax|)`` evaluated IN THE DTYPE OF THE STATISTICS
    (fp16 for fp16 activations: the quotient is formed in fp32 and rounded to fp16, which is
    what torch's Half division kernel does), then ``clamp_(eps)`` with eps = 2^-23 (exactly
    representable as an fp16 subnormal).  zero_point = int64 zeros.
    """
    qmin, qmax = int_bounds(bits)
    min_val = np.asarray(min_val)
    max_val = np.asarray(max_val)
    dt = max_val.dtype
    a = np.abs((min_val.astype(F32) / F32(qmin)).astype(dt))
    b = np.abs((max_val.astype(F32) / F32(qmax)).astype(dt))
    scale = np.maximum(a, b)
    scale = np.maximum(scale, np.asarray(np.finfo(np.float32).eps, dtype=dt))
    return scale, np.zeros(scale.shape, dtype=np.int64)


# ----------------------------------------------------------------------------------
# bf16 activations (InternVL2 / Qwen2-VL).  numpy has no bfloat16: values travel as uint16 bit
# patterns; bf16 -> fp32 is exact (the pattern is the upper half of the fp32 word), fp32 -> bf16
# rounds to nearest even on the dropped 16 bits (what torch's .to(bfloat16) does).
# ----------------------------------------------------------------------------------
def bf16_to_f32(bits):
    return (np.asarray(bits, dtype=np.uint16).astype(np.uint32) << 16).view(F32)


def f32_to_bf16_bits(x):
    u = np.ascontiguousarray(np.asarray(x, dtype=F32)).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16  # round to nearest even (finite inputs)
    return r.astype(np.uint16)


def minmax_params_bf16(min_bits, max_bits, bits=8):
    """minmax_params for bf16 statistics: every torch op on bf16 tensors computes in fp32 and
    rounds the result to bf16 (minmax.py:41-46 evaluated in the dtype of the statistics)."""
    qmin, qmax = int_bounds(bits)
    rb = lambda v: bf16_to_f32(f32_to_bf16_bits(v))  # noqa: E731
    a = np.abs(rb(bf16_to_f32(min_bits) / F32(qmin)))
    b = np.abs(rb(bf16_to_f32(max_bits) / F32(qmax)))
    scale = np.maximum(np.maximum(a, b), rb(np.asarray(np.finfo(np.float32).eps, dtype=F32)))
    return f32_to_bf16_bits(scale)


def act_fake_quant_bf16(x_bits, scale_bits, bits=8):
    """BaseQuantizer.forward on bf16 activations: codes from the fp32 quotient, dequantised in
    fp32, rounded into bf16 (base.py:44-50, uniform.py:28-43).  Returns (codes fp32, y bits)."""
    codes = act_quant_codes(bf16_to_f32(x_bits), bf16_to_f32(scale_bits), 0, bits)
    return codes, f32_to_bf16_bits(codes * bf16_to_f32(scale_bits))


# ----------------------------------------------------------------------------------
# PercentileObserver (fake_quant/observer/percentile.py:26-79)
# ----------------------------------------------------------------------------------


def torch_quantile_f32(sorted_f32, q):
    """torch.quantile(v.float(), q) restated (ATen Sorting.cpp quantile_compute, linear
    interpolation): q is materialised in the INPUT dtype (fp32), rank = q*(n-1) in fp32,
    lerp(lo, hi, w) = w<0.5 ? lo + w*(hi-lo) : hi - (hi-lo)*(1-w) in fp32."""
    n = sorted_f32.shape[0]
    qf = F32(q)
    rank = F32(qf * F32(n - 1))
    below = int(np.floor(rank))
    above = int(np.ceil(rank))
    w = F32(rank - F32(below))
    lo = F32(sorted_f32[below])
    hi = F32(sorted_f32[above])
    diff = F32(hi - lo)
    if w < F32(0.5):
        return F32(lo + F32(w * diff))
    return F32(hi - F32(diff * F32(F32(1) - w)))


class PercentileObserver:
    """percentile.py:9-79.  layer_wise only (:28); first call clamps with 0 (:45,51), later
    calls are an EMA with sigma (:47-55) -> order dependent.  Statistics are fp32.
    Only the torch.quantile branch (:31-32) is restated; the numpy fallback the reference
    takes for > 2^24 elements (:33-43) is out of the oracle's pinned range."""

    def __init__(self, bits=8, sigma=0.01, alpha=0.99999):
        self.bits = bits
        self.sigma = sigma
        self.alpha = alpha
        self.max_val = None
        self.min_val = None

    def update(self, v):
        flat = np.sort(np.asarray(v).reshape(-1).astype(F32), kind="stable")
        cur_max = torch_quantile_f32(flat, self.alpha)
        cur_min = torch_quantile_f32(flat, 1.0 - self.alpha)
        if self.max_val is None:
            self.max_val = np.maximum(cur_max, F32(0))
        else:
            self.max_val = F32(self.max_val + F32(F32(self.sigma) * F32(cur_max - self.max_val)))
        if self.min_val is None:
            self.min_val = np.minimum(cur_min, F32(0))
        else:
            self.min_val = F32(self.min_val + F32(F32(self.sigma) * F32(cur_min - self.min_val)))

    def get_quantization_params(self):
        return minmax_params(F32(self.min_val), F32(self.max_val), self.bits)


# ----------------------------------------------------------------------------------
# WeightQuantizer (fake_quant/quant_utils.py:415-524) + sym_quant (:42-53)
# ----------------------------------------------------------------------------------


def weight_find_params(W, bits=4, mse=False, norm=2.4, grid=100, maxshrink=0.8):
    """Per-output-channel symmetric scale (quant_utils.py:446-466, perchannel=True, sym=True).
    All statistics are fp32 even for fp16 W (``tmp`` is fp32, :458).  Optional clip search
    (:473-500).  Returns scale [N,1] fp32."""
    x = np.asarray(W)
    x = x.reshape(x.shape[0], -1)
    maxq = F32(2 ** (bits - 1) - 1)
    xmin = np.minimum(x.min(1).astype(F32), F32(0))
    xmax = np.maximum(x.max(1).astype(F32), F32(0))
    xmax = np.maximum(np.maximum(np.abs(xmin), xmax), F32(1e-5))
    scale = (xmax / maxq).astype(F32)
    if mse:
        xf = x.astype(F32)
        best = np.full(x.shape[0], np.inf, dtype=F32)
        for i in range(int(maxshrink * grid)):
            p = 1 - i / grid
            xmax1 = (F32(p) * xmax).astype(F32)
            scale1 = (xmax1 / maxq).astype(F32)
            q = scale1[:, None] * np.clip(np.rint(xf / scale1[:, None]), -(maxq + 1), maxq)
            q = np.abs(q.astype(F32) - xf) ** F32(norm)
            err = q.sum(1, dtype=F32)
            better = err < best
            best[better] = err[better]
            scale[better] = scale1[better]
    return scale.reshape(-1, 1)


def weight_quant_codes(W, scale, bits=4):
    """sym_quant (quant_utils.py:42-45): q = clamp(round(x / scale), -(maxq+1), maxq);
    fp16 W is promoted to fp32 by the fp32 scale."""
    maxq = 2 ** (bits - 1) - 1
    x = np.asarray(W).astype(F32)
    return np.clip(np.rint(x / np.asarray(scale, dtype=F32)), -(maxq + 1), maxq).astype(np.int8)


def weight_fake_quant(W, scale, bits=4):
    """WeightQuantizer.quantize (quant_utils.py:512-518): scale * q, cast back to W.dtype."""
    W = np.asarray(W)
    q = weight_quant_codes(W, scale, bits).astype(F32)
    return (np.asarray(scale, dtype=F32) * q).astype(W.dtype)


# ----------------------------------------------------------------------------------
# int4 packing (fake_quant/quant_utils.py:56-94)
# ----------------------------------------------------------------------------------


def pack_i4(q):
    """pack_i4 (:61-68): even index -> low nibble, odd index -> high nibble, two's complement,
    along the last dim."""
    q = np.asarray(q)
    assert q.min() >= -8 and q.max() <= 7
    u = np.where(q < 0, 16 + q, q).astype(np.uint8)
    return (u[..., 0::2] | (u[..., 1::2] << 4)).astype(np.uint8)


def unpack_i4(p):
    """unpack_i4 (:72-94)."""
    p = np.asarray(p, dtype=np.uint8)
    lo = (p & 0x0F).astype(np.int8)
    lo[lo >= 8] -= 16
    hi = ((p & 0xF0) >> 4).astype(np.int8)
    hi[hi >= 8] -= 16
    out = np.empty(p.shape[:-1] + (p.shape[-1] * 2,), dtype=np.int32)
    out[..., 0::2] = lo
    out[..., 1::2] = hi
    return out


# ----------------------------------------------------------------------------------
# Hadamard (fake_quant/hadamard_utils.py:9-128)
# ----------------------------------------------------------------------------------

_HADK_ORDER = [172, 156, 140, 108, 60, 52, 36, 28, 40, 20, 12]  # get_hadK priority (:28-75)
_DEFAULT_SIZE = [172, 156, 140, 108, 60, 52, 36, 28, 20, 12, 1]  # auto_pad_size (:6)
_hadk_cache = None


def _load_hadk():
    global _hadk_cache
    if _hadk_cache is None:
        with open(_HADK_PATH, "rb") as f:
            blob = f.read()
        assert blob[:8] == b"MQHADK1\0"
        (count,) = struct.unpack_from("<I", blob, 8)
        off = 12
        tab = {}
        for _ in range(count):
            (k,) = struct.unpack_from("<I", blob, off)
            off += 4
            nb = (k * k + 7) // 8
            bits = np.unpackbits(np.frombuffer(blob, np.uint8, nb, off), bitorder="little")[: k * k]
            off += nb
            tab[k] = (bits.astype(F32) * 2 - 1).reshape(k, k)
        _hadk_cache = tab
    return _hadk_cache


def is_pow2(n):
    """hadamard_utils.py:194-195."""
    return (n & (n - 1) == 0) and (n > 0)


def get_hadK(n, transpose=False):
    """hadamard_utils.py:26-76: first factor in priority order that divides n; the cofactor
    must be a power of two (assert).  Returns (hadK fp32 [K,K] or None, K)."""
    for k in _HADK_ORDER:
        if n % k == 0:
            assert is_pow2(n // k)
            h = _load_hadk()[k]
            return (h.T.copy() if transpose else h.copy()), k
    assert is_pow2(n)
    return None, 1


def auto_pad_size(n):
    """hadamard_utils.py:9-23."""
    for size in _DEFAULT_SIZE:
        if n % size == 0 and is_pow2(n // size):
            return n
    new_n = math.inf
    for size in _DEFAULT_SIZE:
        hadk = math.ceil(n / size)
        hadk = 2 ** math.ceil(math.log2(hadk))
        if hadk * size < new_n:
            new_n = hadk * size
    return new_n


def fwht_last(x):
    """Unnormalised Sylvester-ordered Walsh-Hadamard transform along the last dim, in the
    dtype of x (butterflies of hadamard_utils.py:84-91 restated; identical to x @ H_n^T)."""
    x = np.array(x, copy=True)
    n = x.shape[-1]
    assert is_pow2(n)
    h = 1
    shp = x.shape
    x = x.reshape(-1, n)
    while h < n:
        y = x.reshape(x.shape[0], n // (2 * h), 2, h)
        a = y[:, :, 0, :] + y[:, :, 1, :]
        b = y[:, :, 0, :] - y[:, :, 1, :]
        x = np.stack([a, b], axis=2).reshape(-1, n)
        h *= 2
    return x.reshape(shp)


def fht(x, scale):
    """fast_hadamard_transform.hadamard_transform(x, scale) (third-party, see module doc):
    fp32 butterflies, times fp32 scale, result rounded to x.dtype."""
    x = np.asarray(x)
    return (fwht_last(x.astype(F32)) * F32(scale)).astype(x.dtype)


def matmul_hadU(X, transpose=False, dtype=np.float64):
    """Pure-torch twin (hadamard_utils.py:79-100) in ``dtype``: Y = X (hadK (x) H_{n/K})^T / sqrt(n)."""
    X = np.asarray(X).astype(dtype)
    n = X.shape[-1]
    hadK, K = get_hadK(n, transpose)
    t = X.reshape(-1, K, n // K)
    t = fwht_last(t)
    if K > 1:
        t = np.matmul(hadK.astype(dtype)[None], t)
    # ``torch.tensor(n).sqrt()`` is an fp32 value (hadamard_utils.py:100)
    return (t.reshape(X.shape) / dtype(np.sqrt(F32(n)))).astype(dtype)


def matmul_hadU_cuda(X, hadK=None, K=None, fp32_had=False):
    """Online transform as the wrapper runs it (hadamard_utils.py:115-128 called from
    quant_utils.py:334-341), with the reference's rounding points:

      K == 1 : FHT(x, 1/sqrt(n))                       -> x.dtype           (:117-120)
      K  > 1 : view [-1,K,n/K]; FHT(last, 1/sqrt(n))   -> x.dtype           (:123-126)
               hadK.to(dtype) @ .  (fp32 accumulate)    -> x.dtype           (:127)
      fp32_had: the whole chain runs on x.float() and is rounded once at the end
                (quant_utils.py:336-339).
    The scale is float32(1/sqrt(n)) (torch.tensor(n).sqrt() is fp32).
    """
    X = np.asarray(X)
    n = X.shape[-1]
    if K is None:
        hadK, K = get_hadK(n)
    scale = F32(1.0) / np.sqrt(F32(n))
    work = X.astype(F32) if fp32_had else X
    if K == 1:
        out = fht(work, scale)
    else:
        t = fht(work.reshape(-1, K, n // K), scale)
        # fp16 operands, fp32 accumulation (batched BLAS; same values as the einsum it replaces)
        out = np.matmul(hadK.astype(F32)[None], t.astype(F32)).astype(work.dtype)
        out = out.reshape(X.shape)
    return out.astype(X.dtype)


def pad_last(x, new_size):
    """utils.revise_down_input (fake_quant/utils.py:465-471): right zero-pad the last dim."""
    x = np.asarray(x)
    pad = new_size - x.shape[-1]
    if pad <= 0:
        return x
    return np.concatenate([x, np.zeros(x.shape[:-1] + (pad,), x.dtype)], axis=-1)


# ----------------------------------------------------------------------------------
# The wrapper (fake_quant/quant_utils.py:330-391)
# ----------------------------------------------------------------------------------


def linear_f16(x, W, bias):
    """fp16 nn.Linear as cuBLAS/MKL run it: fp16 operands, fp32 accumulation, fp16 result."""
    y = np.asarray(x).astype(F32) @ np.asarray(W).astype(F32).T
    if bias is not None:
        y = y + np.asarray(bias).astype(F32)
    return y.astype(F16)


def qlinear_fake(x, Wfq, bias, act_scale, online_had=False, fp32_had=False, split=False,
                 pad_to=None):
    """ActQuantWrapper.forward (quant_utils.py:330-391), static int8 activations,
    weights already fake-quantised (fp16 values), out_quantizer 16 bit.
      1. optional zero-pad (pre-hook) and online full Hadamard             (:334-341)
      2. split : x[...,1:] = Q(x[...,1:]); y = L1.float()(x0) + L2.float()(x[1:])  fp32  (:367-376)
         else  : x = Q(x); y = module(x)  fp16                               (:377-384)
    """
    x = np.asarray(x)
    if pad_to is not None:
        x = pad_last(x, pad_to)
    if online_had:
        x = matmul_hadU_cuda(x, fp32_had=fp32_had)
    Wfq = np.asarray(Wfq)
    if split:
        xq = x.copy()
        xq[..., 1:] = act_fake_quant(x[..., 1:], act_scale)
        y = xq[..., 0:1].astype(F32) @ Wfq[:, 0:1].astype(F32).T
        y2 = xq[..., 1:].astype(F32) @ Wfq[:, 1:].astype(F32).T
        if bias is not None:
            y2 = y2 + np.asarray(bias).astype(F32)
        return (y + y2).astype(x.dtype)
    xq = act_fake_quant(x, act_scale)
    return linear_f16(xq, Wfq, bias)


def qlinear_int(codes_a, s_a, codes_w, s_w, bias=None, x0=None, w0=None):
    """Real-integer restatement of the same layer: exact int32 accumulation of int8 x int4
    codes, then the epilogue the CUDA kernel specifies (include/mquant_b200.h):

        y[m,n] = fp16( fp32(acc[m,n]) * (s_a[m] * s_w[n])  (+ x0[m]*w0[n])  (+ bias[n]) )

    every operation a single fp32 rounding in that order (fmaf is NOT used).  s_a is a scalar
    or [M]; s_w [N].  This is what the W4A8 GEMM must reproduce bit-for-bit; its distance
    from :func:`qlinear_fake` is the fp16-rounding tolerance of SURVEY.md 8(c)."""
    acc = np.asarray(codes_a, dtype=np.int32) @ np.asarray(codes_w, dtype=np.int32).T
    s_a = np.broadcast_to(np.asarray(s_a, dtype=F32).reshape(-1, 1), (acc.shape[0], 1))
    s = (s_a * np.asarray(s_w, dtype=F32).reshape(1, -1)).astype(F32)
    y = (acc.astype(F32) * s).astype(F32)
    if x0 is not None:
        r1 = (np.asarray(x0).astype(F32).reshape(-1, 1) * np.asarray(w0).astype(F32).reshape(1, -1)).astype(F32)
        y = (y + r1).astype(F32)
    if bias is not None:
        y = (y + np.asarray(bias).astype(F32).reshape(1, -1)).astype(F32)
    return y.astype(F16)


# ----------------------------------------------------------------------------------
# dynamic (non-static) per-token quantizer (fake_quant/quant_utils.py:205-268, sym branch)
# ----------------------------------------------------------------------------------


def silu_mul(a1, a2):
    """a1 * F.silu(a2) on fp16 tensors (QWenMLP.forward, model/modeling_qwen_opt.py:385): silu in fp32
    (x / (1 + exp(-x))), rounded to fp16 (the tensor F.silu returns), product rounded to fp16."""
    a1, a2 = np.asarray(a1, F16), np.asarray(a2, F16)
    x = a2.astype(F32)
    with np.errstate(over="ignore"):
        s = (x / (F32(1) + np.exp(-x, dtype=F32))).astype(F32).astype(F16)
    return (a1.astype(F32) * s.astype(F32)).astype(F16)


def gptq_hessian(H, nsamples, x, batch):
    """GPTQ.add_batch (fake_quant/gptq/gptq_utils.py:182-195) in float64 (the check value for the
    fp32 tensor-core sum): returns (H', nsamples')."""
    x = np.asarray(x, F16).astype(np.float64).reshape(-1, x.shape[-1])
    n1 = nsamples + batch
    return np.asarray(H, np.float64) * (nsamples / n1) + (2.0 / n1) * (x.T @ x), n1


def gptq_block(W1, Hinv1, scale, maxq):
    """The column loop of GPTQ.fasterquant for one block (gptq_utils.py:254-283), fp32 operation by
    fp32 operation: q = scale * clamp(round(w / scale), -(maxq+1), maxq) (quant_utils.py:42-57),
    Losses = (w - q)^2 / d^2, err = (w - q) / d, W1[:, i:] -= err (x) Hinv1[i, i:] (rounded product,
    rounded difference).  Returns (W1', Q1, Err1, Losses1)."""
    W1 = np.array(W1, F32)
    Hinv1 = np.asarray(Hinv1, F32)
    s = np.asarray(scale, F32).reshape(-1)
    count = W1.shape[1]
    Q1, Err1, L1 = np.zeros_like(W1), np.zeros_like(W1), np.zeros_like(W1)
    for i in range(count):
        w = W1[:, i].copy()
        d = Hinv1[i, i]
        q = (s * np.clip(np.rint((w / s).astype(F32)), -(maxq + 1), maxq).astype(F32)).astype(F32)
        diff = (w - q).astype(F32)
        Q1[:, i] = q
        L1[:, i] = ((diff * diff).astype(F32) / (d * d).astype(F32)).astype(F32)
        err = (diff / d).astype(F32)
        Err1[:, i] = err
        W1[:, i:] = (W1[:, i:] - (err[:, None] * Hinv1[i, i:][None, :]).astype(F32)).astype(F32)
    return W1, Q1, Err1, L1


def rmsn(x, mean_dim, eps=1e-5):
    """RMSN.forward (fake_quant/module_util.py:55-61) for fp16 input: fp32 arithmetic, fp16 result.
    (np.sum's pairwise fp32 summation is ONE valid order of torch's reduction; tests compare within
    the stated ulp budget.)"""
    xf = np.asarray(x, F16).astype(F32)
    variance = (xf * xf).sum(-1, keepdims=True, dtype=F32) / F32(mean_dim)
    r = F32(1.0) / np.sqrt(variance + F32(eps), dtype=F32)
    return (xf * r).astype(F16)


def dynamic_per_token_scales(x, bits=8, clip_ratio=1.0):
    """find_params, per-token symmetric (:236-250): min/max with 0 in fp32 (``tmp`` fp32),
    times clip_ratio, scale = xmax / maxq, rows that are all zero get scale 1."""
    x = np.asarray(x)
    x2 = x.reshape(-1, x.shape[-1])
    maxq = F32(2 ** (bits - 1) - 1)
    xmin = (np.minimum(x2.min(1).astype(F32), F32(0)) * F32(clip_ratio)).astype(F32)
    xmax = (np.maximum(x2.max(1).astype(F32), F32(0)) * F32(clip_ratio)).astype(F32)
    xmax = np.maximum(np.abs(xmin), xmax)
    scale = (xmax / maxq).astype(F32)
    scale[xmax == 0] = 1
    return scale


def dynamic_quant_codes(x, scale, bits=8):
    """sym_quant (quant_utils.py:42-45) with a per-token fp32 scale: x (fp16) / scale (fp32)
    promotes to fp32."""
    maxq = 2 ** (bits - 1) - 1
    x2 = np.asarray(x).reshape(-1, np.asarray(x).shape[-1]).astype(F32)
    return np.clip(np.rint(x2 / scale[:, None]), -(maxq + 1), maxq).astype(F32)
