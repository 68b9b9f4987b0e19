"""Thin torch-facing wrappers over the C ABI: device pointers in, torch tensors out.

torch is plumbing here (allocation, streams); all arithmetic happens in libmquant_b200.so.
Every op requires CUDA tensors and raises otherwise -- there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import MQ_FLAG_FAST_HAD, MQ_FLAG_FP32_HAD, MQ_FLAG_SPLIT, check, lib


import functools


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Run the op on the device (and that device's current stream) of its tensor arguments, not on
    whatever device happens to be current: models sharded over several GPUs in one process
    (device_map / split_model) call wrappers whose tensors live on cuda:1 while cuda:0 is current.
    All tensor arguments must share one device."""

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for t in list(args) + list(kwargs.values()):
            if torch.is_tensor(t) and t.is_cuda:
                if dev is None:
                    dev = t.device
                elif t.device != dev:
                    raise _lib.MQuantLibraryError(
                        f"{fn.__name__}: tensor arguments on different devices ({dev} and {t.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MQuantLibraryError("mquant_b200 ops need CUDA tensors (no CPU fallback)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _as_rows(x):
    """[..., C] -> 2-D view with unit column stride (copy only if the layout forces it)."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
        x2 = x2.contiguous()
    return x2


def _ld(x2):
    return x2.stride(0) if x2.shape[0] > 1 else max(x2.shape[1], x2.stride(0))


def make_scales(s_text, s_visual=None, device=None):
    """float32[2] device tensor {s_text, s_visual} from python floats or 0-dim tensors."""
    if s_visual is None:
        s_visual = s_text
    vals = []
    for s in (s_text, s_visual):
        if torch.is_tensor(s):
            vals.append(s.detach().to(device=device, dtype=torch.float32).reshape(()))
        else:
            vals.append(torch.tensor(float(s), dtype=torch.float32, device=device))
    return torch.stack(vals)


@_on_device
def act_quant_i8(x, scales, modality=None, n_out=None, split=False, want_row_scales=True):
    """mq_act_quant_i8.  x: [..., C] fp16.  Returns (codes int8 [M, n_out], s_row fp32 [M] | None,
    x0 fp16 [M] | None)."""
    _need_cuda(x, scales, modality)
    assert x.dtype == torch.float16, "activations must be fp16"
    x2 = _as_rows(x)
    M, C = x2.shape
    n_out = C if n_out is None else n_out
    q = torch.empty((M, n_out), dtype=torch.int8, device=x.device)
    srow = torch.empty((M,), dtype=torch.float32, device=x.device) if want_row_scales else None
    x0 = torch.empty((M,), dtype=torch.float16, device=x.device) if split else None
    check(lib().mq_act_quant_i8(_ptr(x2), _ld(x2), _ptr(q), _ptr(srow), _ptr(x0), _ptr(modality),
                                _ptr(scales), M, C, n_out, MQ_FLAG_SPLIT if split else 0, _stream()),
          "mq_act_quant_i8")
    return q, srow, x0


@_on_device
def gptq_hessian_update(x, H, nsamples_before, batch):
    """mq_gptq_hessian_update: GPTQ.add_batch (fake_quant/gptq/gptq_utils.py:182-195) on the rows of
    x ([T, C] fp16): H <- H * n/(n+t) + 2/(n+t) X^T X with n = nsamples_before, t = batch.  In place."""
    _need_cuda(x, H)
    assert x.dtype == torch.float16 and H.dtype == torch.float32 and H.is_contiguous()
    x2 = _as_rows(x)
    T, C = x2.shape
    assert H.shape == (C, C)
    n1 = nsamples_before + batch
    check(lib().mq_gptq_hessian_update(_ptr(x2), _ld(x2), T, C, _ptr(H), float(nsamples_before) / n1, 2.0 / n1, _stream()),
          "mq_gptq_hessian_update")
    return H


@_on_device
def gptq_block(W, i1, i2, Hinv, scale, maxq):
    """mq_gptq_block_f32 on the columns i1:i2 of W ([rows, columns] fp32, updated in place) with the
    upper Cholesky factor Hinv.  Returns (Q1, Err1, Losses1), each [rows, i2 - i1]."""
    _need_cuda(W, Hinv, scale)
    assert W.dtype == torch.float32 and Hinv.dtype == torch.float32 and W.stride(1) == 1 and Hinv.stride(1) == 1
    rows, count = W.shape[0], i2 - i1
    scale = scale.reshape(-1).to(torch.float32).contiguous()
    assert scale.numel() == rows
    Q = torch.empty((rows, count), dtype=torch.float32, device=W.device)
    Err = torch.empty_like(Q)
    Losses = torch.empty_like(Q)
    wv, hv = W[:, i1:i2], Hinv[i1:i2, i1:i2]
    check(lib().mq_gptq_block_f32(wv.data_ptr(), W.stride(0), hv.data_ptr(), Hinv.stride(0), _ptr(scale), float(maxq),
                                  _ptr(Q), _ptr(Err), _ptr(Losses), rows, count, _stream()), "mq_gptq_block_f32")
    return Q, Err, Losses


@_on_device
def rmsn_quant_i8(x, mean_dim, eps, scales=None, modality=None, n_out=None, want_codes=True, want_f16=False,
                  want_row_scales=True):
    """mq_rmsn_quant_i8: RMSN (fake_quant/module_util.py:42-61) fused with the static INT8 quantiser.
    x: [..., C] fp16.  Returns dict(y=fp16 [M, C] | None, q=int8 [M, n_out] | None, s_row)."""
    _need_cuda(x, scales, modality)
    assert x.dtype == torch.float16, "activations must be fp16"
    x2 = _as_rows(x)
    M, C = x2.shape
    n_out = C if n_out is None else n_out
    y = torch.empty((M, C), dtype=torch.float16, device=x.device) if want_f16 else None
    q = torch.empty((M, n_out), dtype=torch.int8, device=x.device) if want_codes else None
    srow = torch.empty((M,), dtype=torch.float32, device=x.device) if (want_codes and want_row_scales) else None
    check(lib().mq_rmsn_quant_i8(_ptr(x2), _ld(x2), _ptr(q), _ptr(y), _ptr(srow), _ptr(modality), _ptr(scales), M, C,
                                 n_out, float(mean_dim), float(eps), _stream()), "mq_rmsn_quant_i8")
    return {"y": y, "q": q, "s_row": srow}


@_on_device
def act_fake_quant(x, scales, modality=None, bits=8, per_channel=False, out=None):
    """mq_act_fake_quant_f16: quant->dequant, same shape/dtype as x."""
    _need_cuda(x, scales, modality)
    assert x.dtype == torch.float16
    if x.shape[-1] % 8 != 0 and not per_channel and modality is None:
        # one scale for the whole tensor and a last dimension the 128-bit accesses cannot take (the
        # patch-embedding convolutions see [B, 3, H, W]): element-wise, so any 2-D view will do --
        # flatten, pad to a multiple of 8 and pick the widest power-of-two row that divides it
        flat = x.reshape(-1)
        n = flat.numel()
        n8 = (n + 7) // 8 * 8
        if n8 != n:
            flat = torch.nn.functional.pad(flat, (0, n8 - n))
        width = min(n8 & -n8, 4096)
        y = act_fake_quant(flat.reshape(-1, width), scales, None, bits, False).reshape(-1)[:n].reshape(x.shape)
        if out is not None:
            out.copy_(y)
            return out
        return y
    x2 = _as_rows(x)
    M, C = x2.shape
    y = torch.empty((M, C), dtype=torch.float16, device=x.device) if out is None else out.reshape(M, C)
    check(lib().mq_act_fake_quant_f16(_ptr(x2), _ld(x2), _ptr(y), _ld(y), _ptr(modality), _ptr(scales),
                                      1 if per_channel else 0, bits, M, C, _stream()),
          "mq_act_fake_quant_f16")
    return y.reshape(x.shape)


@_on_device
def act_quant_i8_dynamic(x, clip_ratio=1.0, n_out=None):
    _need_cuda(x)
    assert x.dtype == torch.float16
    x2 = _as_rows(x)
    M, C = x2.shape
    n_out = C if n_out is None else n_out
    q = torch.empty((M, n_out), dtype=torch.int8, device=x.device)
    srow = torch.empty((M,), dtype=torch.float32, device=x.device)
    check(lib().mq_act_quant_i8_dynamic(_ptr(x2), _ld(x2), _ptr(q), _ptr(srow), M, C, n_out,
                                        float(clip_ratio), _stream()), "mq_act_quant_i8_dynamic")
    return q, srow


@_on_device
def hadamard(x, had_k=1, hadk=None, n=None, scales=None, modality=None, want_codes=False,
             want_f16=True, split=False, fp32_had=False, want_row_scales=True, fast_had=False):
    """mq_hadamard_quant_i8.  Returns dict(y=fp16 [M,n] | None, q=int8 [M,n] | None, s_row, x0).
    fast_had: MQ_FLAG_FAST_HAD (n / had_k in {256, 512}: moved rounding point, see include/mquant_b200.h)."""
    _need_cuda(x, hadk, scales, modality)
    assert x.dtype == torch.float16
    x2 = _as_rows(x)
    M, C = x2.shape
    n = C if n is None else n
    dev = x.device
    y = torch.empty((M, n), dtype=torch.float16, device=dev) if want_f16 else None
    q = torch.empty((M, n), dtype=torch.int8, device=dev) if want_codes else None
    srow = torch.empty((M,), dtype=torch.float32, device=dev) if (want_codes and want_row_scales) else None
    x0 = torch.empty((M,), dtype=torch.float16, device=dev) if split else None
    flags = (MQ_FLAG_SPLIT if split else 0) | (MQ_FLAG_FP32_HAD if fp32_had else 0) | (MQ_FLAG_FAST_HAD if fast_had else 0)
    if hadk is not None:
        assert hadk.dtype == torch.float16 and hadk.is_contiguous() and hadk.shape == (had_k, had_k)
    check(lib().mq_hadamard_quant_i8(_ptr(x2), _ld(x2), _ptr(q), _ptr(y), _ptr(srow), _ptr(x0),
                                     _ptr(modality), _ptr(scales), _ptr(hadk), had_k, M, C, n, flags,
                                     _stream()), "mq_hadamard_quant_i8")
    return {"y": y, "q": q, "s_row": srow, "x0": x0}


@_on_device
def hadamard_transform(x, scale):
    """mq_hadamard_transform_f16: y = fp16(fp32(x . H_n) * scale) along the last dim (the seam of
    fast_hadamard_transform.hadamard_transform); any leading shape, result contiguous."""
    _need_cuda(x)
    assert x.dtype == torch.float16
    x2 = _as_rows(x)
    M, n = x2.shape
    y = torch.empty((M, n), dtype=torch.float16, device=x.device)
    check(lib().mq_hadamard_transform_f16(_ptr(x2), _ld(x2), _ptr(y), n, M, n, float(scale), _stream()),
          "mq_hadamard_transform_f16")
    return y.reshape(x.shape)


@_on_device
def observer_minmax(x, state, per_channel=False):
    """mq_observer_minmax: merge x's min/max into `state` (fp32 [2] or [2*C], zero-initialised)."""
    _need_cuda(x, state)
    assert x.dtype == torch.float16 and state.dtype == torch.float32
    x2 = _as_rows(x)
    M, C = x2.shape
    assert state.numel() == (2 * C if per_channel else 2)
    check(lib().mq_observer_minmax(_ptr(x2), _ld(x2), M, C, 1 if per_channel else 0, _ptr(state),
                                   _stream()), "mq_observer_minmax")
    return state


@_on_device
def observer_minmax_msq(x, modality, state_text, state_visual, seen):
    """mq_observer_minmax_msq: one pass, rows with modality == 0 -> state_text, others ->
    state_visual (fp32 [2] each), row counts added to `seen` (int32 [2])."""
    _need_cuda(x, modality, state_text, state_visual, seen)
    assert x.dtype == torch.float16 and modality.dtype == torch.uint8 and seen.dtype == torch.int32
    x2 = _as_rows(x)
    M, C = x2.shape
    assert modality.numel() == M and state_text.numel() == 2 and state_visual.numel() == 2 and seen.numel() == 2
    check(lib().mq_observer_minmax_msq(_ptr(x2), _ld(x2), M, C, _ptr(modality), _ptr(state_text),
                                       _ptr(state_visual), _ptr(seen), _stream()), "mq_observer_minmax_msq")


_pct_ws = {}


@_on_device
def observer_percentile(x, alpha=0.99999):
    """mq_observer_percentile -> fp32 [2] = {quantile(1-alpha), quantile(alpha)}."""
    _need_cuda(x)
    assert x.dtype == torch.float16
    x2 = _as_rows(x)
    M, C = x2.shape
    nbytes = lib().mq_observer_percentile_workspace()
    key = (x.device, torch.cuda.current_stream().cuda_stream)
    ws = _pct_ws.get(key)
    if ws is None:
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        _pct_ws[key] = ws
    out = torch.empty((2,), dtype=torch.float32, device=x.device)
    check(lib().mq_observer_percentile(_ptr(x2), _ld(x2), M, C, float(alpha), _ptr(out), _ptr(ws),
                                       nbytes, _stream()), "mq_observer_percentile")
    return out


@_on_device
def weight_quant_rtn(W, bits=4, scales=None):
    """mq_weight_quant_rtn.  W: [N, K] fp16.  Returns (codes int8 [N,K], s_w fp32 [N])."""
    _need_cuda(W, scales)
    assert W.dtype == torch.float16 and W.dim() == 2
    W = W.contiguous()
    N, K = W.shape
    codes = torch.empty((N, K), dtype=torch.int8, device=W.device)
    s_w = torch.empty((N,), dtype=torch.float32, device=W.device)
    if scales is not None:
        scales = scales.detach().to(torch.float32).reshape(-1).contiguous()
        assert scales.numel() == N
    check(lib().mq_weight_quant_rtn(_ptr(W), W.stride(0), _ptr(codes), _ptr(s_w), _ptr(scales), N, K,
                                    bits, _stream()), "mq_weight_quant_rtn")
    return codes, s_w


@_on_device
def pack_w4(codes, zero_col0=False):
    _need_cuda(codes)
    assert codes.dtype == torch.int8 and codes.dim() == 2
    codes = codes.contiguous()
    N, K = codes.shape
    nbytes = lib().mq_w4_packed_bytes(N, K)
    if nbytes == 0:
        raise _lib.MQuantLibraryError(f"pack_w4: K={K} must be a positive multiple of 128")
    w4 = torch.empty((nbytes,), dtype=torch.uint8, device=codes.device)
    check(lib().mq_pack_w4(_ptr(codes), codes.stride(0), _ptr(w4), N, K, MQ_FLAG_SPLIT if zero_col0 else 0,
                           _stream()), "mq_pack_w4")
    return w4


@_on_device
def repack_w4_from_ref(ref_packed, N, K):
    _need_cuda(ref_packed)
    assert ref_packed.dtype == torch.uint8 and ref_packed.shape == (N, K // 2)
    ref_packed = ref_packed.contiguous()
    nbytes = lib().mq_w4_packed_bytes(N, K)
    w4 = torch.empty((nbytes,), dtype=torch.uint8, device=ref_packed.device)
    check(lib().mq_repack_w4_from_ref(_ptr(ref_packed), _ptr(w4), N, K, _stream()), "mq_repack_w4_from_ref")
    return w4


@_on_device
def unpack_w4(w4, N, K):
    _need_cuda(w4)
    codes = torch.empty((N, K), dtype=torch.int8, device=w4.device)
    check(lib().mq_unpack_w4(_ptr(w4), _ptr(codes), K, N, K, _stream()), "mq_unpack_w4")
    return codes


@_on_device
def w4a8_gemm(q, w4, s_a, s_w, N, bias=None, x0=None, w0=None, out=None, simt=False):
    """mq_w4a8_gemm.  q: int8 [M, K]; s_a: fp32 [M] (per row) or [1]; s_w fp32 [N]."""
    _need_cuda(q, w4, s_a, s_w, bias, x0, w0)
    assert q.dtype == torch.int8 and q.dim() == 2 and q.stride(1) == 1
    M, K = q.shape
    if s_a.numel() not in (1, M):
        raise _lib.MQuantLibraryError("w4a8_gemm: s_a must have 1 or M elements")
    per_row = int(s_a.numel() == M)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float16, device=q.device)
    lda = q.stride(0) if M > 1 else max(K, q.stride(0))
    ldo = out.stride(0) if M > 1 else max(N, out.stride(0))
    if simt:
        check(lib().mq_w4a8_gemm_simt(_ptr(q), lda, _ptr(w4), _ptr(s_a), per_row, _ptr(s_w), _ptr(bias), _ptr(x0),
                                      _ptr(w0), _ptr(out), ldo, M, N, K, _stream()), "mq_w4a8_gemm_simt")
        return out
    ws = _gemm_workspace(q.device, M, N, K)
    check(lib().mq_w4a8_gemm_ws(_ptr(q), lda, _ptr(w4), _ptr(s_a), per_row, _ptr(s_w), _ptr(bias), _ptr(x0),
                                _ptr(w0), _ptr(out), ldo, M, N, K, _ptr(ws), 0 if ws is None else ws.numel(),
                                _stream()), "mq_w4a8_gemm")
    return out


def interleave_gated(v1, v2):
    """Rows (or entries) of w1 / w2 in the order of the gated GEMM: blocks of 128 of v1 and v2 alternate
    (mq_w4a8_gemm_gated_ws).  v1, v2: [F, ...] with F % 128 == 0."""
    F = v1.shape[0]
    if F % 128 != 0 or v2.shape != v1.shape:
        raise _lib.MQuantLibraryError("interleave_gated: two equal shapes with a multiple of 128 rows")
    rest = v1.shape[1:]
    return torch.stack((v1.reshape(F // 128, 128, *rest), v2.reshape(F // 128, 128, *rest)), dim=1).reshape(2 * F, *rest)


@_on_device
def silu_mul_interleaved(a, F, out=None):
    """mq_silu_mul_interleaved_f16: a = [M, 2F] output of a GEMM over interleave_gated weights -> [M, F]."""
    _need_cuda(a)
    assert a.dtype == torch.float16 and a.dim() == 2 and a.shape[1] == 2 * F and a.stride(1) == 1
    M = a.shape[0]
    if out is None:
        out = torch.empty((M, F), dtype=torch.float16, device=a.device)
    check(lib().mq_silu_mul_interleaved_f16(_ptr(a), _ld(a), _ptr(out), _ld(out), M, F, _stream()),
          "mq_silu_mul_interleaved_f16")
    return out


@_on_device
def w4a8_gemm_gated(q, w4, s_a, s_w, F, bias=None, out=None):
    """mq_w4a8_gemm_gated_ws: w4 / s_w / bias over interleave_gated(w1, w2) rows (N2 = 2F) -> fp16 [M, F] =
    a1 * silu(a2) (QWenMLP.forward, model/modeling_qwen_opt.py:382-386)."""
    _need_cuda(q, w4, s_a, s_w, bias)
    assert q.dtype == torch.int8 and q.dim() == 2 and q.stride(1) == 1
    M, K = q.shape
    if s_a.numel() not in (1, M):
        raise _lib.MQuantLibraryError("w4a8_gemm_gated: s_a must have 1 or M elements")
    if out is None:
        out = torch.empty((M, F), dtype=torch.float16, device=q.device)
    ws = _gemm_workspace(q.device, M, 2 * F, K)
    nws = 0 if ws is None else ws.numel()
    tmp = None
    if M > 0 and lib().mq_w4a8_gemm_plan(M, 2 * F, K, nws) != 4:
        tmp = torch.empty((M, 2 * F), dtype=torch.float16, device=q.device)
    lda = q.stride(0) if M > 1 else max(K, q.stride(0))
    ldo = out.stride(0) if M > 1 else max(F, out.stride(0))
    check(lib().mq_w4a8_gemm_gated_ws(_ptr(q), lda, _ptr(w4), _ptr(s_a), int(s_a.numel() == M), _ptr(s_w), _ptr(bias),
                                      _ptr(out), ldo, M, 2 * F, K, _ptr(tmp), _ptr(ws), nws, _stream()),
          "mq_w4a8_gemm_gated")
    return out


_gemm_ws = {}


def _gemm_workspace(device, M, N, K):
    """Zero-initialised, self-cleaning split-K workspace for decode-sized GEMMs (one per device and
    stream, grown on demand)."""
    need = lib().mq_w4a8_gemm_workspace_bytes(M, N, K)
    if need == 0:
        return None
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    ws = _gemm_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _gemm_ws[key] = ws
    return ws
