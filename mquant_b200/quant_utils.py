"""Quantized-linear path: API mirror of fake_quant/quant_utils.py:1-720 on B200 kernels.

Same names, attributes and call protocol as the reference so that exam/quant_*.py and
fake_quant/gptq/* run on top unchanged:

  ActQuantizer      quant_utils.py:97-268     configure / forward / find_params / quantize / free
  ActQuantWrapper   quant_utils.py:271-391    == QuantLinear; .module .weight .bias .quantizer
                                              .out_quantizer .had_K .K .online_full_had
                                              .online_partial_had .had_dim .fp32_had .split .L1 .L2
  WeightQuantizer   quant_utils.py:415-524    configure / find_params / quantize / enabled / ready
  add_actquant, find_qlayers, *_add_act_qaunt, model_open_calibrate, ...   quant_utils.py:543-720

What is new (additive):
  * the wrapper has a REAL integer path: ``freeze()`` packs the (already fake-quantised) weight
    into INT4 and from then on ``forward`` = [online Hadamard +] static INT8 quantisation (one
    fused kernel) -> tcgen05 W4A8 GEMM with the scale/bias/rank-1 epilogue.  Before ``freeze()``
    the wrapper computes what the reference computes (fp16 fake-quant activations into a torch
    Linear) so calibration and GPTQ see the same tensors.
  * MSQ: ``ActQuantizer.configure(..., msq=True)`` keeps separate text / visual statistics and
    scales; the per-token flag comes from ``mquant_b200.msq.modality_context``.
There is no CPU execution path: forward on CPU tensors raises.
"""
from __future__ import annotations

import math
import warnings
from collections import OrderedDict

import torch

from . import hadamard_utils, msq, ops
from .bit_type import BIT_TYPE_DICT
from .observer import build_observer
from .quantizer import build_quantizer


def get_minq_maxq(bits, sym):
    if sym:
        maxq = torch.tensor(2 ** (bits - 1) - 1)
        minq = -maxq - 1
    else:
        maxq = torch.tensor(2 ** bits - 1)
        minq = 0
    return minq, maxq


# ---- integer-code helpers (reference :27-94) -------------------------------------------------

def asym_quant(x, scale, zero, maxq):
    scale = scale.to(x.device)
    zero = zero.to(x.device)
    q = torch.clamp(torch.round(x / scale) + zero, 0, maxq)
    return q, scale, zero


def asym_dequant(q, scale, zero):
    return scale * (q - zero)


def asym_quant_dequant(x, scale, zero, maxq):
    return asym_dequant(*asym_quant(x, scale, zero, maxq))


def sym_quant(x, scale, maxq):
    scale = scale.to(x.device)
    q = torch.clamp(torch.round(x / scale), -(maxq + 1), maxq)
    return q, scale


def sym_dequant(q, scale):
    return scale * q


def sym_quant_dequant(x, scale, maxq):
    return sym_dequant(*sym_quant(x, scale, maxq))


def two_compl(x, bits: int):
    return torch.where(x < 0, 2 ** bits + x, x)


def pack_i4(q):
    """Interchange format of the reference (:61-68): even index -> low nibble."""
    assert torch.is_signed(q), "The tensor to be packed should be signed int"
    assert torch.all(torch.logical_and(q >= -8, q <= 7))
    u = two_compl(q.to(dtype=torch.int8), 4).to(torch.uint8)
    return u[:, 0::2] | (u[:, 1::2] << 4)


def unpack_i4(x: torch.Tensor):
    assert x.dtype == torch.uint8, "The tensor to be unpacked should be stored in uint8"
    lo = (x & 0x0F).to(torch.int8)
    lo = torch.where(lo >= 8, lo - 16, lo)
    hi = ((x & 0xF0) >> 4).to(torch.int8)
    hi = torch.where(hi >= 8, hi - 16, hi)
    return torch.stack((lo, hi), dim=-1).reshape(*x.shape[:-1], x.shape[-1] * 2).to(torch.int32)


# ---- activation quantizer ---------------------------------------------------------------------

def _observed(observer):
    """Has this observer seen data?  (Without materialising the statistics.)"""
    state = getattr(observer, "state", None)
    return state is not None if hasattr(observer, "state") else observer.max_val is not None


class ActQuantizer(torch.nn.Module):
    """Reference quant_utils.py:97-268."""

    def __init__(self, act_per_tensor=False):
        super().__init__()
        self.register_buffer("maxq", torch.tensor(0))
        self.register_buffer("scale", torch.zeros(1))
        self.register_buffer("zero", torch.zeros(1))
        self.bits = 16
        self.act_per_tensor = act_per_tensor
        self.static = False
        self.msq = False
        self.sym = True

    def free(self):
        self.zero = None
        self.scale = None
        self._dyn = None

    def configure(self, bits, groupsize=-1, sym=False, clip_ratio=1.0, act_per_tensor=False, static=False,
                  observer_type="minmax", calibration_mode="layer_wise", msq=False):
        _, self.maxq = get_minq_maxq(bits, sym)
        self.bits = bits
        self.groupsize = groupsize
        self.sym = sym
        self.clip_ratio = clip_ratio
        self.act_per_tensor = act_per_tensor
        assert 0 < self.clip_ratio <= 1, "Clip ratio should be in (0, 1]"
        self.static = static
        self.msq = bool(msq) and static
        self._fast_obs = None
        if self.static:
            bit_type_a = BIT_TYPE_DICT[f"int{bits}"]  # KeyError for unsupported widths, as the reference
            self.observer = build_observer(observer_type, "activation", bit_type_a, calibration_mode)
            self.quantizer = build_quantizer("uniform", bit_type_a, self.observer, "activation")
            self.observer_visual = (build_observer(observer_type, "activation", bit_type_a, calibration_mode)
                                    if self.msq else None)
            self.calibrate = False
            self.last_calibrate = False
            self.quant = False

    # -- static path ---------------------------------------------------------------------------
    def _modality(self, x):
        if not self.msq:
            return None
        ntok = x.numel() // x.shape[-1]
        return msq.current_modality(ntok, x.device)

    def _observe(self, x):
        # Lean path for the common calibration call (fused MSQ min/max, fp16, states allocated):
        # calibration of ~1000-token samples is bound by the host cost of a wrapper call, not by the
        # 3 us kernel -- skip the generic argument plumbing of ops.* (pointers cached once).
        fast = getattr(self, "_fast_obs", None)
        if fast is not None and x.dtype is torch.float16 and x.is_cuda and x.is_contiguous():
            mask = getattr(msq._state, "mask", None)
            C = x.shape[-1]
            if (mask is not None and mask.device == x.device and mask.numel() * C == x.numel()
                    and C == fast[0] and x.device.index == torch.cuda.current_device()
                    and self.observer.state is fast[5] and self.observer_visual.state is fast[6]
                    and self._seen is fast[7]):  # (somebody may have installed other statistics)
                rc = fast[1](x.data_ptr(), C, mask.numel(), C, mask.data_ptr(), fast[2], fast[3], fast[4],
                             torch.cuda.current_stream().cuda_stream)
                if rc != 0:
                    ops.check(rc, "mq_observer_minmax_msq")
                return
        mod = self._modality(x)
        if mod is None:
            self.observer.update(x)
            return
        x2 = x.reshape(-1, x.shape[-1])
        if x2.is_cuda and x2.dtype == torch.bfloat16:
            for o in (self.observer, self.observer_visual):
                if getattr(o, "_dtype", None) is None and hasattr(o, "_dtype"):
                    o._dtype = x.dtype  # statistics (and the scale) keep the activation dtype
            x2 = x2.clamp(-65504.0, 65504.0).to(torch.float16)
        if self._fused_msq_ok(x2):
            # one kernel: both modalities' min/max and the row counts of either kind, no gathers
            # and no device round trip (mq_observer_minmax_msq)
            for o in (self.observer, self.observer_visual):
                if o.state is None:
                    o.state = torch.zeros(2, dtype=torch.float32, device=x.device)
                    o._dtype, o._channels = (o._dtype or x.dtype), x2.shape[-1]
            if getattr(self, "_seen", None) is None:
                self._seen = torch.zeros(2, dtype=torch.int32, device=x.device)
            ops.observer_minmax_msq(x2, mod, self.observer.state, self.observer_visual.state, self._seen)
            # (the tensors behind these pointers live as long as the observers do)
            self._fast_obs = (x2.shape[-1], ops.lib().mq_observer_minmax_msq, self.observer.state.data_ptr(),
                              self.observer_visual.state.data_ptr(), self._seen.data_ptr(),
                              self.observer.state, self.observer_visual.state, self._seen)
            return
        vis = mod.bool()
        if bool((~vis).any()):
            self.observer.update(x2[~vis])
        if bool(vis.any()):
            self.observer_visual.update(x2[vis])

    def _fused_msq_ok(self, x2):
        from .observer.minmax import MinmaxObserver
        return (type(self.observer) is MinmaxObserver and type(self.observer_visual) is MinmaxObserver
                and not self.observer._per_channel() and x2.is_cuda and x2.dtype == torch.float16
                and x2.shape[-1] % 8 == 0 and x2.stride(-1) == 1)

    def _seen_kinds(self):
        """(text seen, visual seen).  The fused observer counts rows on the device: one read here,
        when the parameters are derived, instead of one per observed batch."""
        has_t = _observed(self.observer)
        has_v = self.msq and _observed(self.observer_visual)
        seen = getattr(self, "_seen", None)
        if seen is not None:
            n_t, n_v = seen.tolist()
            has_t, has_v = has_t and n_t > 0, has_v and n_v > 0
        return has_t, has_v

    def update_quantization_params(self):
        has_t, has_v = self._seen_kinds()
        if has_t:
            self.quantizer.update_quantization_params()
        if has_v:
            sv, _ = self.observer_visual.get_quantization_params()
            if has_t:
                self.quantizer.scale_visual = sv
                self.quantizer._dev_scales = None
            else:  # only visual tokens were ever seen
                self.quantizer.set_scales(sv, sv)

    def forward(self, x):
        if self.static:
            if self.calibrate:
                self._observe(x)
                if self.last_calibrate:
                    self.update_quantization_params()
                return x
            elif self.quant:
                return self.quantizer(x, self._modality(x))
            return x
        if self.bits == 16:
            return x
        if not (self.sym and self.bits == 8 and self.groupsize <= 0 and not self.act_per_tensor):
            raise NotImplementedError("dynamic activation quantisation: only per-token symmetric INT8 "
                                      "has a kernel (the MQuant recipes use static INT8)")
        q, s = ops.act_quant_i8_dynamic(x, self.clip_ratio)
        return (q.float() * s[:, None]).to(x.dtype).reshape(x.shape)

    def quantize(self, x):
        """Integer codes and scales (reference :136-140)."""
        if self.static:
            return self.quantizer.quant(x, modality=self._modality(x)), self.quantizer.scale
        q, s = ops.act_quant_i8_dynamic(x, self.clip_ratio)
        return q.reshape(x.shape), s.reshape(*x.shape[:-1], 1)

    def find_params(self, x):
        """Dynamic per-token scales (reference :205-268); kept for API compatibility: the codes
        and scales are produced together by one kernel in forward()/quantize()."""
        if self.bits == 16:
            return
        _, s = ops.act_quant_i8_dynamic(x, self.clip_ratio)
        self.scale = s.reshape(*x.shape[:-1], 1)
        self.zero = torch.zeros_like(self.scale)


# ---- the wrapper ------------------------------------------------------------------------------

def _prequant_of(x, scales, n_out, modality):
    """INT8 codes an RMSNQuant attached to its output (module_util.py), if they are exactly what
    this consumer would compute: same scale buffer, same padded width, same modality mask."""
    pre = getattr(x, "_mq_prequant", None)
    if pre is not None and pre.matches(scales, n_out, modality):
        return pre.q, pre.s_row
    return None


def _round_up(a, b):
    return (a + b - 1) // b * b


class ActQuantWrapper(torch.nn.Module):
    """Reference quant_utils.py:271-391, plus the real-integer path (see module docstring)."""

    def __init__(self, module, act_per_tensor=False):
        super().__init__()
        assert isinstance(module, (torch.nn.Linear, torch.nn.Conv2d, torch.nn.Conv3d))
        self.module = module
        self.weight = module.weight
        self.bias = module.bias
        self.quantizer = ActQuantizer(act_per_tensor)
        self.out_quantizer = ActQuantizer(act_per_tensor)
        self.register_buffer("had_K", torch.tensor(0))
        self._buffers["had_K"] = None
        self.K = 1
        self.online_full_had = False
        self.online_partial_had = False
        self.had_dim = 0
        self.fp32_had = False
        self.fast_had = False  # opt-in MQ_FLAG_FAST_HAD (n / K in {256, 512}: moved fp16 rounding point, 3x faster)
        self.split = False
        # new
        self.pad_to = None       # fold utils.revise_down_input into the fused kernel
        self.real_quant = False  # set by freeze()
        self._packed = None
        self._group = None       # set by fuse_siblings()

    def extra_repr(self) -> str:
        s = f"Input Quantizer Bits: {self.quantizer.bits}"
        if self.quantizer.bits < 16:
            s += " (Symmetric" if self.quantizer.sym else " (Asymmetric"
            s += " Static Per-Tensor)" if self.quantizer.static else " Per-Token)"
        s += f"\nOutput Quantizer Bits: {self.out_quantizer.bits}"
        if self.real_quant:
            s += "\nW4A8 real-integer path (tcgen05)"
        return s

    def split_weights(self):
        """RMS split (reference :316-328): channel 0 gets its own fp Linear."""
        m = self.module
        self.L1 = torch.nn.Linear(1, m.out_features, bias=False).to(m.weight.device)
        self.L2 = torch.nn.Linear(m.in_features - 1, m.out_features, bias=m.bias is not None).to(m.weight.device)
        self.L1.weight.data = m.weight.data[:, 0:1]
        self.L2.weight.data = m.weight.data[:, 1:]
        if m.bias is not None:
            self.L2.bias.data = m.bias.data

    # -- real-integer path -----------------------------------------------------------------------
    @torch.no_grad()
    def freeze(self, w_scale=None, w_bits=4, check=False):
        """Pack the current (fake-quantised, fp16) weight into INT4 for the W4A8 GEMM.

        w_scale: the per-output-channel scale the weight was quantised with
        (``WeightQuantizer.scale``, shape [N,1]); codes = round(W / scale) are then recovered
        exactly (SURVEY.md 8(c) self-check).  If omitted the weight is re-quantised RTN
        (quant_utils.py:446-466) -- identical codes when the weight already sits on a 4-bit grid.
        check=True verifies that the packed integers reproduce the current weight exactly and raises
        otherwise (model_freeze turns it on whenever weight quantizers were supplied).
        """
        if not isinstance(self.module, torch.nn.Linear):
            raise NotImplementedError("freeze(): only Linear layers have an integer path")
        if not (self.quantizer.static and self.quantizer.bits == 8):
            raise RuntimeError("freeze(): configure the activation quantizer as static INT8 first")
        if w_bits != 4:
            raise NotImplementedError("freeze(): the GEMM takes INT4 weights")
        if self.split:
            W = torch.cat([self.L1.weight.data.to(torch.float16), self.L2.weight.data.to(torch.float16)], 1)
            bias = self.L2.bias
        else:
            W = self.module.weight.data.to(torch.float16)
            bias = self.module.bias
        if not W.is_cuda:
            raise ops._lib.MQuantLibraryError("freeze() needs the weights on a CUDA device")
        N, K = W.shape
        if w_scale is not None:
            w_scale = w_scale.to(W.device)
        if self.split:
            body = W.clone()
            body[:, 0] = 0  # channel 0 is the fp rank-1 term; it must not steer the RTN scale
            codes, s_w = ops.weight_quant_rtn(body, w_bits, w_scale)
            w0 = W[:, 0].contiguous()
        else:
            body = W
            codes, s_w = ops.weight_quant_rtn(W, w_bits, w_scale)
            w0 = None
        if check:
            # the integer weights must BE the weights that were evaluated: a fake-quantised weight
            # sits on its grid, so codes * scale reproduces it (to fp16 rounding).  A GPTQ / clipped
            # weight re-quantised without its scale (max|W|/7 instead) does not.
            orig = (torch.cat([self.L1.weight.data, self.L2.weight.data], 1) if self.split
                    else self.module.weight.data)
            # (L1 / L2 turn fp32 on their first reference-path call, quant_utils.py:374-375: compare
            # what the fp16 weights were)
            cmp_dtype = torch.float16 if orig.dtype == torch.float32 else orig.dtype
            back = (codes.float() * s_w[:, None]).to(cmp_dtype)  # the reference's own expression (:52)
            ref_w = orig.to(cmp_dtype).clone()
            if self.split:
                ref_w[:, 0] = 0
            bad = (back != ref_w).any(dim=1)
            if bool(bad.any()):
                n_bad = int(bad.sum())
                raise RuntimeError(
                    f"freeze(): {n_bad} of {N} output channels are not reproduced by round(W/scale)*scale"
                    + (" with the supplied scale" if w_scale is not None else
                       " with the RTN scale max|W|/7 -- pass the WeightQuantizer scale the weight was "
                       "quantised with (GPTQ / clip search), or check=False to re-quantise RTN"))
        Kp = _round_up(max(K, self.pad_to or 0), 128)
        if Kp != K:
            codes = torch.nn.functional.pad(codes, (0, Kp - K))
        self._packed = {
            "w4": ops.pack_w4(codes, zero_col0=self.split),
            "s_w": s_w, "w0": w0, "N": N, "K": K, "Kp": Kp,
            "bias": None if bias is None else bias.data.to(torch.float16).contiguous(),
        }
        self.real_quant = True
        return self

    def _forward_real(self, x):
        if self._group is not None:
            return self._group.forward(self, x)
        pk = self._packed
        qz = self.quantizer.quantizer
        x2 = x.reshape(-1, x.shape[-1])
        M = x2.shape[0]
        scales = qz.device_scales(x.device)
        modality = self.quantizer._modality(x) if qz.scale_visual is not None else None
        if self.online_partial_had and not self.online_full_had:
            x2 = self._partial_had(x2)
        if self.online_full_had:
            n = self.pad_to or x2.shape[-1]
            hk = hadamard_utils.hadK_device(self.K, x.device) if self.K > 1 else None
            r = ops.hadamard(x2, had_k=self.K, hadk=hk, n=n, scales=scales, modality=modality,
                             want_codes=True, want_f16=False, split=self.split, fp32_had=self.fp32_had,
                             want_row_scales=modality is not None, fast_had=self.fast_had)
            q, srow, x0 = r["q"], r["s_row"], r["x0"]
            if q.shape[1] != pk["Kp"]:
                q = torch.nn.functional.pad(q, (0, pk["Kp"] - q.shape[1]))
        else:
            pre = None if self.split else _prequant_of(x, scales, pk["Kp"], modality)
            if pre is not None:  # codes emitted by the RMSNQuant in front (module_util.py)
                q, srow, x0 = pre[0], pre[1], None
            else:
                q, srow, x0 = ops.act_quant_i8(x2, scales, modality, n_out=pk["Kp"], split=self.split,
                                               want_row_scales=modality is not None)
        s_a = srow if modality is not None else scales[0:1]
        out = ops.w4a8_gemm(q, pk["w4"], s_a, pk["s_w"], pk["N"], bias=pk["bias"], x0=x0, w0=pk["w0"])
        return out.reshape(*x.shape[:-1], pk["N"])

    # -- reference-equivalent (fake-quant) path ---------------------------------------------------
    def _rotate(self, x):
        if self.online_full_had:
            n = self.pad_to or x.shape[-1]
            hk = hadamard_utils.hadK_device(self.K, x.device) if self.K > 1 else None
            if x.dtype not in (torch.float16, torch.bfloat16):
                raise NotImplementedError("online Hadamard: fp16 / bf16 activations only")
            xh = x if x.dtype == torch.float16 else x.clamp(-65504.0, 65504.0).to(torch.float16)
            y = ops.hadamard(xh, had_k=self.K, hadk=hk, n=n, fp32_had=self.fp32_had, fast_had=self.fast_had)["y"]
            return y.reshape(*x.shape[:-1], n).to(x.dtype)
        if self.online_partial_had:
            return self._partial_had(x)
        if self.pad_to and self.pad_to > x.shape[-1]:
            return torch.nn.functional.pad(x, (0, self.pad_to - x.shape[-1]))
        return x

    def _partial_had(self, x):
        """Per-head online Hadamard (reference quant_utils.py:343-365): the transform runs ACROSS
        the n/had_dim heads for each of the had_dim positions, scale 1/sqrt(n/had_dim)."""
        from .fht import hadamard_transform

        x_dtype = x.dtype
        if self.fp32_had:
            x = x.float()
        init_shape = x.shape
        heads = init_shape[-1] // self.had_dim
        if self.K == 1:
            x = hadamard_transform(x.reshape(-1, heads, self.had_dim).transpose(1, 2).contiguous(),
                                   scale=1 / math.sqrt(heads)).transpose(1, 2)
        else:
            x = (self.had_K.to(x.dtype) @ x.reshape(-1, heads, self.had_dim)) / math.sqrt(heads)
        if self.fp32_had:
            x = x.to(x_dtype)
        return x.reshape(init_shape)

    def forward(self, x):
        x_dtype = x.dtype
        q = self.quantizer
        if (self.real_quant and q.static and q.quant and not q.calibrate and self.out_quantizer.bits >= 16):
            if x.dtype == torch.float16:
                return self._forward_real(x)
            if x.dtype == torch.bfloat16:
                # the integer path computes in fp16 in / fp16 out (exact conversion of in-range bf16);
                # a sibling group converts once and keys its cache on the ORIGINAL tensor
                if self._group is not None:
                    return self._group.forward(self, x).to(x_dtype)
                return self._forward_real(x.clamp(-65504.0, 65504.0).to(torch.float16)).to(x_dtype)

        x = self._rotate(x)
        if self.split:
            if q.static:
                if q.calibrate:
                    obs = x.clone()
                    obs[..., 0] = 0  # the reference observes x[..., 1:]; 0 is already folded in
                    q(obs)
                elif q.quant:
                    x0 = x[..., 0].clone()
                    x = q(x)
                    x[..., 0] = x0
            elif q.bits < 16:
                x0 = x[..., 0].clone()
                x = q(x.clone().index_fill_(-1, torch.zeros(1, dtype=torch.long, device=x.device), 0))
                x[..., 0] = x0
            x1 = self.L1.float()(x[..., 0:1].float())
            x2 = self.L2.float()(x[..., 1:].float())
            x = (x1 + x2).to(x_dtype)
        else:
            if q.static or q.bits < 16:
                x = q(x)
            x = self.module(x).to(x_dtype)
        if self.out_quantizer.bits < 16:
            x = self.out_quantizer(x).to(x_dtype)
        return x


QuantLinear = ActQuantWrapper  # the north-star's name for the drop-in unit


class ActRotateWrapper(torch.nn.Module):
    """Reference quant_utils.py:394-412 (dense rotation of two inputs; offline experiments only)."""

    def __init__(self, module, QMatrix):
        super().__init__()
        self.module = module
        self.register_buffer("q_matrix", QMatrix)
        self.fp32_had = False

    def forward(self, x, y):
        x_dtype = x.dtype
        if self.fp32_had:
            x = (x.float() @ self.q_matrix).to(x_dtype)
            y.copy_((y.float() @ self.q_matrix).to(y.dtype))
        else:
            x = x @ self.q_matrix
            y.copy_(y @ self.q_matrix.to(y.dtype))
        return self.module(x, y).to(x_dtype)


# ---- weight quantizer ---------------------------------------------------------------------------

class WeightQuantizer(torch.nn.Module):
    """Reference quant_utils.py:415-524 ("From GPTQ Repo").  Statistics are torch reductions over
    an [N, K] matrix (offline, once per layer); the per-channel symmetric RTN that the W4A8
    recipes use also exists as a kernel (ops.weight_quant_rtn) which ``freeze`` calls."""

    def __init__(self, shape=1):
        super().__init__()
        self.register_buffer("maxq", torch.tensor(0))
        self.register_buffer("scale", torch.zeros(shape))
        self.register_buffer("zero", torch.zeros(shape))

    def configure(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=0.8):
        self.bits = bits
        self.perchannel = perchannel
        self.sym = sym
        self.mse = mse
        self.norm = norm
        self.grid = grid
        self.maxshrink = maxshrink
        self.maxq = torch.tensor(2 ** (bits - 1) - 1) if sym else torch.tensor(2 ** bits - 1)

    def find_params(self, x):
        if self.bits == 16:
            return
        dev = x.device
        self.maxq = self.maxq.to(dev)
        shape = x.shape
        x = x.flatten(1) if self.perchannel else x.flatten().unsqueeze(0)
        zeros = torch.zeros(x.shape[0], device=dev)
        xmin = torch.minimum(x.min(1)[0], zeros)
        xmax = torch.maximum(x.max(1)[0], zeros)
        if self.sym:
            xmax = torch.maximum(torch.abs(xmin), xmax).clamp(min=1e-5)
            self.scale = xmax / self.maxq
            self.zero = torch.zeros_like(self.scale)
        else:
            both0 = (xmin == 0) & (xmax == 0)
            xmin[both0] = -1
            xmax[both0] = +1
            self.scale = (xmax - xmin).clamp(min=1e-5) / self.maxq
            self.zero = torch.round(-xmin / self.scale)
        if self.mse:
            best = torch.full([x.shape[0]], float("inf"), device=dev)
            for i in range(int(self.maxshrink * self.grid)):
                p = 1 - i / self.grid
                xmin1, xmax1 = p * xmin, p * xmax
                if self.sym:
                    scale1 = xmax1 / self.maxq
                    zero1 = torch.zeros_like(scale1)
                    q = sym_quant_dequant(x, scale1.unsqueeze(1), self.maxq)
                else:
                    scale1 = (xmax1 - xmin1) / self.maxq
                    zero1 = torch.round(-xmin1 / scale1)
                    q = asym_quant_dequant(x, scale1.unsqueeze(1), zero1.unsqueeze(1), self.maxq)
                q -= x
                q.abs_()
                q.pow_(self.norm)
                err = torch.sum(q, 1)
                better = err < best
                if torch.any(better):
                    best[better] = err[better]
                    self.scale[better] = scale1[better]
                    self.zero[better] = zero1[better]
        if not self.perchannel:
            self.scale = self.scale.repeat(shape[0])
            self.zero = self.zero.repeat(shape[0])
        bshape = [-1] + [1] * (len(shape) - 1)
        self.scale = self.scale.reshape(bshape)
        self.zero = self.zero.reshape(bshape)

    def quantize(self, x):
        x_dtype = x.dtype
        if self.ready() and self.bits < 16:
            if self.sym:
                return sym_quant_dequant(x, self.scale, self.maxq).to(x_dtype)
            return asym_quant_dequant(x, self.scale, self.zero, self.maxq).to(x_dtype)
        return x

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)


# ---- module-tree surgery and flag broadcast (reference :543-720) ---------------------------------

def add_actquant(module, act_per_tensor=False, name="", layers=[torch.nn.Linear]):
    """Wrap every exact-type match (reference :626-662; `type(tmp) in layers`, Sequential and
    ModuleList children included)."""
    if isinstance(module, ActQuantWrapper):
        return
    for attr in dir(module):
        tmp = getattr(module, attr)
        if type(tmp) in layers:
            setattr(module, attr, ActQuantWrapper(tmp, act_per_tensor))
        if type(tmp) == torch.nn.Sequential:
            replaced = OrderedDict()
            for cname, child in tmp.named_children():
                replaced[cname] = ActQuantWrapper(child, act_per_tensor) if type(child) in layers else child
            setattr(module, attr, torch.nn.Sequential(replaced))
        if type(tmp) == torch.nn.ModuleList:
            replaced = [ActQuantWrapper(c, act_per_tensor) if type(c) in layers else c for c in tmp.children()]
            setattr(module, attr, torch.nn.ModuleList(replaced))
    for name1, child in module.named_children():
        add_actquant(child, act_per_tensor, name + "." + name1 if name != "" else name1, [torch.nn.Linear])


def add_actquant_for_mlp1(module, act_per_tensor=False, name="", layers=[torch.nn.Linear]):
    module.mlp1[1] = ActQuantWrapper(module.mlp1[1], act_per_tensor)
    module.mlp1[3] = ActQuantWrapper(module.mlp1[3], act_per_tensor)


def find_qlayers(module, layers=[torch.nn.Linear, ActQuantWrapper], name=""):
    if type(module) in layers:
        return {name: module}
    res = {}
    for name1, child in module.named_children():
        res.update(find_qlayers(child, layers=layers, name=name + "." + name1 if name != "" else name1))
    return res


def internvl_add_act_qaunt(model, args):
    if args.quant_llm:
        add_actquant(model.model.language_model.model, args.act_per_tensor)
    if args.quant_visual_clip:
        emb = model.model.vision_model.embeddings
        emb.patch_embedding = ActQuantWrapper(emb.patch_embedding, args.act_per_tensor)
        add_actquant(model.model.vision_model.encoder, args.act_per_tensor)
    if args.quant_cross_attention:
        add_actquant_for_mlp1(model.model, args.act_per_tensor)


def qwen2vl_add_act_qaunt(model, args):
    if args.quant_llm:
        add_actquant(model.model.model, args.act_per_tensor)
    if args.quant_visual_clip:
        pe = model.model.visual.patch_embed
        pe.proj = ActQuantWrapper(pe.proj, args.act_per_tensor)
        add_actquant(model.model.visual.blocks, args.act_per_tensor)
    if args.quant_cross_attention:
        add_actquant(model.model.visual.merger, args.act_per_tensor)


def qwenvl_add_act_qaunt(model, args):
    if args.quant_llm:
        add_actquant(model.transformer.h, args.act_per_tensor)
    if args.quant_visual_clip:
        vis = model.transformer.visual
        vis.conv1 = ActQuantWrapper(vis.conv1, args.act_per_tensor)
        add_actquant(vis.transformer, args.act_per_tensor)
    if args.quant_cross_attention:
        add_actquant(model.transformer.visual.attn_pool, args.act_per_tensor)
        model.transformer.visual.proj_fc = ActQuantWrapper(model.transformer.visual.proj_fc, args.act_per_tensor)


def minicpmv_add_act_qaunt(model, args):
    if args.quant_llm:
        add_actquant(model.llm.model.layers, args.act_per_tensor)
    if args.quant_visual_clip:
        emb = model.vpm.embeddings
        emb.patch_embedding = ActQuantWrapper(emb.patch_embedding, args.act_per_tensor)
        add_actquant(model.vpm.encoder, args.act_per_tensor)
    if args.quant_cross_attention:
        add_actquant(model.resampler, args.act_per_tensor)


@torch.no_grad()
def fuse_internvl(model):
    """Fold InternViT layer-scale vectors into the following Linear (reference :527-541)."""
    for layer in model.model.vision_model.encoder.layers:
        layer.attn.proj.weight.data *= layer.ls1.data.view(-1, 1)
        layer.mlp.fc2.weight.data *= layer.ls2.data.view(-1, 1)
        if hasattr(layer.attn.proj, "bias"):
            layer.attn.proj.bias.data *= layer.ls1.data
        if hasattr(layer.mlp.fc2, "bias"):
            layer.mlp.fc2.bias.data *= layer.ls2.data
        layer.ls1[:] = 1
        layer.ls2[:] = 1


def _set_flag(model, args, flag, value):
    skip = getattr(args, "skip_names", None) or []
    for name, layer in find_qlayers(model, layers=[ActQuantWrapper]).items():
        if any(p in name for p in skip):
            continue
        setattr(layer.quantizer, flag, value)
    return model


def model_open_calibrate(model, args):
    return _set_flag(model, args, "calibrate", True)


def model_open_last_calibrate(model, args):
    return _set_flag(model, args, "last_calibrate", True)


def model_close_calibrate(model, args):
    return _set_flag(model, args, "calibrate", False)


def model_quant(model, args):
    return _set_flag(model, args, "quant", True)


def model_no_quant(model, args):
    return _set_flag(model, args, "quant", False)


class _SiblingGroup:
    """Frozen QuantLinears that are always fed the SAME tensor (q/k/v, gate/up): the first member
    called with an input quantises it once and runs ONE W4A8 GEMM over the concatenated output
    channels; the siblings called next with that very tensor get their column slice of the result
    (one int8 write shared by 3 resp. 2 GEMMs).  Results are bit-identical to
    separate calls: same codes, same scales, same per-channel epilogue.

    "The same tensor" is decided by object identity + version counter; the group keeps a reference
    to the input while results are pending, so the address cannot be recycled meanwhile."""

    def __init__(self, members):
        self.members = list(members)
        first = members[0]
        pks = [m._packed for m in members]
        self.offsets, off = [], 0
        for pk in pks:
            self.offsets.append(off)
            off += pk["N"]
        self.N = off
        self.Kp = pks[0]["Kp"]
        self.w4 = torch.cat([pk["w4"] for pk in pks])  # tile-major layout: n-blocks are contiguous
        self.s_w = torch.cat([pk["s_w"].reshape(-1) for pk in pks]).contiguous()
        if any(pk["bias"] is not None for pk in pks):
            self.bias = torch.cat([pk["bias"] if pk["bias"] is not None else
                                   torch.zeros(pk["N"], dtype=torch.float16, device=self.w4.device) for pk in pks])
            # adding an exact +0.0 bias leaves every fp32 sum unchanged (no -0.0 can occur: the
            # accumulator term is +-finite or +0.0)
        else:
            self.bias = None
        self.first = first
        self._x = None
        self._ver = None
        self._out = None
        self._served = set()

    def release(self):
        self._x = self._out = self._ver = None
        self._served = set()

    @staticmethod
    def _version_of(x):
        try:
            return x._version
        except RuntimeError:  # inference tensors do not track versions: never share across calls
            return None

    @torch.no_grad()
    def forward(self, member, x):
        idx = self.members.index(member)
        ver = self._version_of(x)
        hit = (ver is not None and self._x is x and self._ver == ver and idx not in self._served)
        if not hit and (ver is None or (self._x is not None and idx in self._served)):
            # sharing cannot be proven for this call (inference tensors carry no version counter;
            # a member called twice in a row is not the q/k/v pattern): run the member's OWN GEMM
            # instead of the concatenated one -- same result, a third (half) of the work
            return self._forward_single(member, x)
        if not hit:
            w = self.first
            qz = w.quantizer.quantizer
            x2 = x.reshape(-1, x.shape[-1])
            if x2.dtype == torch.bfloat16:
                x2 = x2.clamp(-65504.0, 65504.0).to(torch.float16)
            scales = qz.device_scales(x.device)
            modality = w.quantizer._modality(x) if qz.scale_visual is not None else None
            pre = _prequant_of(x, scales, self.Kp, modality)
            if pre is not None:
                q, srow = pre
            else:
                q, srow, _ = ops.act_quant_i8(x2, scales, modality, n_out=self.Kp, split=False,
                                              want_row_scales=modality is not None)
            s_a = srow if modality is not None else scales[0:1]
            self._out = ops.w4a8_gemm(q, self.w4, s_a, self.s_w, self.N, bias=self.bias)
            self._x, self._ver, self._served = x, ver, set()
        self._served.add(idx)
        N = member._packed["N"]
        y = self._out[:, self.offsets[idx]:self.offsets[idx] + N].reshape(*x.shape[:-1], N)
        if len(self._served) == len(self.members):
            self.release()
        return y


def _forward_single(self, member, x):
    pk = member._packed
    qz = member.quantizer.quantizer
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype == torch.bfloat16:
        x2 = x2.clamp(-65504.0, 65504.0).to(torch.float16)
    scales = qz.device_scales(x.device)
    modality = member.quantizer._modality(x) if qz.scale_visual is not None else None
    pre = _prequant_of(x, scales, pk["Kp"], modality)
    if pre is not None:
        q, srow = pre
    else:
        q, srow, _ = ops.act_quant_i8(x2, scales, modality, n_out=pk["Kp"], split=False,
                                      want_row_scales=modality is not None)
    s_a = srow if modality is not None else scales[0:1]
    out = ops.w4a8_gemm(q, pk["w4"], s_a, pk["s_w"], pk["N"], bias=pk["bias"])
    return out.reshape(*x.shape[:-1], pk["N"])


_SiblingGroup._forward_single = _forward_single


def fuse_siblings(wrappers):
    """Make frozen wrappers that read the same input share one activation quantisation and one
    GEMM launch (see _SiblingGroup).  Refuses (returns None, leaves the wrappers untouched) unless
    it is exact: identical static scales, no online rotation / split, same K, output widths that
    are multiples of the 256-channel weight tile (the last one may be ragged)."""
    wrappers = list(wrappers)
    if len(wrappers) < 2:
        return None
    f = wrappers[0]
    for i, w in enumerate(wrappers):
        if not (isinstance(w, ActQuantWrapper) and w.real_quant and w._group is None):
            return None
        if w.online_full_had or w.online_partial_had or w.split or w.out_quantizer.bits < 16:
            return None
        if w._packed["Kp"] != f._packed["Kp"] or w._packed["K"] != f._packed["K"]:
            return None
        if i < len(wrappers) - 1 and w._packed["N"] % 256 != 0:
            return None
        a, b = w.quantizer.quantizer, f.quantizer.quantizer
        if w.quantizer.msq != f.quantizer.msq or not torch.equal(a.scale.float().cpu(), b.scale.float().cpu()):
            return None
        if (a.scale_visual is None) != (b.scale_visual is None):
            return None
        if a.scale_visual is not None and not torch.equal(a.scale_visual.float().cpu(), b.scale_visual.float().cpu()):
            return None
    g = _SiblingGroup(wrappers)
    for w in wrappers:
        w._group = g
    return g


class _GatedPair:
    """The two input projections of a gated MLP as ONE W4A8 GEMM whose epilogue applies the gate
    (QWenMLP.forward, model/modeling_qwen_opt.py:382-386: ``a1 * F.silu(a2)``):
    ``value`` = the wrapper of a1 (Qwen-VL ``w1``, Llama-style ``up_proj``, InternLM2 ``w3``), ``gate`` =
    the wrapper whose output goes through silu (``w2`` / ``gate_proj`` / ``w1``).  The packed weight
    interleaves their rows in blocks of 128 (ops.interleave_gated); the [M, 2F] intermediate is never
    written for batches the 256 x 256 kernel serves.  Same fp16 rounding points as the reference."""

    def __init__(self, value, gate):
        pv, pg = value._packed, gate._packed
        self.value, self.gate = value, gate
        self.F, self.Kp = pv["N"], pv["Kp"]
        codes = ops.interleave_gated(ops.unpack_w4(pv["w4"], self.F, self.Kp), ops.unpack_w4(pg["w4"], self.F, self.Kp))
        self.w4 = ops.pack_w4(codes)
        self.s_w = ops.interleave_gated(pv["s_w"].reshape(-1), pg["s_w"].reshape(-1)).contiguous()
        self.bias = None
        if pv["bias"] is not None or pg["bias"] is not None:
            z = torch.zeros(self.F, dtype=torch.float16, device=self.w4.device)  # (+0.0 leaves every fp32 sum unchanged)
            self.bias = ops.interleave_gated(pv["bias"] if pv["bias"] is not None else z,
                                             pg["bias"] if pg["bias"] is not None else z).contiguous()

    @torch.no_grad()
    def forward(self, x):
        w = self.value
        qz = w.quantizer.quantizer
        x2 = x.reshape(-1, x.shape[-1])
        out_dtype = x2.dtype
        if x2.dtype == torch.bfloat16:
            x2 = x2.clamp(-65504.0, 65504.0).to(torch.float16)
        scales = qz.device_scales(x.device)
        modality = w.quantizer._modality(x) if qz.scale_visual is not None else None
        pre = _prequant_of(x, scales, self.Kp, modality)
        if pre is not None:
            q, srow = pre
        else:
            q, srow, _ = ops.act_quant_i8(x2, scales, modality, n_out=self.Kp, split=False,
                                          want_row_scales=modality is not None)
        s_a = srow if modality is not None else scales[0:1]
        y = ops.w4a8_gemm_gated(q, self.w4, s_a, self.s_w, self.F, bias=self.bias)
        if out_dtype != torch.float16:
            y = y.to(out_dtype)
        return y.reshape(*x.shape[:-1], self.F)

    __call__ = forward


def fuse_gated_pair(value, gate):
    """-> _GatedPair, or None (wrappers untouched) unless the fusion is exact: both frozen, fp16
    Linears of the same shape with F % 128 == 0, identical static activation scales, no online
    rotation / split / output quantiser."""
    f = value
    for w in (value, gate):
        if not (isinstance(w, ActQuantWrapper) and w.real_quant and w._packed is not None):
            return None
        if w.online_full_had or w.online_partial_had or w.split or w.out_quantizer.bits < 16:
            return None
        if w._packed["Kp"] != f._packed["Kp"] or w._packed["K"] != f._packed["K"] or w._packed["N"] != f._packed["N"]:
            return None
        if w.module.weight.dtype != torch.float16:  # (bf16 models round a1 / silu / the product to bf16)
            return None
        a, b = w.quantizer.quantizer, f.quantizer.quantizer
        if w.quantizer.msq != f.quantizer.msq or not torch.equal(a.scale.float().cpu(), b.scale.float().cpu()):
            return None
        if (a.scale_visual is None) != (b.scale_visual is None):
            return None
        if a.scale_visual is not None and not torch.equal(a.scale_visual.float().cpu(), b.scale_visual.float().cpu()):
            return None
    if f._packed["N"] % 128 != 0:
        return None
    return _GatedPair(value, gate)


_GATED_MLP_NAMES = (("w1", "w2", "c_proj"),             # Qwen-VL: c_proj(w1(x) * silu(w2(x)))  (modeling_qwen_opt.py:382-386)
                    ("up_proj", "gate_proj", "down_proj"),  # Qwen2-VL / MiniCPM-V (Llama style)
                    ("w3", "w1", "w2"))                  # InternLM2 (InternVL2): w2(silu(w1(x)) * w3(x))


def fuse_gated_mlp(mlp):
    """Route ``mlp.forward`` through one gated GEMM + the down projection.  Recognises the three
    naming schemes of the supported models; the activation must be SiLU.  Returns the _GatedPair or
    None when the module was left alone."""
    for value, gate, down in _GATED_MLP_NAMES:
        if all(isinstance(getattr(mlp, n, None), ActQuantWrapper) for n in (value, gate, down)):
            act = getattr(mlp, "act_fn", None)
            if act is not None and not isinstance(act, torch.nn.SiLU) and getattr(act, "__name__", "") != "silu":
                return None
            pair = fuse_gated_pair(getattr(mlp, value), getattr(mlp, gate))
            if pair is None:
                return None
            down_w = getattr(mlp, down)
            mlp._mq_gated = pair
            mlp.forward = lambda hidden_states, _p=pair, _d=down_w: _d(_p(hidden_states))
            return pair
    return None


@torch.no_grad()
def model_freeze(model, args=None, weight_quantizers=None):
    """New: switch every calibrated Linear wrapper to the real W4A8 path.  ``weight_quantizers`` is
    the dict the reference's gptq entry points return (name -> WeightQuantizer); names ending in
    ".module" / ".L2" are matched to their wrapper."""
    skip = getattr(args, "skip_names", None) or []
    scales = {}
    for k, wq in (weight_quantizers or {}).items():
        base = k.rsplit(".", 1)[0] if k.endswith((".module", ".L2")) else k
        scales[base] = wq.scale
    n = 0
    for name, layer in find_qlayers(model, layers=[ActQuantWrapper]).items():
        if any(p in name for p in skip) or not isinstance(layer.module, torch.nn.Linear):
            continue
        q = layer.quantizer
        if not (q.static and q.bits == 8 and q.quantizer.scale is not None):
            continue
        ws = None
        for key, s in scales.items():
            if key == name or key.endswith("." + name) or name.endswith("." + key):
                ws = s
                break
        if weight_quantizers and ws is None:
            warnings.warn(f"model_freeze: no weight-quantizer scale matches '{name}'; its weight is "
                          "re-quantised RTN and verified against the fp16 weight", stacklevel=2)
        layer.freeze(ws, check=bool(weight_quantizers))
        n += 1
    return n


@torch.no_grad()
def update_all_quantization_params(model, args=None):
    """Derive the scales of EVERY calibrated static wrapper of ``model`` at once: one device-to-host
    copy of all layer-wise MinmaxObserver statistics, the reference's scale expression
    (observer/minmax.py:30-46) evaluated on the host over all of them together, one copy back.
    Same values as calling ``ActQuantizer.update_quantization_params()`` wrapper by wrapper (which
    costs a device round trip per observer: 66 ms for 160 wrappers, profiles/r1e_multi_gpu_n2.log).
    Other observer kinds fall back to their own update.  Returns the number of wrappers updated."""
    from .observer.minmax import MinmaxObserver

    skip = getattr(args, "skip_names", None) or []
    fast, slow = [], []
    for name, layer in find_qlayers(model, layers=[ActQuantWrapper]).items():
        q = layer.quantizer
        if any(p in name for p in skip) or not getattr(q, "static", False):
            continue
        obs = [q.observer] + ([q.observer_visual] if q.msq else [])
        if all(type(o) is MinmaxObserver and not o._per_channel() for o in obs) and q.observer.symmetric:
            if any(o.state is not None for o in obs):
                fast.append(q)
        elif any(_observed(o) for o in obs):
            slow.append(q)
    for q in slow:
        q.update_quantization_params()
    if not fast:
        return len(slow)
    dev = next(o.state.device for q in fast for o in ([q.observer] + ([q.observer_visual] if q.msq else []))
               if o.state is not None)
    # three concatenations (not one per wrapper: 160 wrappers x a handful of tiny kernels is 10 ms), and no
    # per-wrapper tensor indexing on the host: everything below is a handful of vector operations plus
    # one loop that only stores views (merge of 160 wrappers: 6.3 -> ~2 ms)
    n = len(fast)
    zero2 = torch.zeros(2, dtype=torch.float32, device=dev)
    t_states = [q.observer.state if q.observer.state is not None else zero2 for q in fast]
    v_states = [q.observer_visual.state if (q.msq and q.observer_visual.state is not None) else zero2 for q in fast]
    counted = [q for q in fast if getattr(q, "_seen", None) is not None]
    parts = [torch.cat(t_states + v_states)]
    if counted:
        parts.append(torch.cat([q._seen for q in counted]).to(torch.float32))  # exact below 2^24 rows
    flat = torch.cat(parts).cpu()  # ONE copy of all statistics and row counts
    stats = flat[:4 * n].view(2, n, 2)  # [text | visual][wrapper][min, max]
    seen_rows = iter(flat[4 * n:].view(-1, 2).tolist())
    has_t, has_v = [], []
    for q in fast:
        if getattr(q, "_seen", None) is not None:
            nt, nv = next(seen_rows)
        else:  # (gather path / restored statistics: a state that exists has been seen)
            nt, nv = float(q.observer.state is not None), float(q.msq and q.observer_visual.state is not None)
        has_t.append(nt > 0)
        has_v.append(bool(q.msq) and nv > 0)
    # the reference's expression, in the dtype of the statistics, for all observers at once
    by_dtype = {}
    for i, q in enumerate(fast):
        dt = q.observer._dtype or (q.observer_visual._dtype if q.msq else None) or torch.float16
        by_dtype.setdefault(dt, []).append(i)
    bt = fast[0].observer.bit_type
    eps = torch.finfo(torch.float32).eps
    zp0 = torch.zeros((), dtype=torch.int64, device=dev)
    for dt, idx in by_dtype.items():
        sel = torch.tensor(idx)
        h = stats[:, sel, :]
        st = torch.max(torch.abs(h[0, :, 0].to(dt) / bt.lower_bound), torch.abs(h[0, :, 1].to(dt) / bt.upper_bound)).clamp_(eps)
        sv = torch.max(torch.abs(h[1, :, 0].to(dt) / bt.lower_bound), torch.abs(h[1, :, 1].to(dt) / bt.upper_bound)).clamp_(eps)
        ht = torch.tensor([has_t[i] for i in idx])
        hv = torch.tensor([has_v[i] for i in idx])
        st = torch.where(~ht & hv, sv, st)  # only visual tokens were ever seen
        sv = torch.where(hv, sv, st)
        # text/visual scales of every wrapper as one device tensor; each quantizer keeps views of it
        pairs = torch.stack([st, sv], 1).to(dev)  # [m, 2] in dt
        f32 = pairs.float().contiguous()
        s_text, s_vis, s_f32 = pairs[:, 0].unbind(0), pairs[:, 1].unbind(0), f32.unbind(0)
        for j, i in enumerate(idx):
            if not (has_t[i] or has_v[i]):
                continue
            # plain attributes of the quantizer module (never parameters / buffers): stored without
            # nn.Module.__setattr__'s registry checks
            d = fast[i].quantizer.__dict__
            d["scale"] = s_text[j]
            d["zero_point"] = zp0
            d["scale_visual"] = s_vis[j] if has_v[i] else None
            d["_dev_scales"] = s_f32[j]
    return n + len(slow)


# ---- calibration entry points the drivers call (reference :1079-1129) ------------------------------

def _calib_generate(model, args, dataset, calib_num, kwargs_attr, restore):
    from tqdm import tqdm

    lt = len(dataset.data)
    step = math.ceil(lt / calib_num)
    print("Calibrating...")
    model_open_calibrate(model.model, args)
    kw = getattr(model, kwargs_attr)
    saved = kw.get("max_new_tokens") if restore else None
    kw["max_new_tokens"] = 20
    for i in tqdm(range(0, lt, step)):
        if i + step >= lt:
            print("last calibrate")
            model_open_last_calibrate(model.model, args)
            kw["max_new_tokens"] = 1
        if hasattr(model, "use_custom_prompt") and model.use_custom_prompt(args.dataset_name):
            struct = model.build_prompt(dataset.data.iloc[i], dataset=args.dataset_name)
        else:
            struct = dataset.build_prompt(dataset.data.iloc[i])
        model.generate(message=struct, dataset=args.dataset_name)
    if restore:
        kw["max_new_tokens"] = saved
    else:
        setattr(model, kwargs_attr, {})
    model_close_calibrate(model.model, args)
    print("Calibrate End...")
    model_quant(model.model, args)


def calib_vqa_plus(model, args, dataset, calib_num):
    """Reference quant_utils.py:1079-1102: ~calib_num evenly spaced samples of a VLMEvalKit dataset
    through ``model.generate`` with the observers on (20 new tokens, 1 on the last sample, where the
    scales are derived), then quantisation on.  ``model`` is the VLMEvalKit wrapper (``.model`` holds
    the network, ``.kwargs`` the generation arguments)."""
    _calib_generate(model, args, dataset, calib_num, "kwargs", restore=False)


def calib_qwen2vl_plus(model, args, dataset, calib_num):
    """Reference quant_utils.py:1105-1129: the same protocol for the Qwen2-VL wrapper, whose
    generation arguments live in ``.generate_kwargs`` and are restored afterwards."""
    _calib_generate(model, args, dataset, calib_num, "generate_kwargs", restore=True)


# ---- the same protocol with caller-supplied batches (synthetic calibration, tests, bench) -----------

@torch.no_grad()
def calibrate_static(model, args, batches, forward=None):
    """N forward passes with observers on, parameters set on the last one, then quant on.
    ``batches``: iterable of model inputs; ``forward(model, batch)`` defaults to ``model(batch)``."""
    forward = forward or (lambda m, b: m(b))
    batches = list(batches)
    model_open_calibrate(model, args)
    for i, b in enumerate(batches):
        if i == len(batches) - 1:
            model_open_last_calibrate(model, args)
        forward(model, b)
    model_close_calibrate(model, args)
    model_quant(model, args)
    return model
