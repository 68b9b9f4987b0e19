"""RMS normalisation in front of the quantized Linears.

``RMSN`` mirrors fake_quant/module_util.py:42-61 (the weight-free norm that
``rotation_utils.fuse_layer_norms`` leaves behind); for fp16 CUDA activations it runs
``mq_rmsn_quant_i8``.  ``RMSNQuant`` additionally emits the INT8 codes its consumers would
compute from the normalised tensor -- the frozen wrappers (``quant_utils.ActQuantWrapper`` /
sibling groups) that are fed its output pick them up instead of launching their own quantiser:
one read of the hidden state (2 B / element) and one write of the codes instead of
read + write + read + write.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["RMSN", "RMSNQuant", "fuse_rmsn_quant", "PreQuant"]


def _rmsn_torch(x, mean_dim, eps):
    """The reference expression (module_util.py:55-61), for dtypes / devices the kernel does not take."""
    input_dtype = x.dtype
    if x.dtype == torch.float16:
        x = x.to(torch.float32)
    variance = x.pow(2).sum(-1, keepdim=True) / mean_dim
    x = x * torch.rsqrt(variance + eps)
    return x.to(input_dtype)


class RMSN(torch.nn.Module):
    """Root-mean-square normalisation without weight (fake_quant/module_util.py:42-61)."""

    def __init__(self, mean_dim: int, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.mean_dim = mean_dim
        self.weight = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.is_cuda and x.dtype == torch.float16 and x.shape[-1] % 8 == 0:
            y = ops.rmsn_quant_i8(x, self.mean_dim, self.eps, want_codes=False, want_f16=True)["y"]
            return y.reshape(x.shape)
        return _rmsn_torch(x, self.mean_dim, self.eps)


class PreQuant:
    """INT8 codes that travel with a normalised activation tensor (attribute ``_mq_prequant``)."""

    __slots__ = ("q", "s_row", "scales", "modality", "n_out")

    def __init__(self, q, s_row, scales, modality, n_out):
        self.q, self.s_row, self.scales, self.modality, self.n_out = q, s_row, scales, modality, n_out

    def matches(self, scales, n_out, modality):
        """Valid for a consumer iff it would have quantised with the very same scale buffer, padding
        and modality mask."""
        return self.scales is scales and self.n_out == n_out and self.modality is modality


class RMSNQuant(RMSN):
    """RMSN whose output carries the consumers' INT8 codes.

    ``consumer`` is one of the frozen wrappers that read this norm's output (q/k/v resp. w1/w2 share
    their static scales, see quant_utils.fuse_siblings).  ``emit_y=False`` skips the fp16 store: the
    returned tensor is then UNINITIALISED and only valid as the argument of those consumers."""

    def __init__(self, mean_dim: int, eps=1e-5, consumer=None, emit_y=True):
        super().__init__(mean_dim, eps)
        object.__setattr__(self, "_consumer", consumer)  # (not a submodule: it lives elsewhere in the tree)
        self.emit_y = emit_y

    def _source(self):
        """The wrapper whose scale buffer the consumer(s) quantise with (a sibling group uses its first member's)."""
        w = self._consumer
        if w is not None and getattr(w, "_group", None) is not None:
            w = w._group.first
        return w

    def _fusable(self, x):
        w = self._source()
        if w is None or not (x.is_cuda and x.dtype == torch.float16 and x.shape[-1] % 8 == 0):
            return False
        q = w.quantizer
        return bool(w.real_quant and q.static and q.quant and not q.calibrate and not w.online_full_had
                    and not w.online_partial_had and not w.split and w.out_quantizer.bits >= 16
                    and w._packed is not None and w._packed["K"] == x.shape[-1])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self._fusable(x):
            return super().forward(x)
        w = self._source()
        qz = w.quantizer.quantizer
        scales = qz.device_scales(x.device)
        modality = w.quantizer._modality(x) if qz.scale_visual is not None else None
        n_out = w._packed["Kp"]
        r = ops.rmsn_quant_i8(x, self.mean_dim, self.eps, scales=scales, modality=modality, n_out=n_out,
                              want_codes=True, want_f16=self.emit_y, want_row_scales=modality is not None)
        y = r["y"].reshape(x.shape) if self.emit_y else torch.empty_like(x)
        y._mq_prequant = PreQuant(r["q"], r["s_row"], scales, modality, n_out)
        return y


def fuse_rmsn_quant(parent, name, consumer, emit_y=True):
    """Replace ``parent.<name>`` (an RMSN: attributes ``mean_dim`` and ``eps``) by an RMSNQuant feeding
    ``consumer``.  Returns the new module."""
    old = getattr(parent, name)
    new = RMSNQuant(old.mean_dim, old.eps, consumer=consumer, emit_y=emit_y)
    new.to(old.weight.device)
    setattr(parent, name, new)
    return new
