"""UniformQuantizer on the GPU (mirror of fake_quant/quantizer/uniform.py:8-43 driven by
BaseQuantizer.forward, base.py:44-50).

``scale`` / ``zero_point`` keep the reference's types (scale in the activation dtype, int64 zero
point).  MSQ extension: ``scale_visual`` is a second per-tensor scale used for tokens whose
modality flag is set; with ``scale_visual is None`` the quantizer is exactly the reference's.
"""
import torch

from .. import ops
from .base import BaseQuantizer


class UniformQuantizer(BaseQuantizer):
    def __init__(self, bit_type, observer, module_type):
        super().__init__(bit_type, observer, module_type)
        self.scale = None
        self.zero_point = None
        self.scale_visual = None
        self._dev_scales = None  # fp32 [2] {text, visual} (layer_wise) or [C] (channel_wise)

    def update_quantization_params(self, *args, **kwargs):
        self.scale, self.zero_point = self.observer.get_quantization_params(*args, **kwargs)
        self._dev_scales = None

    def set_scales(self, scale, scale_visual=None):
        """Install calibrated scales directly (checkpoint load / MSQ calibration)."""
        self.scale = scale
        self.scale_visual = scale_visual
        self.zero_point = torch.zeros_like(scale, dtype=torch.int64)
        self._dev_scales = None

    @property
    def per_channel(self):
        return self.scale is not None and self.scale.dim() > 0 and self.scale.numel() > 1

    def device_scales(self, device):
        if self.scale is None:
            raise RuntimeError("UniformQuantizer used before calibration (scale is None)")
        if self._dev_scales is None or self._dev_scales.device != device:
            if self.per_channel:
                self._dev_scales = self.scale.detach().to(device=device, dtype=torch.float32).contiguous()
            else:
                self._dev_scales = ops.make_scales(self.scale, self.scale_visual, device)
        return self._dev_scales

    def _check_symmetric(self):
        if not self.bit_type.signed:
            raise NotImplementedError("unsigned activation types are unreachable from ActQuantizer.configure")

    def quant(self, inputs, scale=None, zero_point=None, modality=None):
        """INT8 codes (int8 tensor, same shape).  The reference returns them as fp32."""
        self._check_symmetric()
        if self.bit_type.bits != 8 or self.per_channel:
            raise NotImplementedError("integer codes are produced for per-tensor INT8 only")
        if inputs.dtype not in (torch.float16, torch.bfloat16):
            raise NotImplementedError("integer codes: fp16 / bf16 activations only")
        x = inputs if inputs.dtype == torch.float16 else inputs.clamp(-65504.0, 65504.0).to(torch.float16)
        q, _, _ = ops.act_quant_i8(x, self.device_scales(x.device), modality, want_row_scales=False)
        return q.reshape(inputs.shape)

    def dequantize(self, inputs, scale=None, zero_point=None):
        s = self.scale if scale is None else scale
        return (inputs.float() * s.float())

    def forward(self, inputs, modality=None):
        self._check_symmetric()
        if inputs.dtype == torch.bfloat16:
            return self._forward_bf16(inputs, modality)
        if inputs.dtype != torch.float16:
            raise NotImplementedError("mquant_b200 quantises fp16 / bf16 activations (fp32 does not convert exactly)")
        return ops.act_fake_quant(inputs, self.device_scales(inputs.device), modality,
                                  bits=self.bit_type.bits, per_channel=self.per_channel)

    def _forward_bf16(self, inputs, modality):
        """bf16 activations (InternVL2 and Qwen2-VL run in bf16).  The kernels take fp16; a
        bf16 value converts to fp16 EXACTLY while 2^-14 <= |x| <= 65504, so the INT8 codes are the
        reference's; the dequantisation ``(q * scale).to(x.dtype)`` is then done with the
        reference's own expression (uniform.py:34-43) to get its rounding into the input dtype.
        Values beyond +-65504 saturate (they have code +-127/-128 for every sane scale), values
        below 2^-14 lose bits that cannot reach the code unless scale < 2^-13."""
        if self.bit_type.bits != 8 or self.per_channel:
            raise NotImplementedError("bf16 activations: per-tensor INT8 only (cast the model to fp16 otherwise)")
        lim = 65504.0
        xh = inputs.clamp(-lim, lim).to(torch.float16)
        x2 = xh.reshape(-1, xh.shape[-1])
        q, srow, _ = ops.act_quant_i8(x2, self.device_scales(inputs.device), modality,
                                      want_row_scales=modality is not None and self.scale_visual is not None)
        if srow is not None:
            y = q.float() * srow[:, None]
        else:
            y = q.float() * self.scale.to(inputs.device).float()
        # torch.round keeps the sign of a tiny negative quotient: a zero code of a strictly negative
        # input dequantises to -0.0 (x = -0.0 itself becomes +0.0 through "+ zero_point")
        y = torch.where((q == 0) & (x2 < 0), torch.full_like(y, -0.0), y)
        return y.to(inputs.dtype).reshape(inputs.shape)
