"""Observer base (mirror of fake_quant/observer/base.py:5-35)."""
import torch


def layerwise_rows(v, pad_ok):
    """A 2-D [rows, C] fp16-kernel-friendly view of ``v`` for LAYER-WISE statistics, which do not
    depend on the layout: activations [..., C] with C % 8 == 0 pass through; anything else (the
    patch-embedding convolutions see [B, 3, H, W] / [N, 3, 2, 14, 14] inputs, reference
    base.py:15-28 permutes them to channels-last with C = 3) is flattened and re-cut into rows of a
    power-of-two width.  ``pad_ok``: a tail that does not fill a row may be zero padded (exact for
    min/max, which fold 0 in anyway; not for quantiles)."""
    C = v.shape[-1]
    if v.dim() >= 2 and C % 8 == 0 and v.stride(-1) == 1:
        return v.reshape(-1, C)
    flat = v.reshape(-1)
    n = flat.numel()
    if n % 8 != 0:
        if not pad_ok:
            raise NotImplementedError(f"layer-wise observer: {n} elements is not a multiple of 8")
        flat = torch.nn.functional.pad(flat, (0, 8 - n % 8))
        n = flat.numel()
    width = 8
    while width < 4096 and n % (width * 2) == 0:
        width *= 2
    return flat.reshape(-1, width)


class BaseObserver:
    def __init__(self, module_type, bit_type, calibration_mode):
        self.module_type = module_type
        self.bit_type = bit_type
        self.calibration_mode = calibration_mode
        self.max_val = None
        self.min_val = None
        self.eps = torch.finfo(torch.float32).eps

    def update(self, v):
        raise NotImplementedError

    def get_quantization_params(self, *args, **kwargs):
        raise NotImplementedError

    # shared by minmax / percentile: reference minmax.py:30-52 == percentile.py:57-79
    def _params_from_minmax(self):
        # Evaluated on the HOST with the reference's own torch expressions: CUDA torch divides by a
        # scalar through a reciprocal multiply (1 ulp off in fp32), the CPU kernels divide.  The
        # statistics are 2 (or 2*C) numbers and this runs once per calibration.
        dev = self.max_val.device
        max_val, min_val = self.max_val.cpu(), self.min_val.cpu()
        qmax, qmin = self.bit_type.upper_bound, self.bit_type.lower_bound
        if self.symmetric:
            # evaluated in the dtype of the statistics (fp16 for fp16 activations)
            scale = torch.max(torch.abs(min_val / qmin), torch.abs(max_val / qmax))
            scale.clamp_(self.eps)
            zero_point = torch.zeros_like(max_val, dtype=torch.int64)
        else:
            scale = (max_val - min_val) / float(qmax - qmin)
            scale.clamp_(self.eps)
            zero_point = qmin - torch.round(min_val / scale)
            zero_point.clamp_(qmin, qmax)
        return scale.to(dev), zero_point.to(dev)
