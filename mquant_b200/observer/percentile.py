"""PercentileObserver (mirror of fake_quant/observer/percentile.py:9-79).

The two order statistics come from mq_observer_percentile (exact counting select, same fp32
rank/lerp arithmetic as torch.quantile); the first-call clamp with 0 and the EMA (sigma) stay
on the host exactly as the reference writes them -- they are order dependent, so this observer
cannot be sharded across ranks bit-exactly (replicas only).
"""
import torch

from .. import ops
from .base import BaseObserver, layerwise_rows


class PercentileObserver(BaseObserver):
    def __init__(self, module_type, bit_type, calibration_mode, percentile_sigma=0.01,
                 percentile_alpha=0.99999):
        super().__init__(module_type, bit_type, calibration_mode)
        self.percentile_sigma = percentile_sigma
        self.percentile_alpha = percentile_alpha
        self.symmetric = self.bit_type.signed

    def update(self, v):
        assert self.calibration_mode == "layer_wise"
        v = layerwise_rows(v.detach(), pad_ok=False)  # the quantiles of the flattened tensor
        qs = ops.observer_percentile(v, self.percentile_alpha)
        cur_min, cur_max = qs[0], qs[1]
        if self.max_val is None:
            self.max_val = torch.max(cur_max, torch.zeros_like(cur_max))
        else:
            self.max_val = self.max_val + self.percentile_sigma * (cur_max - self.max_val)
        if self.min_val is None:
            self.min_val = torch.min(cur_min, torch.zeros_like(cur_min))
        else:
            self.min_val = self.min_val + self.percentile_sigma * (cur_min - self.min_val)

    def get_quantization_params(self, *args, **kwargs):
        return self._params_from_minmax()
