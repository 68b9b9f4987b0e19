"""MinmaxObserver on the GPU (mirror of fake_quant/observer/minmax.py:7-52).

The running statistics live in one fp32 device buffer that mq_observer_minmax merges into with
atomics ({min, max} or {min[C], max[C]}).  max/min of fp16 values are exact, so the buffer can
also be merged across ranks with allreduce(MAX) on {max, -min} (mquant_b200.dist) and stays
bit-identical to the reference's sequential update.
"""
import torch

from .. import ops
from .base import BaseObserver, layerwise_rows


class MinmaxObserver(BaseObserver):
    def __init__(self, module_type, bit_type, calibration_mode):
        super().__init__(module_type, bit_type, calibration_mode)
        self.symmetric = self.bit_type.signed
        self.state = None      # fp32 device buffer, see above
        self._dtype = None     # dtype of the observed activations (statistics keep it)
        self._channels = None

    # min_val / max_val (the reference's attributes) are VIEWS of the device buffer, materialised
    # when somebody reads them: an update is one kernel launch, not one launch plus four tiny
    # slice/convert kernels (calibration of ~1000-token samples is launch bound).
    @property
    def min_val(self):
        if self.state is None:
            return None
        return (self.state[:self._channels] if self._per_channel() else self.state[0]).to(self._dtype)

    @min_val.setter
    def min_val(self, v):  # BaseObserver.__init__ assigns None; callers may install statistics
        self._assign(v, 0)

    @property
    def max_val(self):
        if self.state is None:
            return None
        return (self.state[self._channels:] if self._per_channel() else self.state[1]).to(self._dtype)

    @max_val.setter
    def max_val(self, v):
        self._assign(v, 1)

    def _assign(self, v, which):
        """Assignment to min_val / max_val (as the reference's plain attributes allow, e.g. to
        restore saved statistics) writes through to the device buffer."""
        if v is None:
            return
        v = torch.as_tensor(v)
        C = v.numel()
        if self.state is None:
            n = 2 * C if self._per_channel() else 2
            self.state = torch.zeros(n, dtype=torch.float32, device=v.device)
            self._dtype, self._channels = v.dtype, C
        half = self.state.numel() // 2
        self.state[which * half:(which + 1) * half] = v.reshape(-1).to(self.state.device, torch.float32)

    def _per_channel(self):
        return self.calibration_mode != "layer_wise"

    def update(self, v):
        if self.module_type != "activation":
            raise NotImplementedError("mquant_b200 observes activations only")
        v = v.detach()
        if not self._per_channel():
            v = layerwise_rows(v, pad_ok=True)  # conv inputs: the scalar min/max ignore the layout
        elif v.dim() == 4:
            v = v.permute(0, 2, 3, 1)
        C = v.shape[-1]
        if self._per_channel() and C % 8 != 0:
            raise NotImplementedError(f"channel-wise MinmaxObserver: {C} channels (needs a multiple of 8); "
                                      "the drivers calibrate convolutions layer-wise")
        if self.state is None:
            n = 2 * C if self._per_channel() else 2
            self.state = torch.zeros(n, dtype=torch.float32, device=v.device)
            # (a caller that already converted bf16 activations has recorded their dtype)
            self._dtype, self._channels = (self._dtype or v.dtype), C
        if v.dtype == torch.bfloat16:
            # bf16 statistics: max/min commute with the (monotone, in-range exact) conversion;
            # _dtype keeps the input dtype so that the scale is derived in it, as the reference does
            v = v.clamp(-65504.0, 65504.0).to(torch.float16)
        ops.observer_minmax(v, self.state, per_channel=self._per_channel())

    def _refresh(self):
        """Kept for callers that modified .state directly (mquant_b200.dist): nothing to do, the
        views are derived on access."""

    def get_quantization_params(self, *args, **kwargs):
        return self._params_from_minmax()
