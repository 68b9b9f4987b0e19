"""BaseQuantizer (mirror of fake_quant/quantizer/base.py:6-50)."""
import torch.nn as nn


class BaseQuantizer(nn.Module):
    def __init__(self, bit_type, observer, module_type):
        super().__init__()
        self.bit_type = bit_type
        self.observer = observer
        self.module_type = module_type

    def update_quantization_params(self, *args, **kwargs):
        pass

    def quant(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def dequantize(self, inputs, scale=None, zero_point=None):
        raise NotImplementedError

    def forward(self, inputs):
        raise NotImplementedError
