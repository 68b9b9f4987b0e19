"""build_quantizer (mirror of fake_quant/quantizer/build.py).  log2 is unused by the reference
drivers and out of scope."""
from .uniform import UniformQuantizer

str2quantizer = {"uniform": UniformQuantizer}


def build_quantizer(quantizer_str, bit_type, observer, module_type):
    if quantizer_str not in str2quantizer:
        raise NotImplementedError(f"quantizer '{quantizer_str}' is outside the mquant_b200 hot path")
    return str2quantizer[quantizer_str](bit_type, observer, module_type)
