from .base import BaseQuantizer
from .build import build_quantizer
from .uniform import UniformQuantizer

__all__ = ["BaseQuantizer", "UniformQuantizer", "build_quantizer"]
