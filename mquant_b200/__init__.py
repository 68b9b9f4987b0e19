"""mquant_b200 -- B200-native implementation of MQuant's quantized-linear hot path.

Python host code mirrors the reference's ``fake_quant`` API (quant_utils / hadamard_utils /
observer / quantizer / bit_type); all arithmetic runs in ``libmquant_b200.so`` (hand-written
sm_100a CUDA behind the C ABI of include/mquant_b200.h).  No CPU fallback.
"""
from . import bit_type, checkpoint, hadamard_utils, msq, ops, quant_utils, utils  # noqa: F401
from .checkpoint import load_packed, save_packed  # noqa: F401
from ._lib import LIB_PATH, MQuantLibraryError, lib  # noqa: F401
from .quant_utils import (ActQuantizer, ActQuantWrapper, QuantLinear, WeightQuantizer,  # noqa: F401
                          add_actquant, find_qlayers)

__version__ = "0.1.0"
