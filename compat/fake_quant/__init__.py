"""Drop-in shim: put ``<repo>/compat`` FIRST on PYTHONPATH and the reference drivers'
``from fake_quant import quant_utils, hadamard_utils, utils`` (exam/quant_qwenvl.py:5-10) resolve
to the B200-native modules.  Sub-packages the hot path does not own (gptq, rotation_utils, ...)
are looked up in the reference checkout named by $MQUANT_REFERENCE_ROOT (its fake_quant/ dir is
appended to this package's search path), so they keep running unchanged on top of our
WeightQuantizer / ActQuantWrapper.
"""
import os
import sys

import mquant_b200
from mquant_b200 import bit_type, hadamard_utils, observer, quant_utils, quantizer, utils  # noqa: F401

for _name in ("bit_type", "hadamard_utils", "observer", "quant_utils", "quantizer", "utils"):
    sys.modules[__name__ + "." + _name] = getattr(mquant_b200, _name, None) or sys.modules["mquant_b200." + _name]
sys.modules[__name__ + ".observer.minmax"] = sys.modules["mquant_b200.observer.minmax"]
sys.modules[__name__ + ".observer.percentile"] = sys.modules["mquant_b200.observer.percentile"]
sys.modules[__name__ + ".quantizer.uniform"] = sys.modules["mquant_b200.quantizer.uniform"]

_ref = os.environ.get("MQUANT_REFERENCE_ROOT")
if _ref and os.path.isdir(os.path.join(_ref, "fake_quant")):
    __path__.append(os.path.join(_ref, "fake_quant"))
