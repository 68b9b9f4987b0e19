"""Import the UNMODIFIED reference (``/root/reference``) on CPU -- generator-side only.

TEST INFRASTRUCTURE.  This module exists so that ``oracle/gen_golden.py`` and
``tools/gen_hadk_table.py`` can execute the reference's own Python in the build
container and freeze its outputs as fixtures.  ``/root/reference`` does not exist
on the GPU box, so nothing under ``tests/ -m gpu``, ``bench.py`` or
``__graft_entry__.smoke()`` may import this file.

Two third-party modules the reference imports are absent from this image and are
stubbed exactly as SURVEY.md section 8(c) describes:

* ``fast_hadamard_transform`` (Dao-AILab, unpinned git HEAD, docs/install.md:14-16):
  ``hadamard_transform(x, scale)`` = ``x @ H_n^T * scale`` along the last dim in
  Sylvester order, fp32 internally, result in ``x.dtype``.
* ``accelerate`` (only used by fake_quant/utils.py:413-433, never on the path).
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def _sylvester(n):
    import numpy as np
    h = np.array([[1.0]], dtype=np.float64)
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h


def import_reference():
    import torch  # noqa: F401  (must precede the stubs: transformers probes modules)
    import transformers  # noqa: F401

    if "fast_hadamard_transform" not in sys.modules:
        fht = types.ModuleType("fast_hadamard_transform")

        def hadamard_transform(x, scale=1.0):
            n = x.shape[-1]
            h = torch.from_numpy(_sylvester(n)).to(torch.float32)
            return ((x.float() @ h.T) * float(scale)).to(x.dtype)

        fht.hadamard_transform = hadamard_transform
        sys.modules["fast_hadamard_transform"] = fht
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.dispatch_model = None
        acc.infer_auto_device_map = None
        accu = types.ModuleType("accelerate.utils")
        accu.get_balanced_memory = None
        acc.utils = accu
        sys.modules["accelerate"] = acc
        sys.modules["accelerate.utils"] = accu
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from fake_quant import quant_utils, hadamard_utils  # noqa: E402
    return quant_utils, hadamard_utils
