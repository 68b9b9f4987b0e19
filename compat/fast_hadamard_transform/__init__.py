"""Drop-in for the third-party ``fast_hadamard_transform`` CUDA extension -- the ONE native seam the
reference sits on (imported at fake_quant/hadamard_utils.py:2, quant_utils.py:7,
rotation_utils.py:11; called at hadamard_utils.py:118,124,168,178,188 and quant_utils.py:351).

With ``<repo>/compat`` first on PYTHONPATH, ``import fast_hadamard_transform`` resolves here and
``hadamard_transform(x, scale)`` runs libmquant_b200.so (C ABI ``mq_hadamard_transform_f16``) for
fp16 CUDA activations.  The implementation lives in ``mquant_b200.fht`` so that the product
package owns it; this file only gives it the original module name.
"""
from mquant_b200.fht import hadamard_transform  # noqa: F401

__all__ = ["hadamard_transform"]
