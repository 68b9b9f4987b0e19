"""``hadamard_transform(x, scale)``: the call signature of fast_hadamard_transform (Dao-AILab; the
reference's native seam, see compat/fast_hadamard_transform) on the B200 library.

    y[..., :] = (x[..., :] @ H_n) * scale        H_n Sylvester, fp32 inside, result in x.dtype

* fp16 CUDA tensors whose last dim is a power of two in [64, 8192] -- the online activation
  rotations (hadamard_utils.py:118,124; quant_utils.py:351) -- run ``mq_hadamard_transform_f16``.
* every other dtype / size / device is OFFLINE model preparation in the reference (weight rotation,
  hadamard_utils.py:168-188, rotation_utils.py) or a small per-head transform: a device-agnostic
  torch butterfly in fp32 (fp64 for fp64 inputs), zero-padded to the next power of two like the
  original.  fp16 CPU tensors raise: the activation path has no CPU fallback.
Inference only (the reference never differentiates through it).
"""
from __future__ import annotations

import torch

from . import ops
from .hadamard_utils import _fwht_last


def _next_pow2(n):
    p = 1
    while p < n:
        p *= 2
    return p


def hadamard_transform(x, scale=1.0):
    n = x.shape[-1]
    if x.dtype == torch.float16 and not x.is_cuda:
        raise ops._lib.MQuantLibraryError("hadamard_transform: fp16 activations must be CUDA tensors (no CPU fallback)")
    if x.is_cuda and x.dtype == torch.float16 and 64 <= n <= 8192 and n & (n - 1) == 0:
        return ops.hadamard_transform(x, float(scale))
    npad = _next_pow2(n)
    work = torch.float64 if x.dtype == torch.float64 else torch.float32
    t = x.to(work)
    if npad != n:
        t = torch.nn.functional.pad(t, (0, npad - n))
    y = _fwht_last(t.contiguous()) * scale
    if npad != n:
        y = y[..., :n]
    return y.to(x.dtype)
