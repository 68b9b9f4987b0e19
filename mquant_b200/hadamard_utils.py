"""Hadamard utilities: API mirror of fake_quant/hadamard_utils.py:1-195.

* ``get_hadK`` reads the 11 non-power-of-two factors from ``data/hadk.bin`` (1 bit per entry,
  frozen from the reference's literals by tools/gen_hadk_table.py) instead of 97 k lines of
  floats.
* ``matmul_hadU_cuda`` is the ONLINE transform.  fp16 CUDA inputs go to the fused kernel
  (mq_hadamard_quant_i8 with only the fp16 output requested); fp32/fp64 inputs -- the offline
  weight rotation of ``apply_exact_had_to_linear`` -- run a device-agnostic torch butterfly in the
  input dtype (one-time model prep above the boundary, not the hot path).
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np
import torch

from . import ops

_HADK_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hadk.bin")
_PRIORITY = [172, 156, 140, 108, 60, 52, 36, 28, 40, 20, 12]  # get_hadK order (:28-75)
default_size = [172, 156, 140, 108, 60, 52, 36, 28, 20, 12, 1]
_table = None
_dev_cache = {}


def _load_table():
    global _table
    if _table is None:
        blob = open(_HADK_PATH, "rb").read()
        if blob[:8] != b"MQHADK1\0":
            raise RuntimeError("corrupt hadk.bin")
        (count,) = struct.unpack_from("<I", blob, 8)
        off, tab = 12, {}
        for _ in range(count):
            (k,) = struct.unpack_from("<I", blob, off)
            off += 4
            nb = (k * k + 7) // 8
            bits = np.unpackbits(np.frombuffer(blob, np.uint8, nb, off), bitorder="little")[: k * k]
            off += nb
            tab[k] = torch.from_numpy((bits.astype(np.float32) * 2 - 1).reshape(k, k))
        _table = tab
    return _table


def is_pow2(n):
    return (n & (n - 1) == 0) and (n > 0)


def get_hadK(n, transpose=False):
    """(hadK fp32 [K,K] | None, K): first factor (in the reference's priority order) dividing n;
    the cofactor must be a power of two."""
    for k in _PRIORITY:
        if n % k == 0:
            assert is_pow2(n // k)
            h = _load_table()[k]
            return (h.T.contiguous() if transpose else h.clone()), k
    assert is_pow2(n)
    return None, 1


def hadK_device(K, device):
    """fp16 +-1 copy of the K x K factor on `device` (cached), as the kernels want it."""
    key = (K, str(device))
    t = _dev_cache.get(key)
    if t is None:
        t = _load_table()[K].to(device=device, dtype=torch.float16).contiguous()
        _dev_cache[key] = t
    return t


def auto_pad_size(n):
    for size in default_size:
        if n % size == 0 and is_pow2(n // size):
            return n
    best = math.inf
    for size in default_size:
        p = 2 ** math.ceil(math.log2(math.ceil(n / size)))
        if p * size < best:
            best = p * size
    return best


def _fwht_last(t):
    """Unnormalised Sylvester WHT along the last dim, in t.dtype, any device."""
    n = t.shape[-1]
    shape = t.shape
    t = t.reshape(-1, n)
    h = 1
    while h < n:
        t = t.reshape(-1, n // (2 * h), 2, h)
        t = torch.stack((t[:, :, 0, :] + t[:, :, 1, :], t[:, :, 0, :] - t[:, :, 1, :]), dim=2).reshape(-1, n)
        h *= 2
    return t.reshape(shape)


def matmul_hadU(X, transpose=False):
    """Y = X (hadK (x) H_{n/K})^T / sqrt(n), pure torch in X.dtype (reference :79-100)."""
    n = X.shape[-1]
    hadK, K = get_hadK(n, transpose)
    t = _fwht_last(X.reshape(-1, K, n // K))
    if K > 1:
        t = hadK.view(1, K, K).to(t) @ t
    return t.reshape(X.shape) / torch.tensor(n).sqrt()


def matmul_hadUt(X):
    return matmul_hadU(X, transpose=True)


def random_hadamard_matrix(size, device):
    Q = torch.randint(low=0, high=2, size=(size,)).to(torch.float64)
    Q = torch.diag(Q * 2 - 1)
    return matmul_hadU(Q).to(device)


def matmul_hadU_cuda(X, hadK, K):
    """Online transform (reference :115-128)."""
    n = X.shape[-1]
    if X.is_cuda and X.dtype == torch.float16:
        hk = hadK_device(K, X.device) if K > 1 else None
        if K > 1 and hadK is not None and not torch.equal(hadK.to("cpu", torch.float32), _load_table()[K]):
            hk = hadK.to(device=X.device, dtype=torch.float16).contiguous()
        return ops.hadamard(X, had_k=K, hadk=hk, n=n)["y"].reshape(X.shape)
    if X.dtype == torch.float16:
        raise RuntimeError("matmul_hadU_cuda: fp16 activations must be CUDA tensors (no CPU fallback)")
    scale = 1.0 / torch.tensor(n).sqrt()
    if K == 1:
        return _fwht_last(X.contiguous()) * scale.to(X)
    t = _fwht_last(X.reshape(-1, K, n // K).contiguous()) * scale.to(X)
    t = hadK.to(t.device).to(t.dtype) @ t
    return t.reshape(X.shape)


def apply_exact_had_to_linear(module, had_dim=-1, output=False):
    """Offline rotation of a Linear's weight so that the online transform cancels
    (reference :135-191).  fp32 math on the weight's device."""
    assert isinstance(module, torch.nn.Linear)
    in_features, out_features = module.in_features, module.out_features
    if had_dim != -1:
        assert is_pow2(had_dim), "Hadamard dimension must be a power of 2!"
    W = module.weight.data
    dtype, dev = W.dtype, W.device
    Wf = W.float()
    Bf = module.bias.data.float() if module.bias is not None else None
    if had_dim == -1:
        if output:
            hk, K = get_hadK(out_features)
            Wf = matmul_hadU_cuda(Wf.t(), hk, K).t()
            if Bf is not None:
                module.bias.data = matmul_hadU_cuda(Bf.view(1, -1), hk, K).view(-1).to(device=dev, dtype=dtype)
        else:
            hk, K = get_hadK(in_features)
            Wf = matmul_hadU_cuda(Wf, hk, K)
    else:
        s = 1.0 / math.sqrt(had_dim)
        if output:
            Wt = Wf.t()
            shp = Wt.shape
            Wf = (_fwht_last(Wt.reshape(-1, shp[-1] // had_dim, had_dim).contiguous()) * s).reshape(shp).t()
            if Bf is not None:
                bs = Bf.shape
                Bf = (_fwht_last(Bf.reshape(-1, bs[-1] // had_dim, had_dim).contiguous()) * s).reshape(bs)
                module.bias.data = Bf.to(device=dev, dtype=dtype)
        else:
            n = Wf.shape[1]
            Wf = (_fwht_last(Wf.reshape(-1, n // had_dim, had_dim).contiguous()) * s).reshape(W.shape)
    module.weight.data = Wf.to(device=dev, dtype=dtype)
