"""torch-CPU port of the reference wrapper forward -- TEST / BASELINE INFRASTRUCTURE ONLY.

Same algorithm as oracle/mq_oracle.py:qlinear_fake (ActQuantWrapper.forward,
fake_quant/quant_utils.py:330-391), written with the torch CPU operators the reference itself
executes (x.float() / scale, round, clamp, * scale, .to(fp16), F.linear in fp16) so that the
``cpu_baseline`` / ``--impl reference`` timings of bench.py use the same multi-threaded MKL
kernels the reference would use on these host cores (numpy's fp16 conversions are single
threaded and would make the baseline unfairly slow).  Only tests/ and bench.py's CPU legs may
import this file.  Pinned by tests/test_oracle_golden.py::test_torch_port_matches_reference.
"""
import torch
import torch.nn.functional as F

from . import mq_oracle as O


def fwht_last(t):
    n = t.shape[-1]
    shape = t.shape
    t = t.reshape(-1, n)
    h = 1
    while h < n:
        t = t.reshape(-1, n // (2 * h), 2, h)
        t = torch.stack((t[:, :, 0, :] + t[:, :, 1, :], t[:, :, 0, :] - t[:, :, 1, :]), dim=2).reshape(-1, n)
        h *= 2
    return t.reshape(shape)


def matmul_hadU_cuda(x, hadK, K):
    """hadamard_utils.py:115-128 with fast_hadamard_transform restated (fp32 butterflies)."""
    n = x.shape[-1]
    scale = 1.0 / torch.tensor(n).sqrt()
    if K == 1:
        return (fwht_last(x.float()) * scale).to(x.dtype)
    t = (fwht_last(x.reshape(-1, K, n // K).float()) * scale).to(x.dtype)
    t = hadK.to(t.dtype) @ t
    return t.reshape(x.shape)


def fake_quant(x, scale, lo=-128, hi=127):
    """BaseQuantizer.forward (quantizer/base.py:44-50) + UniformQuantizer (uniform.py:20-43)."""
    xf = x.float()
    q = (xf / scale + 0).round().clamp(lo, hi)
    return (q * scale).to(x.dtype)


def qlinear_fake(x, Wfq, bias, scale, hadK=None, K=None):
    if K is not None:
        x = matmul_hadU_cuda(x, hadK, K)
    x = fake_quant(x, scale)
    return F.linear(x, Wfq, bias)


def hadK_tensor(n):
    h, k = O.get_hadK(n)
    return (None if h is None else torch.from_numpy(h)), k
