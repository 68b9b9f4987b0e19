"""GPTQ on the device kernels.

The reference's ``GPTQ`` (fake_quant/gptq/gptq_utils.py:171-310) spends its time in two loops;
both run in ``libmquant_b200.so`` here:

* ``add_batch``   -> ``mq_gptq_hessian_update``: H = H n/(n+t) + 2/(n+t) X^T X on tcgen05.  The
  activations are fp16 values, every product is exact in fp32, so this is the reference's fp32
  matmul up to summation order;
* ``fasterquant`` -> ``mq_gptq_block_f32`` for the column-by-column loop of each block (bit-exact
  against the eager loop on the same Cholesky factor).  Damping, the two Cholesky factorisations
  and the block-to-block update ``W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:]`` stay torch calls.

Supported configuration = what the drivers use for W4: symmetric, one scale per output channel,
no groups, no act-order.  ``accelerate_reference()`` installs the two methods into the
reference's class (the entry points ``from fake_quant.gptq.gptq_utils import GPTQ``) and keeps the
reference's own ``fasterquant`` for every other configuration; ``allreduce_hessians`` merges the
Hessians of calibration samples sharded over ranks with ONE all-reduce.
"""
from __future__ import annotations

import math

import torch

from . import ops

__all__ = ["GPTQ", "accelerate_reference", "allreduce_hessians"]


def add_batch(self, inp, out=None):
    """GPTQ.add_batch (gptq_utils.py:182-195): running mean of 2 X^T X over the samples seen."""
    if inp.dim() == 2:
        inp = inp.unsqueeze(0)
    batch = inp.shape[0]
    rows = inp.reshape(-1, inp.shape[-1])
    if rows.is_cuda and rows.dtype == torch.float16 and rows.shape[-1] % 8 == 0 and self.H.is_cuda:
        ops.gptq_hessian_update(rows, self.H, self.nsamples, batch)
        self.nsamples += batch
        return
    # dtypes the kernel does not take (bf16 / fp32 activations): the reference's expression
    self.H *= self.nsamples / (self.nsamples + batch)
    self.nsamples += batch
    x = math.sqrt(2 / self.nsamples) * rows.t().float()
    self.H += x.matmul(x.t())


def kernel_config_ok(self, blocksize, groupsize, actorder, static_groups):
    q = self.quantizer
    return bool(self.layer.weight.is_cuda and blocksize <= 128 and groupsize == -1 and not actorder and not static_groups
                and getattr(q, "sym", False) and getattr(q, "bits", 16) < 16)


def fasterquant(self, blocksize=128, percdamp=0.01, groupsize=-1, actorder=False, static_groups=False):
    """GPTQ.fasterquant (gptq_utils.py:197-303) for the supported configuration."""
    if not kernel_config_ok(self, blocksize, groupsize, actorder, static_groups):
        fallback = getattr(self, "_fasterquant_reference", None)
        if fallback is None:
            raise NotImplementedError("mquant_b200.gptq: symmetric per-channel GPTQ without groups / act-order only "
                                      "(other configurations run the reference's own loop: accelerate_reference())")
        return fallback(blocksize=blocksize, percdamp=percdamp, groupsize=groupsize, actorder=actorder,
                        static_groups=static_groups)
    layer_w = self.layer.weight.data
    W = layer_w.clone().float()
    if not self.quantizer.ready():
        self.quantizer.find_params(W)
    if self.quantizer.scale.numel() != W.shape[0]:
        raise NotImplementedError("mquant_b200.gptq: one scale per output channel expected")
    H = self.H
    del self.H
    dead = torch.diag(H) == 0
    H[dead, dead] = 1
    W[:, dead] = 0
    H.diagonal().add_(percdamp * torch.mean(torch.diag(H)))
    H = torch.cholesky_inverse(torch.linalg.cholesky(H))
    try:
        Hinv = torch.linalg.cholesky(H, upper=True).contiguous()  # (the upper factor comes back as a transposed view)
    except Exception:  # noqa: BLE001  (not positive definite: the reference falls back to RTN)
        self.layer.weight.data = self.quantizer.quantize(W).to(layer_w.dtype)
        return
    Q = torch.empty_like(W)
    self.Losses = torch.empty_like(W)
    for i1 in range(0, self.columns, blocksize):
        i2 = min(i1 + blocksize, self.columns)
        Q1, Err1, Losses1 = ops.gptq_block(W, i1, i2, Hinv, self.quantizer.scale, self.quantizer.maxq)
        Q[:, i1:i2] = Q1
        self.Losses[:, i1:i2] = Losses1 / 2
        if i2 < self.columns:
            W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])
    self.layer.weight.data = Q.reshape(layer_w.shape).to(layer_w.dtype)
    if torch.any(torch.isnan(self.layer.weight.data)):
        raise ValueError("NaN in weights")


class GPTQ:
    """Stand-alone class with the reference's attributes (layer, dev, rows, columns, H, nsamples, quantizer)."""

    def __init__(self, layer):
        self.layer = layer
        self.dev = layer.weight.device
        self.rows, self.columns = layer.weight.shape[0], layer.weight.shape[1]
        self.H = torch.zeros((self.columns, self.columns), device=self.dev)
        self.nsamples = 0

    add_batch = add_batch
    fasterquant = fasterquant

    def free(self):
        self.H = self.Losses = None


def accelerate_reference():
    """Install the kernel-backed methods into fake_quant.gptq.gptq_utils.GPTQ.  The reference's own
    fasterquant stays reachable for the configurations the kernel does not cover.  Returns True when
    the reference package was importable."""
    try:
        from fake_quant.gptq import gptq_utils as ref  # noqa: PLC0415
    except Exception:  # noqa: BLE001
        return False
    if getattr(ref.GPTQ, "_mq_accelerated", False):
        return True
    ref.GPTQ._fasterquant_reference = ref.GPTQ.fasterquant
    ref.GPTQ.add_batch = add_batch
    ref.GPTQ.fasterquant = fasterquant
    ref.GPTQ._mq_accelerated = True
    return True


@torch.no_grad()
def allreduce_hessians(gptqs, group=None):
    """Calibration samples sharded over ranks: rank r holds H_r = (2 / n_r) sum X^T X over its n_r
    samples; the Hessian of all samples is sum_r n_r H_r / sum_r n_r.  ONE all-reduce(SUM) over the
    concatenation of every layer's n_r H_r followed by the n_r."""
    import torch.distributed as dist

    gptqs = list(gptqs)
    if not gptqs or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return gptqs
    dev = gptqs[0].H.device
    flat = torch.cat([(g.H * float(g.nsamples)).reshape(-1) for g in gptqs] +
                     [torch.tensor([float(g.nsamples) for g in gptqs], device=dev)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    counts = flat[-len(gptqs):].tolist()
    off = 0
    for g, n in zip(gptqs, counts):
        m = g.H.numel()
        g.H.copy_((flat[off:off + m] / max(n, 1.0)).reshape(g.H.shape))
        g.nsamples = int(round(n))
        off += m
    return gptqs
