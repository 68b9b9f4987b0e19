"""AIFS: attention-invariant flexible switch -- visual tokens first.

The reference only sketches it in comments (model/modeling_qwen_opt.py:717-727, 778-792): per
sequence ``torch.cat((h[a+1:b], h[:a+1], h[b:]))`` for the image span ``(a, b)`` reported by
``img_pos`` (:604-606), with the rotary table permuted the same way.  This module states the
whole contract:

* ``plan(img_pos, seq_len)``: the permutation that moves every image span to the front of its
  sequence (spans keep their order, text keeps its order), its inverse, and the number of visual
  tokens per sequence;
* ``reorder / restore``: apply it to ``[B, S, ...]`` tensors (hidden states, rotary tables, labels);
* ``causal_mask``: the attention mask that keeps causal attention INVARIANT under the reorder:
  query i may look at key j iff ``orig_pos[j] <= orig_pos[i]`` -- the plain lower triangle is wrong
  once visual tokens sit in front of the text that preceded them;
* ``modality_mask``: after the reorder the MSQ mask is ``[1]*n_vis + [0]*n_text`` per sequence, so
  the per-token scale select of ``mq_act_quant_i8`` / ``mq_hadamard_quant_i8`` /
  ``mq_rmsn_quant_i8`` degenerates to two contiguous segments with one scalar scale each, and the
  W4A8 GEMM's per-row scale vector is piecewise constant.

Position-wise operators (the quantized Linears, norms, the MLP) commute with a token permutation;
attention commutes with it once the mask and the rotary phases travel with the tokens.  tests/
test_host_cpu.py checks both statements numerically.
"""
from __future__ import annotations

import torch

__all__ = ["plan", "reorder", "restore", "causal_mask", "modality_mask"]


def plan(img_pos, seq_len: int, batch: int, device=None):
    """img_pos: iterable of (sequence index, a, b): the image tokens of that sequence are
    positions a+1 .. b-1 (a / b are the image-start / image-end markers, modeling_qwen_opt.py:604-606).
    Returns dict(perm=[B,S] int64 with new[i] = old[perm[i]], inv=[B,S], n_visual=[B])."""
    vis = torch.zeros(batch, seq_len, dtype=torch.bool)
    for i, a, b in img_pos:
        i, a, b = int(i), int(a), int(b)
        if not (0 <= a < b <= seq_len):
            raise ValueError(f"aifs.plan: image span ({a}, {b}) outside a sequence of {seq_len} tokens")
        vis[i, a + 1:b] = True
    pos = torch.arange(seq_len).expand(batch, seq_len)
    # stable sort by "is text": visual tokens first, both groups in their original order
    key = (~vis).to(torch.int64) * seq_len + pos
    perm = key.argsort(dim=1)
    inv = perm.argsort(dim=1)
    out = {"perm": perm, "inv": inv, "n_visual": vis.sum(dim=1)}
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out


def _gather(x, idx):
    idx = idx.to(x.device)
    view = idx.reshape(idx.shape + (1,) * (x.dim() - 2)).expand(idx.shape + x.shape[2:])
    return x.gather(1, view)


def reorder(x, p):
    """[B, S, ...] -> visual tokens first."""
    return _gather(x, p["perm"])


def restore(x, p):
    """Inverse of ``reorder``."""
    return _gather(x, p["inv"])


def causal_mask(p):
    """[B, S, S] bool, True = query (row) may attend to key (column), for REORDERED tokens: the
    causal relation of the original positions."""
    orig = p["perm"]  # original position of the token now at index i
    return orig.unsqueeze(2) >= orig.unsqueeze(1)


def modality_mask(p, seq_len: int):
    """[B*S] uint8 MSQ mask of the reordered, flattened batch (1 = visual): two contiguous segments per sequence."""
    n = p["n_visual"].reshape(-1, 1)
    return (torch.arange(seq_len, device=n.device).unsqueeze(0) < n).to(torch.uint8).reshape(-1)
