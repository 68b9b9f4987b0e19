"""Modality-Specific Static quantisation (MSQ) plumbing.

The reference's ``ActQuantWrapper.forward(x)`` has no mask argument (and the released code has
one scale per Linear), so the per-token visual/text flag travels out of band:
the model root sets it once per forward (image spans are known where the vision tokens are
spliced in, model/modeling_qwen_opt.py:715-716) and every QuantLinear whose quantizer has a
visual scale reads it.  Default (no context) = all text, which is also the decode case (S = 1).
"""
from __future__ import annotations

import contextlib
import threading

import torch

_state = threading.local()


def current_modality(num_tokens=None, device=None):
    m = getattr(_state, "mask", None)
    if m is None:
        return None
    if num_tokens is not None and m.numel() != num_tokens:
        return None  # a tensor with a different token count (e.g. visual tower): all "text"
    if device is not None and m.device != device:
        m = m.to(device)
    return m


@contextlib.contextmanager
def modality_context(mask):
    """mask: bool/uint8 tensor, one entry per token of the flattened [B*S] batch; nonzero = visual."""
    prev = getattr(_state, "mask", None)
    _state.mask = None if mask is None else mask.reshape(-1).to(torch.uint8).contiguous()
    try:
        yield
    finally:
        _state.mask = prev
