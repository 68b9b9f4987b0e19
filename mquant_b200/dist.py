"""Multi-GPU plumbing for the two places the path shards.

* Calibration: samples are split across ranks; every rank runs the observers on its shard; the
  running min/max of all MinmaxObservers are merged with ONE allreduce(MAX) over a packed fp32
  buffer [max ; -min] (a few KB -- latency bound, NCCL over NVLink).  max/min are associative,
  commutative and exact, so the merged statistics (and the scales every rank then derives) are
  bit-identical to single-GPU sequential calibration.  Percentile/EMA observers are order
  dependent and are NOT merged (replicas only).
* Batched prefill: token batches are data parallel with replicated weights and no collective.

One process per GPU (torchrun); works on the gloo backend with CPU tensors for the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .observer.minmax import MinmaxObserver
from .quant_utils import ActQuantWrapper, find_qlayers


def shard_indices(n_items: int, rank: int, world: int, weights=None):
    """Shard of calibration samples: round-robin, or -- given per-sample weights such as
    token counts -- the samples sorted by weight and dealt in serpentine order (ranks 0..w-1, w-1..0, ...):
    equal counts and near-equal total weight per rank, so no rank waits for another at the merge.  Min/max
    statistics do not depend on which rank sees which sample."""
    if weights is None:
        return list(range(rank, n_items, world))
    assert len(weights) == n_items
    order = sorted(range(n_items), key=lambda i: (-weights[i], i))
    mine = []
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world)
        if (k if rnd % 2 == 0 else world - 1 - k) == rank:
            mine.append(i)
    return sorted(mine)


def _minmax_observers(model):
    obs = []
    for name, layer in sorted(find_qlayers(model, layers=[ActQuantWrapper]).items()):
        q = layer.quantizer
        if not getattr(q, "static", False):
            continue
        for o in (q.observer, getattr(q, "observer_visual", None)):
            if isinstance(o, MinmaxObserver):
                obs.append((name, o))
    return obs


def pack_minmax(states):
    """[{min..., max...}] fp32 buffers -> one flat [max ; -min] buffer (so a single MAX merges both)."""
    if states and all(st.numel() == 2 for st in states):
        # layer-wise observers (the drivers' configuration): three kernels for any number of them
        sign = torch.tensor([-1.0, 1.0], device=states[0].device)
        return (torch.stack(states) * sign).flip(1).reshape(-1)  # rows {max, -min}
    parts = []
    for st in states:
        half = st.numel() // 2
        parts.append(st[half:])
        parts.append(-st[:half])
    return torch.cat(parts) if parts else torch.zeros(0)


def unpack_minmax(flat, states):
    if states and all(st.numel() == 2 for st in states):
        sign = torch.tensor([-1.0, 1.0], device=flat.device)
        rows = flat.view(-1, 2).flip(1) * sign  # back to {min, max}
        torch._foreach_copy_(list(states), list(rows.unbind(0)))
        return
    off = 0
    for st in states:
        half = st.numel() // 2
        st[half:] = flat[off:off + half]
        off += half
        st[:half] = -flat[off:off + half]
        off += half


def merge_observer_states(states, group=None):
    """allreduce(MAX) of a list of MinmaxObserver state buffers, in place."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return states
    flat = pack_minmax(states).contiguous()
    if flat.numel():
        dist.all_reduce(flat, op=dist.ReduceOp.MAX, group=group)
        unpack_minmax(flat, states)
    return states


def merge_calibration(model, group=None, channels_hint=None):
    """Merge every static MinmaxObserver of ``model`` across ranks and derive the scales.

    ONE collective: the running {max, -min} of every observer and, for MSQ quantizers, the
    "modality seen" row counts travel in a single packed fp32 buffer through one
    ``all_reduce(MAX)`` (a few KB over NVLink: latency bound); the scales of all wrappers then come
    from ONE device-to-host copy (``quant_utils.update_all_quantization_params``).  Round 1 issued
    an extra 8-byte all_reduce per wrapper and a device round trip per observer: 66 ms for 160
    wrappers at N = 2, which capped strong scaling; this is < 1 ms.

    Observers that saw no data on a rank get a zero state first (layer-wise: two zeros, no
    communication needed; channel-wise: the channel count is agreed with one extra
    allreduce(MAX), only when such an observer exists).  Returns the number of observers merged."""
    from .quant_utils import update_all_quantization_params

    obs = _minmax_observers(model)
    if not obs:
        return 0
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    dev = None
    for _, o in obs:
        if o.state is not None:
            dev = o.state.device
            break
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    msq_quantizers = []
    for _, layer in sorted(find_qlayers(model, layers=[ActQuantWrapper]).items()):
        q = layer.quantizer
        if getattr(q, "static", False) and getattr(q, "msq", False):
            if getattr(q, "_seen", None) is None:
                # gather path / installed statistics: what this rank saw is known on the host
                q._seen = torch.tensor([int(q.observer.state is not None), int(q.observer_visual.state is not None)],
                                       dtype=torch.int32, device=dev)
            msq_quantizers.append(q)
    # missing states: layer-wise observers are two numbers; channel-wise ones need the channel count
    need_sizes = [o for _, o in obs if o.state is None and o._per_channel()]
    if distributed and (need_sizes or any(o._per_channel() for _, o in obs)):
        sizes = torch.tensor([0 if o.state is None else o.state.numel() for _, o in obs], dtype=torch.int64, device=dev)
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
        sizes = sizes.tolist()
    else:
        sizes = [2 if o.state is None and not o._per_channel() else (0 if o.state is None else o.state.numel())
                 for _, o in obs]
    live = []
    for (_, o), n in zip(obs, sizes):
        if n == 0:
            continue
        if o.state is None:
            o.state = torch.zeros(n, dtype=torch.float32, device=dev)
            o._dtype = o._dtype or torch.float16
            o._channels = n // 2
        live.append(o)
    if distributed:
        states = [o.state for o in live]
        seen = torch.cat([q._seen for q in msq_quantizers]).to(torch.float32) if msq_quantizers else \
            torch.zeros(0, device=dev)  # exact below 2^24 rows
        flat = torch.cat([pack_minmax(states), seen]).contiguous()
        dist.all_reduce(flat, op=dist.ReduceOp.MAX, group=group)  # the ONE data-path collective
        n_stats = sum(s.numel() for s in states)
        unpack_minmax(flat[:n_stats], states)
        if msq_quantizers:
            # (the MAX over ranks of the row counts: "seen anywhere" is all the scales depend on)
            merged = flat[n_stats:].to(torch.int32).view(-1, 2)
            torch._foreach_copy_([q._seen for q in msq_quantizers], list(merged.unbind(0)))
    for o in live:
        o._refresh()
    # the scales were last derived from rank-local statistics: recompute them from the merged ones
    update_all_quantization_params(model)
    return len(live)
