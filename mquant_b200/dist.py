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


def shard_indices(n_items: int, rank: int, world: int):
    """Round-robin shard of calibration samples."""
    return list(range(rank, n_items, world))


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
    parts = []
    for st in states:
        half = st.numel() // 2
        parts.append(st[half:])
        parts.append(-st[:half])
    return torch.cat(parts) if parts else torch.zeros(0)


def unpack_minmax(flat, states):
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
    """Merge every static MinmaxObserver of `model` across ranks and refresh its min/max views.
    Observers that saw no data on a rank get a zero state of the right size first (needs one rank
    to have seen data: sizes are agreed via an allreduce(MAX) of the element counts)."""
    obs = _minmax_observers(model)
    if not obs:
        return 0
    # which modality each MSQ quantizer has seen, as known on the host (gather path), BEFORE
    # missing observer states are zero-filled below
    for _, layer in find_qlayers(model, layers=[ActQuantWrapper]).items():
        q = layer.quantizer
        if getattr(q, "static", False) and getattr(q, "msq", False) and getattr(q, "_seen", None) is None:
            def seen(o):
                if o is None:
                    return 0
                return int((o.state if hasattr(o, "state") else o.max_val) is not None)
            q._seen_host = (seen(q.observer), seen(q.observer_visual))
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    dev = None
    for _, o in obs:
        if o.state is not None:
            dev = o.state.device
            break
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    sizes = torch.tensor([0 if o.state is None else o.state.numel() for _, o in obs], dtype=torch.int64, device=dev)
    if distributed:
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX, group=group)
    live = []
    for (name, o), n in zip(obs, sizes.tolist()):
        if n == 0:
            continue
        if o.state is None:
            o.state = torch.zeros(n, dtype=torch.float32, device=dev)
            o._dtype = o._dtype or torch.float16
            o._channels = n // 2
        live.append(o)
    merge_observer_states([o.state for o in live], group)
    for o in live:
        o._refresh()
    # row counts of the fused text/visual observer (which modality has been seen anywhere)
    if distributed:
        for _, layer in sorted(find_qlayers(model, layers=[ActQuantWrapper]).items()):
            q = layer.quantizer
            if getattr(q, "static", False) and getattr(q, "msq", False) and q.observer.state is not None:
                if getattr(q, "_seen", None) is None:
                    # a rank that used the gather path knows on the host what it saw
                    q._seen = torch.tensor(list(getattr(q, "_seen_host", (0, 0))), dtype=torch.int32,
                                           device=q.observer.state.device)
                dist.all_reduce(q._seen, op=dist.ReduceOp.SUM, group=group)
    # the scales were last derived from rank-local statistics: recompute them from the merged ones
    for _, layer in find_qlayers(model, layers=[ActQuantWrapper]).items():
        q = layer.quantizer
        if getattr(q, "static", False) and (getattr(q.observer, "state", None) is not None or
                                            (q.msq and getattr(q.observer_visual, "state", None) is not None)):
            q.update_quantization_params()
    return len(live)
