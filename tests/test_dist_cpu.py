"""world_size-2 gloo tests of the sharded-calibration merge (host logic; no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mquant_b200 import dist as mqd

    rng = np.random.default_rng(0)
    samples = [(rng.standard_normal((int(rng.integers(4, 32)), 64)) * 3).astype(np.float16) for _ in range(11)]
    mine = mqd.shard_indices(len(samples), rank, world)
    # local statistics as the observer kernel would leave them: {min, max} / {min[C], max[C]}
    states = []
    for per_channel in (False, True):
        ob = O.MinmaxObserver(8, "channel_wise" if per_channel else "layer_wise")
        for i in mine:
            ob.update(samples[i])
        mn = np.asarray(ob.min_val, dtype=np.float32).reshape(-1)
        mx = np.asarray(ob.max_val, dtype=np.float32).reshape(-1)
        states.append(torch.from_numpy(np.concatenate([mn, mx])))
    mqd.merge_observer_states(states)
    ret[rank] = [s.numpy().copy() for s in states]
    dist.destroy_process_group()


def test_sharded_minmax_merge_is_bit_identical_to_sequential():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    samples = [(rng.standard_normal((int(rng.integers(4, 32)), 64)) * 3).astype(np.float16) for _ in range(11)]
    for k, per_channel in enumerate((False, True)):
        ob = O.MinmaxObserver(8, "channel_wise" if per_channel else "layer_wise")
        for s in samples:
            ob.update(s)
        want = np.concatenate([np.asarray(ob.min_val, np.float32).reshape(-1),
                               np.asarray(ob.max_val, np.float32).reshape(-1)])
        for r in range(world):
            assert np.array_equal(ret[r][k], want)
        # and the scales every rank derives are the sequential ones
        half = want.size // 2
        s_seq, _ = ob.get_quantization_params()
        s_rank, _ = O.minmax_params(ret[0][k][:half].astype(np.float16), ret[0][k][half:].astype(np.float16))
        assert np.array_equal(np.asarray(s_seq).reshape(-1).view(np.uint16), np.asarray(s_rank).reshape(-1).view(np.uint16))


def test_shard_indices_cover_everything():
    from mquant_b200 import dist as mqd

    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in mqd.shard_indices(256, r, world))
        assert got == list(range(256))
    # weighted (token-balanced) shards: still a partition, equal counts, near-equal weight per rank
    rng = np.random.default_rng(0)
    w = rng.integers(256, 2049, 256).tolist()
    for world in (2, 4, 8):
        shards = [mqd.shard_indices(256, r, world, weights=w) for r in range(world)]
        assert sorted(i for sh in shards for i in sh) == list(range(256))
        assert {len(sh) for sh in shards} == {256 // world}
        tot = [sum(w[i] for i in sh) for sh in shards]
        assert max(tot) - min(tot) <= 2048, tot  # round-robin on this draw: up to ~6 % apart


def _worker_model(rank, world, port, ret):
    """merge_calibration on a wrapped model whose ranks saw different modalities: rank 0 only text
    tokens, rank 1 text and visual tokens (statistics installed as the kernels would leave them)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mquant_b200 import dist as mqd, quant_utils as qu

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(16, 8)

    net = Net()
    qu.add_actquant(net)
    q = net.fc.quantizer
    q.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=True)
    text = {0: (-1.5, 2.0), 1: (-3.0, 1.0)}[rank]
    q.observer.state = torch.tensor(text, dtype=torch.float32)
    q.observer._dtype, q.observer._channels = torch.float16, 16
    q._seen = torch.tensor([10, 0 if rank == 0 else 4], dtype=torch.int32)
    if rank == 1:
        q.observer_visual.state = torch.tensor([-8.0, 6.0], dtype=torch.float32)
        q.observer_visual._dtype, q.observer_visual._channels = torch.float16, 16
    n = mqd.merge_calibration(net)
    qz = q.quantizer
    ret[rank] = (n, float(qz.scale), None if qz.scale_visual is None else float(qz.scale_visual), q._seen.tolist())
    dist.destroy_process_group()


def test_merge_calibration_agrees_on_modalities_seen():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_model, args=(world, port, ret), nprocs=world, join=True)
    # merged statistics: text (-3, 2), visual (-8, 6); both ranks derive both scales
    s_text, _ = O.minmax_params(np.float16(-3.0), np.float16(2.0))
    s_vis, _ = O.minmax_params(np.float16(-8.0), np.float16(6.0))
    for r in range(world):
        n, st, sv, seen = ret[r]
        assert n == 2 and seen == [10, 4]  # MAX over ranks of the row counts
        assert np.float16(st) == s_text and sv is not None and np.float16(sv) == s_vis


def _worker_hessian(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mquant_b200 import dist as mqd, gptq as G

    rng = np.random.default_rng(3)
    samples = [rng.standard_normal((int(rng.integers(5, 40)), 48)).astype(np.float16) for _ in range(7)]
    layers = [torch.nn.Linear(48, 8, bias=False), torch.nn.Linear(48, 16, bias=False)]
    gs = [G.GPTQ(l) for l in layers]
    for i in mqd.shard_indices(len(samples), rank, world):
        for k, g in enumerate(gs):  # the second layer sees the samples scaled: two different Hessians in one collective
            H64, n = O.gptq_hessian(g.H.numpy(), g.nsamples, samples[i] * np.float16(1 + k), 1)
            g.H, g.nsamples = torch.from_numpy(H64.astype(np.float32)), n
    G.allreduce_hessians(gs)
    ret[rank] = [(g.H.numpy().copy(), g.nsamples) for g in gs]
    dist.destroy_process_group()


def test_sharded_gptq_hessians_merge_to_the_sequential_hessian():
    """calibration samples sharded over ranks, GPTQ.add_batch (gptq_utils.py:182-195) run
    locally, gptq.allreduce_hessians = ONE all-reduce(SUM) over every layer's n_r * H_r and the counts:
    every rank ends with the Hessian (and nsamples) of a sequential pass over all samples."""
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_hessian, args=(world, port, ret), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    samples = [rng.standard_normal((int(rng.integers(5, 40)), 48)).astype(np.float16) for _ in range(7)]
    for k in range(2):
        H, n = np.zeros((48, 48)), 0
        for s in samples:
            H, n = O.gptq_hessian(H, n, s * np.float16(1 + k), 1)
        for r in range(world):
            Hr, nr = ret[r][k]
            assert nr == n == len(samples)
            assert np.abs(Hr - H).max() <= 1e-6 * np.abs(H).max()
        assert np.array_equal(ret[0][k][0], ret[1][k][0])
