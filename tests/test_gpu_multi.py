"""Multi-GPU tests (need >= 2 GPUs; skipped otherwise): sharded calibration with the NCCL
allreduce-max merge is bit-identical to sequential single-GPU calibration."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model(device):
    from mquant_b200 import quant_utils as qu

    torch.manual_seed(0)

    class Net(torch.nn.Module):  # add_actquant wraps Linear ATTRIBUTES / Sequential members of a parent
        def __init__(self):
            super().__init__()
            self.seq = torch.nn.Sequential(torch.nn.Linear(256, 512, bias=False), torch.nn.Linear(512, 256, bias=False))

        def forward(self, x):
            return self.seq(x)

    m = Net().half().to(device)
    qu.add_actquant(m)
    assert len(qu.find_qlayers(m, layers=[qu.ActQuantWrapper])) == 2
    for l in qu.find_qlayers(m, layers=[qu.ActQuantWrapper]).values():
        l.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax")
    return m


def _samples():
    g = torch.Generator().manual_seed(1)
    return [(torch.randn(int(torch.randint(16, 200, (1,), generator=g)), 256, generator=g) * (1 + i % 5)).half()
            for i in range(24)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    import torch.distributed as dist

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mquant_b200 import dist as mqd, quant_utils as qu

    dev = torch.device("cuda", rank)
    model = _make_model(dev)
    args = type("A", (), {"skip_names": []})()
    samples = _samples()
    qu.model_open_calibrate(model, args)
    with torch.no_grad():
        for i in mqd.shard_indices(len(samples), rank, world):
            model(samples[i].to(dev))
    n = mqd.merge_calibration(model)
    qu.model_close_calibrate(model, args)
    qu.model_quant(model, args)
    scales = [l.quantizer.quantizer.scale.cpu().numpy() for _, l in
              sorted(qu.find_qlayers(model, layers=[qu.ActQuantWrapper]).items())]
    out[rank] = (n, scales)
    dist.destroy_process_group()


def test_sharded_calibration_matches_sequential():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    from mquant_b200 import quant_utils as qu

    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # sequential reference on one GPU, the reference's protocol (params on the last sample)
    model = _make_model(torch.device("cuda", 0))
    args = type("A", (), {"skip_names": []})()
    with torch.no_grad():
        qu.calibrate_static(model, args, [s.cuda() for s in _samples()])
    want = [l.quantizer.quantizer.scale.cpu().numpy() for _, l in
            sorted(qu.find_qlayers(model, layers=[qu.ActQuantWrapper]).items())]
    for r in range(world):
        n, got = out[r]
        assert n == 2
        for a, b in zip(got, want):
            assert a.dtype == np.float16 and np.array_equal(a.view(np.uint16), b.view(np.uint16))
