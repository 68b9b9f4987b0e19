"""BASELINE.json configs[2]: W4A8 static calibration (minmax observers, 256 synthetic image+text
samples) of the InternVL2-8B LLM quantized-linear stack, sharded across the ranks of a torchrun
launch, merged with ONE NCCL allreduce(MAX) (mquant_b200/dist.py) -- bit-identical to sequential
calibration, which --verify checks on every rank.

    python tools/calib_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29511 tools/calib_bench.py --verify

A sample = S_i tokens (S_i in [256, 2048], seed 0) with one visual span; the timed region is the
hot path of calibration (online Hadamard with fp16 output on feed_forward.w2's input + observer
update of every wrapper), not the model's own fp16 GEMMs.  Samples come from a small pool of
resident tensors scaled per sample so that the statistics differ."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="internvl2-8b")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--verify", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mquant_b200 import dist as mqdist, quant_utils
    from mquant_b200.prefill import QuantLinearStack, synthetic_tokens

    args = type("A", (), {"skip_names": []})()
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(256, 2049, (a.samples,), generator=g).tolist()
    gains = (1.0 + torch.rand(a.samples, generator=g)).tolist()
    pool_len = 2048

    def build():
        return QuantLinearStack(a.model, n_layers=a.layers, device=dev, calibrate=False)

    stack = build()
    pool = stack.make_inputs(pool_len, dev, seed=7, visual_span=256, seq_len=pool_len)

    def sample(i):
        S = lengths[i]
        out = {k: (v[:S] * gains[i]).to(torch.float16) if k != "mask" else v[:S] for k, v in pool.items()}
        return out

    def run(stk, indices):
        quant_utils.model_open_calibrate(stk, args)
        for i in indices:
            stk.observe_tokens(sample(i))

    mine = mqdist.shard_indices(a.samples, rank, world)
    run(stack, mine[:2])  # warm-up (observers are reset below)
    stack = build()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    run(stack, mine)
    e1.record()
    n_merged = mqdist.merge_calibration(stack)
    e2.record()
    torch.cuda.synchronize()
    ms_obs, ms_merge = e0.elapsed_time(e1), e1.elapsed_time(e2)
    t = torch.tensor([ms_obs, ms_merge], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_obs, ms_merge = float(t[0]), float(t[1])
    identical = None
    if a.verify:
        ref = build()
        run(ref, range(a.samples))
        for _, layer in quant_utils.find_qlayers(ref, layers=[quant_utils.ActQuantWrapper]).items():
            layer.quantizer.update_quantization_params()
        identical = True
        la = quant_utils.find_qlayers(stack, layers=[quant_utils.ActQuantWrapper])
        lb = quant_utils.find_qlayers(ref, layers=[quant_utils.ActQuantWrapper])
        for name in la:
            qa, qb = la[name].quantizer.quantizer, lb[name].quantizer.quantizer
            for attr in ("scale", "scale_visual"):
                x, y = getattr(qa, attr), getattr(qb, attr)
                if (x is None) != (y is None) or (x is not None and not torch.equal(x.cpu(), y.cpu())):
                    identical = False
        flag = torch.tensor([1 if identical else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        identical = bool(flag.item())
    if rank == 0:
        tokens = sum(lengths)
        print(json.dumps({
            "metric": "w4a8_static_calibration_samples_per_s", "value": a.samples / ((ms_obs + ms_merge) * 1e-3),
            "unit": "samples/s", "n_gpus": world, "samples": a.samples, "tokens": tokens,
            "tokens_per_s": tokens / ((ms_obs + ms_merge) * 1e-3), "ms_observe": ms_obs, "ms_merge": ms_merge,
            "observers_merged": n_merged, "scaling": "strong", "identical_to_sequential": identical,
            "config": {"workload": stack.describe().replace("prefill", "calibration"), "layers": stack.n_layers,
                       "sample_tokens": "256..2048 (seed 0)"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
