#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the MQuant quantized-linear hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm: the oracle port)

A "step" is one pass of the hot path over one batch of synthetic tokens: the 7 x 32 QuantLinear
forwards of the Qwen-VL-Chat 9.6B LLM prefill (BASELINE.json configs[1]) -- per Linear one fused
[online Hadamard +] MSQ static-INT8 quantisation kernel and one tcgen05 W4A8 GEMM.  Attention,
norms and activations are outside the path and are not run.

value      tokens/s of the whole job, inputs resident in HBM, CUDA-event timed, max over ranks
e2e        same through the public API with HOST buffers: per step the token batch's hidden states
           are copied from pinned host memory and the last-token outputs are read back
roofline   dominant kernel (w4a8_gemm_v4_kernel): ops of all GEMM launches of a step / their summed
           CUDA-event durations, against 2 x the measured bf16 tensor peak (INT8 runs at twice the
           bf16 rate on B200; MEASURED_PEAKS.json has no INT8 entry)
cpu_baseline  the oracle port of the reference's fake-quant layer on the host cores, bounded sample
Multi-GPU: token batches are data parallel, weights replicated, no collective (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "w4a8_prefill_tokens_per_s"
UNIT = "tokens/s"
WORKLOAD = "qwenvl-9.6b LLM prefill, quantized-linear path (7 QuantLinear x 32 layers, W4A8, MSQ, online Hadamard K=172)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tokens", type=int, default=16384, help="tokens per GPU per step (16 x 1024)")
    ap.add_argument("--seq-len", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=None, help="default: the model's layer count")
    ap.add_argument("--workload", default="qwenvl-9.6b", choices=["qwenvl-9.6b", "qwen2vl-7b", "internvl2-8b"],
                    help="qwenvl-9.6b = BASELINE configs[1] (the bench line); the others are the LLM shapes "
                         "of configs[3]/[2], with --dynamic-visual for Qwen2-VL's dynamic-resolution spans")
    ap.add_argument("--dynamic-visual", action="store_true")
    ap.add_argument("--no-fuse", action="store_true",
                    help="do not share the activation quantisation / GEMM launch between the Linears of a "
                         "block that read the same tensor (q/k/v, the two up projections)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's fake-quant wrapper (oracle/mq_oracle.py:qlinear_fake)
# ------------------------------------------------------------------------------------------------
def cpu_layer_sample(tokens=1024, reps=1):
    """Time ONE decoder layer's 7 wrapped Linears (fake-quant fp16 semantics, the torch-CPU
    operators the reference executes) on `tokens` tokens with all host threads.  Returns seconds
    per layer (best of reps)."""
    import torch

    from oracle import torch_port as TP

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    hidden, ffn = 4096, 11008
    spec = [(hidden, hidden, True, False)] * 3 + [(hidden, hidden, False, False)] + \
           [(hidden, ffn, False, False)] * 2 + [(ffn, hidden, False, True)]
    layers = []
    for k, n, bias, had in spec:
        W = (torch.randn(n, k, generator=g) * 0.02)
        s_w = W.abs().amax(1, keepdim=True).clamp(min=1e-5) / 7
        Wfq = (s_w * (W / s_w).round().clamp(-8, 7)).half()
        b = (torch.randn(n, generator=g) * 0.1).half() if bias else None
        x = (torch.randn(tokens, k, generator=g) * 3).half()
        hk, K = TP.hadK_tensor(k) if had else (None, None)
        layers.append((x, Wfq, b, torch.tensor(0.12, dtype=torch.float16), hk, K))
    best = float("inf")
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            for x, Wfq, b, s, hk, K in layers:
                TP.qlinear_fake(x, Wfq, b, s, hk, K)
            best = min(best, time.perf_counter() - t0)
    return best


def run_reference_arm(a):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the reference is
    pure Python/PyTorch and /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tokens = 1024
    a.layers = a.layers or 32
    for _ in range(min(a.warmup, 1)):
        cpu_layer_sample(tokens)
    ts = [cpu_layer_sample(tokens) for _ in range(max(1, min(a.steps, 3)))]
    per_layer = sum(ts) / len(ts)
    step_s = per_layer * a.layers  # a full 32-layer pass over `tokens` tokens
    value = tokens / step_s
    sample = f"1 decoder layer (7 fake-quant Linears, online Hadamard K=172) on {tokens} tokens, x{a.layers} layers"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": len(ts), "warmup": min(a.warmup, 1), "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8xint4->int32 (reference: fp16 fake-quant)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "tokens_per_step": tokens, "layers": a.layers},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in out.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (mquant_b200 has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()

    from mquant_b200 import ops
    from mquant_b200.prefill import QuantLinearStack

    dev = torch.device("cuda", local_rank)
    stack = QuantLinearStack(a.workload, n_layers=a.layers, device=dev, fuse_siblings=not a.no_fuse)
    a.layers = stack.n_layers
    inputs = stack.make_inputs(a.tokens, dev, seed=100 + rank, seq_len=a.seq_len,
                               visual_span="dynamic" if a.dynamic_visual else 256)
    n_seq = max(1, a.tokens // a.seq_len)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- value: inputs resident in HBM.  The activation tensors (134 MB / 360 MB each) exceed
    # nothing-cached assumptions: four distinct inputs + 3.2 GB of weights stream per step, far
    # larger than the 126 MB L2, so no explicit flush is needed between steps.
    with ClockSampler(local_rank) as clk:
        ms_step = timed(lambda: stack.forward_tokens(inputs), a.steps, a.warmup)
    clocks = clk.summary()
    tokens_total = a.tokens * world
    value = tokens_total / (ms_step * 1e-3)

    # ---- TTFT of ONE request (seq_len tokens alone on the GPU): the latency end of the metric;
    # ttft_ms_per_request below is the amortised figure of the batched step
    one = {k: (v[:a.seq_len] if hasattr(v, "shape") else v) for k, v in inputs.items()}
    ms_single = timed(lambda: stack.forward_tokens(one), 5, 3)

    # ---- e2e: host buffers -> device -> path -> host, every step
    h_host = torch.empty(inputs["h"].shape, dtype=torch.float16).pin_memory()
    h_host.copy_(inputs["h"].cpu())
    out_host = torch.empty((n_seq, stack.hidden), dtype=torch.float16).pin_memory()
    last_idx = torch.arange(a.seq_len - 1, a.tokens, a.seq_len, device=dev)[:n_seq]
    e2e_inputs = dict(inputs)

    def e2e_step():
        e2e_inputs["h"] = h_host.to(dev, non_blocking=True)
        out = stack.forward_tokens(e2e_inputs)
        out_host.copy_(out.reshape(-1, stack.hidden)[last_idx], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e = timed(e2e_step, a.steps, min(a.warmup, 3))
    e2e_value = tokens_total / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: time every GEMM launch of one step with CUDA events
    gemm_ms, gemm_launches = [], [0]
    orig = ops.w4a8_gemm

    def timed_gemm(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*args, **kw)
        e1.record()
        gemm_ms.append((e0, e1))
        gemm_launches[0] += 1
        return r

    import mquant_b200.quant_utils as _qu
    _qu.ops.w4a8_gemm = timed_gemm
    try:
        stack.forward_tokens(inputs)
        torch.cuda.synchronize()
        gemm_total_ms = sum(e0.elapsed_time(e1) for e0, e1 in gemm_ms)
    finally:
        _qu.ops.w4a8_gemm = orig
    n_gemm = gemm_launches[0]
    gemm_ops = stack.gemm_ops_per_token() * a.tokens
    achieved = gemm_ops / (gemm_total_ms * 1e-3) / 1e12
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        # No INT8 entry in MEASURED_PEAKS.json: INT8 issues at twice the bf16 rate.  The in-step GEMM
        # EXCEEDS 2 x the *sustained* bf16 figure (that run was power limited at 1290 MHz; the INT8
        # kernel holds ~1800 MHz), so the valid upper bound is 2 x the burst figure; the
        # sustained-derived fraction is reported next to it.
        peak = 2.0 * float(pk.get("bf16_tflops", 1590.0))
        peak_sus = 2.0 * float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0)))
        peak_src = "2 x measured bf16_tflops burst (MEASURED_PEAKS.json; INT8 = 2x bf16 rate, no INT8 entry)"
    else:
        peak, peak_sus, peak_src = 2.0 * 1590.0, 2.0 * 1400.0, "2 x fallback bf16 (B200_PROFILING.md)"
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (bytes per
    # launch, averaged over the step's shape mix); only meaningful for the default workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1e_gemm_traffic.json")
    if os.path.exists(tpath) and a.tokens == 16384:
        traffic = json.load(open(tpath)).get("avg_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "w4a8_gemm_v4_kernel", "achieved": achieved, "peak": peak,
                "unit": "TOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "peak_2x_bf16_sustained": peak_sus, "frac_2x_bf16_sustained": achieved / peak_sus,
                "peak_nominal": 4500.0, "frac_nominal": achieved / 4500.0,
                "avg_launch_ms": gemm_total_ms / max(n_gemm, 1), "launches_timed": n_gemm,
                "share_of_step": gemm_total_ms / ms_step}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8xint4->int32", "data": "synthetic",
        "config": {"workload": WORKLOAD if a.workload == "qwenvl-9.6b" else stack.describe(),
                   "tokens_per_gpu_per_step": a.tokens, "seq_len": a.seq_len,
                   "visual_tokens_per_seq": "dynamic 64..seq_len/2" if a.dynamic_visual else 256, "layers": a.layers, "parallelism": f"dp{world}",
                   "sibling_fusion": not a.no_fuse,
                   "l2": "inputs+weights per step (>3 GB) exceed the 126 MB L2; no flush needed"},
        "ttft_ms_per_request": ms_step / n_seq,
        "ttft_single_request_ms": ms_single,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h_host.numel() * 2) * world,
                "d2h_bytes_per_step": int(out_host.numel() * 2) * world},
        "gpu_launches": stack.launches_per_step() * a.steps,
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.workload == "qwenvl-9.6b":
        cores = os.cpu_count() or 1
        tok = 1024
        per_layer = cpu_layer_sample(tok)
        line["cpu_baseline"] = {
            "value": tok / (per_layer * a.layers), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 decoder layer (7 fake-quant Linears, online Hadamard K=172) on {tok} tokens, x{a.layers} layers",
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
