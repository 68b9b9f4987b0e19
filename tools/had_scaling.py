"""Fused Hadamard + INT8 quantiser: fixed cost vs steady state.

Two timings per (n, M), both with CUDA events on the launching stream:
  flushed   one launch per iteration after an L2 flush (tools/microbench.py: the cache is full of the
            flush buffer's DIRTY lines, every line the kernel allocates writes one of them back)
  stream    R launches back to back on R distinct input / output sets (together > 4 x L2, so no launch
            finds its operands in the cache; nothing is flushed), time / R: the figure a pipeline of
            kernels sees, without the ~6 us event floor of a lone launch
and a least-squares fit  t(M) = fixed + per_token * M  over the M sweep.

python tools/had_scaling.py [--json out.json] [--force 0|3|4]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--force", type=int, default=0)
    ap.add_argument("--sizes", default="4096,8192,11008,13824,19968")
    a = ap.parse_args()
    from microbench import peaks, timeit

    from mquant_b200 import _lib, hadamard_utils as hu, ops

    hbm = peaks()[0]
    _lib.lib().mq_debug_force_hadamard_kernel(a.force)
    scales = ops.make_scales(float(torch.tensor(0.1134, dtype=torch.float16)),
                             float(torch.tensor(0.2871, dtype=torch.float16)), "cuda")
    n_in_of = {13824: 13696, 19968: 18944}
    rows = []
    for n in [int(v) for v in a.sizes.split(",")]:
        n_in = n_in_of.get(n, n)
        K_ = hu.get_hadK(n)[1]
        hk = hu.hadK_device(K_, "cuda") if K_ > 1 else None
        pts = []
        for M in (4096, 8192, 16384, 32768):
            bytes_ = (2 * n_in + n) * M + M
            R = max(4, int(4 * 126e6 * 1.5 / bytes_) + 1)
            xs = [(torch.randn(M, n_in, device="cuda") * 3).half() for _ in range(R)]
            mod = (torch.rand(M, device="cuda") < 0.4).to(torch.uint8)
            call = lambda x: ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod, want_codes=True,
                                          want_f16=False)
            sec, best = timeit(lambda: call(xs[0]))
            keep = []

            def burst():
                keep.clear()
                for x in xs:
                    keep.append(call(x))  # outputs stay alive: R distinct output buffers

            for _ in range(2):
                burst()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                torch.cuda._sleep(2000000)  # the host enqueues the burst behind this
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                burst()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3 / R)
            ts.sort()
            st = ts[len(ts) // 2]
            r = {"kernel": f"hadamard_quant_k{K_}", "shape": f"{M}x{n_in}->{n}", "ms_flushed": sec * 1e3,
                 "frac_hbm_flushed": bytes_ / sec / 1e9 / hbm, "ms_stream": st * 1e3,
                 "frac_hbm_stream": bytes_ / st / 1e9 / hbm, "launches_per_burst": R}
            rows.append(r)
            print(json.dumps(r), flush=True)
            pts.append((M, sec, st))
            del xs, keep
        for idx, tag in ((1, "flushed"), (2, "stream")):
            Ms = torch.tensor([p[0] for p in pts], dtype=torch.float64)
            t = torch.tensor([p[idx] for p in pts], dtype=torch.float64)
            A = torch.stack([torch.ones_like(Ms), Ms], 1)
            sol = torch.linalg.lstsq(A, t[:, None]).solution[:, 0]
            fit = {"fit": tag, "n": n, "fixed_us": float(sol[0]) * 1e6, "ns_per_token": float(sol[1]) * 1e9,
                   "steady_frac_hbm": (2 * n_in + n) / float(sol[1]) / 1e9 / hbm}
            rows.append(fit)
            print(json.dumps(fit), flush=True)
    _lib.lib().mq_debug_force_hadamard_kernel(0)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
