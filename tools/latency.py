"""Small-M W4A8 latency as a model sees it: R distinct weight matrices used round-robin (cold in
L2 like the layers of a model, R x N x K / 2 bytes > 126 MB) while the kernel code stays hot --
unlike microbench.py's L2 flush, which also evicts the instructions and adds ~10 us to every
small launch.  CUDA events around `reps` back-to-back launches."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mquant_b200._lib import lib  # noqa: E402
from mquant_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="4096x4096,4096x11008,11008x4096")
    ap.add_argument("--ms", default="1,8,16,32,64,128,256,512,1024")
    ap.add_argument("--kernels", default="default")
    ap.add_argument("--reps", type=int, default=64)
    a = ap.parse_args()
    for shp in a.shapes.split(","):
        K, N = (int(v) for v in shp.split("x"))
        R = max(4, int(200e6 // (N * K // 2)) + 1)
        w4s = [ops.pack_w4(torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda")) for _ in range(R)]
        s_w = torch.rand(N, device="cuda") * 0.01
        for M in (int(v) for v in a.ms.split(",")):
            q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
            s_a = torch.rand(M, device="cuda") * 0.1
            out = torch.empty(M, N, dtype=torch.float16, device="cuda")
            for kern in a.kernels.split(","):
                if kern == "default":
                    lib().mq_debug_force_gemm_kernel(0)
                else:
                    lib().mq_debug_force_gemm_kernel({"v2": 3, "v4": 4, "v5": 5}.get(kern, 0))
                for i in range(R):
                    ops.w4a8_gemm(q, w4s[i], s_a, s_w, N, out=out)
                torch.cuda.synchronize()
                torch.cuda._sleep(2000000)  # let the host queue up the launches
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(a.reps):
                    ops.w4a8_gemm(q, w4s[i % R], s_a, s_w, N, out=out)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / a.reps
                print(json.dumps({"kernel": kern, "M": M, "K": K, "N": N, "us_per_launch": round(us, 2),
                                  "weight_GBps": round(N * K / 2 / us / 1e3, 1), "TOPs": round(2 * M * N * K / us / 1e6, 1)}),
                      flush=True)
        lib().mq_debug_force_gemm_kernel(0)
        del w4s


if __name__ == "__main__":
    main()
