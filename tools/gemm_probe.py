"""Bottleneck probes of the W4A8 GEMM kernels: time one shape under MQ_GEMM_DEBUG bit masks
(results of a probed run are garbage; only the time is meaningful)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mquant_b200._lib import lib  # noqa: E402
from mquant_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernels", default="v2,v4")
    ap.add_argument("--masks", default="0,4,8,12,1,16")
    ap.add_argument("--shapes", default="16384x4096x4096,16384x11008x4096")
    a = ap.parse_args()
    for shp in a.shapes.split(","):
        M, K, N = (int(v) for v in shp.split("x"))
        q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
        codes = torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda")
        w4 = ops.pack_w4(codes)
        s_a = torch.rand(M, device="cuda") * 0.1
        s_w = torch.rand(N, device="cuda") * 0.01
        out = torch.empty(M, N, dtype=torch.float16, device="cuda")
        for k in a.kernels.split(","):
            lib().mq_debug_force_gemm_kernel({"v2": 3, "v4": 4}.get(k, 0))
            for mask in a.masks.split(","):
                os.environ["MQ_GEMM_DEBUG"] = mask
                sec, best = timeit(lambda: ops.w4a8_gemm(q, w4, s_a, s_w, N, out=out), flush=True)
                print(f"{k} M{M} K{K} N{N} dbg={mask}: {sec * 1e3:.4f} ms  ({2 * M * N * K / sec / 1e12:.0f} TOP/s eq)",
                      flush=True)
        os.environ.pop("MQ_GEMM_DEBUG", None)


if __name__ == "__main__":
    main()
