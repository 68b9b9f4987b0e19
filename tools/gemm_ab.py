"""A/B of the tensor-core W4A8 kernels by forcing the dispatcher (mq_debug_force_gemm_kernel): the 128 x 256
kernel (v2), the 256 x 256 kernel (v4), the TMEM-resident-weights kernel (v5) and the default choice.
CUDA events, L2 flushed (tools/microbench.timeit).  python tools/gemm_ab.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mquant_b200 import _lib, ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--Ms", default="256,512,1024,2048,4096,16384")
    a = ap.parse_args()
    rows = []
    shapes = [(4096, 4096), (4096, 12288), (4096, 22016), (11008, 4096)]
    for M in [int(m) for m in a.Ms.split(",")]:
        for K, N in shapes:
            q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
            w4 = ops.pack_w4(torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda"))
            s_a = torch.rand(M, device="cuda") * 0.1
            s_w = torch.rand(N, device="cuda") * 0.01
            out = torch.empty(M, N, dtype=torch.float16, device="cuda")
            r = {"M": M, "K": K, "N": N}
            for name, force in (("default", 0), ("v2", 3), ("v4", 4), ("v5", 5)):
                _lib.lib().mq_debug_force_gemm_kernel(force)
                sec, best = timeit(lambda: ops.w4a8_gemm(q, w4, s_a, s_w, N, out=out), iters=12)
                r[name + "_us"] = round(sec * 1e6, 1)
                r[name + "_TOPs"] = round(2 * M * N * K / sec / 1e12, 0)
            _lib.lib().mq_debug_force_gemm_kernel(0)
            r["plan"] = _lib.lib().mq_w4a8_gemm_plan(M, N, K, _lib.lib().mq_w4a8_gemm_workspace_bytes(M, N, K))
            rows.append(r)
            print(json.dumps(r), flush=True)
            del q, w4, out
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
