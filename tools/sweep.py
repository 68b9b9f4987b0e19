"""BASELINE.json configs[4]: W4A8 qlinear + Hadamard kernel shape sweep
(K, N in {4096, 11008, 13696}, M = 16 .. 16384) against the INT8 / HBM rooflines.

    python tools/sweep.py --json gpurun_out/sweep.json

GEMM rows: TOP/s, fraction of the nominal 4500 TOP/s and of 2 x the measured bf16 burst; for
M <= 64 also the weight-stream GB/s.  Hadamard rows (online rotation of a K-wide input fused with
the static INT8 quantiser): GB/s over 3 B/element against the measured HBM peak.  13696 is padded
to 13824 = 108 x 128 by the reference's auto_pad_size (hadamard_utils.py:9-23)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mquant_b200 import hadamard_utils as hu, ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--dims", default="4096,11008,13696")
    ap.add_argument("--ms", default="16,64,256,1024,4096,16384")
    a = ap.parse_args()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(pk.get("hbm_gbs", 6550.0))
    int8_peak = 2.0 * float(pk.get("bf16_tflops", 1590.0))
    dims = [int(v) for v in a.dims.split(",")]
    Ms = [int(v) for v in a.ms.split(",")]
    rows = []

    def rec(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    for K in dims:
        for N in dims:
            codes = torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda")
            w4 = ops.pack_w4(codes)
            s_w = torch.rand(N, device="cuda") * 0.01
            del codes
            for M in Ms:
                q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
                s_a = torch.rand(M, device="cuda") * 0.1
                out = torch.empty(M, N, dtype=torch.float16, device="cuda")
                sec, best = timeit(lambda: ops.w4a8_gemm(q, w4, s_a, s_w, N, out=out), flush=True)
                tops = 2.0 * M * N * K / sec / 1e12
                byts = N * K // 2 + M * K + 2 * M * N
                rec(kernel="w4a8_gemm", M=M, K=K, N=N, us=sec * 1e6, TOPs=tops, frac_int8_nominal=tops / 4500.0,
                    frac_2x_bf16_burst=tops / int8_peak, GBps=byts / sec / 1e9, frac_hbm=byts / sec / 1e9 / hbm)
                del q, out
            del w4
    scales = ops.make_scales(0.1134, 0.2871, "cuda")
    for n_in in dims:
        n = hu.auto_pad_size(n_in)
        hk_, K_ = hu.get_hadK(n)
        hk = hu.hadK_device(K_, "cuda") if K_ > 1 else None
        for M in Ms:
            x = (torch.randn(M, n_in, device="cuda") * 3).half()
            mod = (torch.rand(M, device="cuda") < 0.25).to(torch.uint8)
            sec, best = timeit(lambda: ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod,
                                                    want_codes=True, want_f16=False))
            byts = 2 * M * n_in + M * n
            rec(kernel="hadamard_quant_i8", M=M, n_in=n_in, n=n, had_k=K_, us=sec * 1e6, GBps=byts / sec / 1e9,
                frac_hbm=byts / sec / 1e9 / hbm)
            del x
    if a.json:
        json.dump({"peaks": {"hbm_gbs": hbm, "int8_2x_bf16_burst_tops": int8_peak, "int8_nominal_tops": 4500.0},
                   "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
