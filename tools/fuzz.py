"""Randomised cross-checks on the GPU (no CPU oracle involved: every case compares a fast kernel
with the simple CUDA-core kernel of the same contract).

  W4A8 GEMM     default dispatch (decode / split-K / 128x256 / 256x256 kernels) vs mq_w4a8_gemm_simt,
                bit-exact, random M, N, K, per-row scales, bias, rank-1 term
  Hadamard      tensor-core kernel vs the CUDA-core kernel, codes within 1 and rarely different
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mquant_b200 import hadamard_utils as hu, ops  # noqa: E402
from mquant_b200._lib import lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=150)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(a.seed)
    cg = torch.Generator().manual_seed(a.seed)

    def ri(lo, hi):
        return int(torch.randint(lo, hi + 1, (1,), generator=cg))

    bad = 0
    for i in range(a.cases):
        big = ri(0, 9) == 0
        M = ri(1, 20000 if big else 3000)
        N = 8 * ri(1, 700 if big else 300)
        K = 128 * ri(1, 24 if big else 16)
        per_row, bias, r1 = ri(0, 1), ri(0, 1), ri(0, 1)
        q = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda", generator=g)
        codes = torch.randint(-8, 8, (N, K), dtype=torch.int8, device="cuda", generator=g)
        if r1:
            codes[:, 0] = 0
        w4 = ops.pack_w4(codes)
        s_a = (torch.rand(M if per_row else 1, device="cuda", generator=g) * 0.1 + 1e-3)
        s_w = torch.rand(N, device="cuda", generator=g) * 0.01 + 1e-4
        b = (torch.randn(N, device="cuda", generator=g)).half() if bias else None
        x0 = torch.randn(M, device="cuda", generator=g).half() if r1 else None
        w0 = torch.randn(N, device="cuda", generator=g).half() if r1 else None
        y = ops.w4a8_gemm(q, w4, s_a, s_w, N, bias=b, x0=x0, w0=w0)
        ref = ops.w4a8_gemm(q, w4, s_a, s_w, N, bias=b, x0=x0, w0=w0, simt=True)
        if not torch.equal(y.view(torch.int16), ref.view(torch.int16)):
            bad += 1
            print(f"GEMM MISMATCH M={M} N={N} K={K} per_row={per_row} bias={bias} r1={r1}: "
                  f"{int((y.view(torch.int16) != ref.view(torch.int16)).sum())} elements", flush=True)
    print(f"gemm: {a.cases} cases, {bad} mismatching")
    hbad = 0
    shapes = [(11008, 11008), (13824, 13696), (14336, 14336), (19968, 18944), (5120, 5120), (3584, 3584), (3072, 3072)]
    for i in range(max(10, a.cases // 5)):
        n, n_in = shapes[ri(0, len(shapes) - 1)]
        M = ri(1, 700)
        _, K_ = hu.get_hadK(n)
        hk = hu.hadK_device(K_, "cuda")
        x = (torch.randn(M, n_in, device="cuda", generator=g) * 3).half()
        mod = (torch.rand(M, device="cuda", generator=g) < 0.4).to(torch.uint8)
        scales = ops.make_scales(0.07 + 0.01 * ri(0, 9), 0.2 + 0.01 * ri(0, 9), "cuda")
        split = bool(ri(0, 1))
        lib().mq_debug_force_hadamard_kernel(0)
        r = ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod, want_codes=True, want_f16=bool(ri(0, 1)),
                         split=split)
        lib().mq_debug_force_hadamard_kernel(1)
        s = ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod, want_codes=True, want_f16=True, split=split)
        lib().mq_debug_force_hadamard_kernel(0)
        d = (r["q"].int() - s["q"].int()).abs()
        if int(d.max()) > 1 or float((d > 0).float().mean()) > 3e-3:
            hbad += 1
            print(f"HADAMARD MISMATCH n={n} n_in={n_in} M={M} split={split}: max {int(d.max())} rate {float((d > 0).float().mean()):.2e}", flush=True)
    print(f"hadamard: {max(10, a.cases // 5)} cases, {hbad} out of tolerance")
    sys.exit(1 if (bad or hbad) else 0)


if __name__ == "__main__":
    main()
