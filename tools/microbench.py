"""Per-kernel timings on one GPU (CUDA events, L2 flushed between iterations).

python tools/microbench.py [--quick] [--json out.json]
Prints achieved GB/s (HBM-bound kernels, algorithmic bytes of SURVEY.md 8(d)) or TOP/s (GEMM,
2*M*N*K) next to the measured peaks of MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        # keep the GPU busy for ~0.25 ms so that the host-side cost of fn() (allocations, ctypes,
        # tensor-map encode) is hidden behind queued work instead of showing up as GPU idle time
        # between the two events
        torch.cuda._sleep(500000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    from mquant_b200 import _lib, hadamard_utils as hu, ops

    hbm, bf16, src = peaks()
    print(f"peaks ({src}): HBM {hbm:.0f} GB/s, bf16 {bf16:.0f} TFLOP/s; INT8 nominal 4500 TOP/s")
    rows = []

    def rec(name, shape, sec, best, bytes_=None, ops_=None):
        r = {"kernel": name, "shape": shape, "ms": sec * 1e3, "ms_best": best * 1e3}
        if bytes_:
            r["GBps"] = bytes_ / sec / 1e9
            r["frac_hbm"] = r["GBps"] / hbm
        if ops_:
            r["TOPs"] = ops_ / sec / 1e12
            r["frac_int8_nominal"] = r["TOPs"] / 4500.0
        rows.append(r)
        print(json.dumps(r))

    Ms = [4096] if a.quick else [1024, 4096, 16384]
    # calibrated-looking scales (fp16 values of max|x|/127); "round" scales such as 0.05 make exact
    # rounding ties ~50x more frequent and exercise the quantiser's slow path instead
    scales = ops.make_scales(float(torch.tensor(0.1134, dtype=torch.float16)),
                             float(torch.tensor(0.2871, dtype=torch.float16)), "cuda")
    want = lambda k: (not a.only) or k in a.only.split(",")

    if want("int8ref"):
        # library INT8 GEMM (cuBLASLt via torch._int_mm): the measured INT8 reference point
        n = 8192
        A = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
        B = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda").t()
        try:
            sec, best = timeit(lambda: torch._int_mm(A, B), flush=False)
            rec("cublaslt_int8_mm", f"{n}^3", sec, best, ops_=2 * n ** 3)
        except Exception as e:  # noqa: BLE001
            print("int8 library GEMM unavailable:", e)
        A16 = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
        B16 = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
        sec, best = timeit(lambda: A16 @ B16, flush=False)
        rec("cublas_bf16_mm", f"{n}^3", sec, best, ops_=2 * n ** 3)
        del A, B, A16, B16

    for M in Ms:
        for C in (4096, 11008):
            x = (torch.randn(M, C, device="cuda") * 3).half()
            mod = (torch.rand(M, device="cuda") < 0.4).to(torch.uint8)
            if want("quant"):
                sec, best = timeit(lambda: ops.act_quant_i8(x, scales, mod))
                rec("act_quant_i8_msq", f"{M}x{C}", sec, best, bytes_=3 * M * C + M)
            if want("rmsn"):
                sec, best = timeit(lambda: ops.rmsn_quant_i8(x, C, 1e-6, scales=scales, modality=mod, want_codes=True,
                                                             want_f16=False))
                rec("rmsn_quant_i8", f"{M}x{C}", sec, best, bytes_=3 * M * C + M)
                sec, best = timeit(lambda: ops.rmsn_quant_i8(x, C, 1e-6, scales=scales, modality=mod, want_codes=True,
                                                             want_f16=True))
                rec("rmsn_quant_i8+y", f"{M}x{C}", sec, best, bytes_=5 * M * C + M)
            if want("fake"):
                sec, best = timeit(lambda: ops.act_fake_quant(x, scales, mod))
                rec("act_fake_quant", f"{M}x{C}", sec, best, bytes_=4 * M * C + M)
            if want("minmax"):
                st = torch.zeros(2, dtype=torch.float32, device="cuda")
                sec, best = timeit(lambda: ops.observer_minmax(x, st))
                rec("observer_minmax", f"{M}x{C}", sec, best, bytes_=2 * M * C)
                stc = torch.zeros(2 * C, dtype=torch.float32, device="cuda")
                sec, best = timeit(lambda: ops.observer_minmax(x, stc, True))
                rec("observer_minmax_cw", f"{M}x{C}", sec, best, bytes_=2 * M * C)
            if want("pct") and M <= 4096:
                sec, best = timeit(lambda: ops.observer_percentile(x))
                rec("observer_percentile", f"{M}x{C}", sec, best, bytes_=4 * M * C)
            if want("dyn"):
                sec, best = timeit(lambda: ops.act_quant_i8_dynamic(x))
                rec("act_quant_dynamic", f"{M}x{C}", sec, best, bytes_=3 * M * C)
            if want("had"):
                hk_, K_ = hu.get_hadK(C)
                hk = hu.hadK_device(K_, "cuda") if K_ > 1 else None
                sec, best = timeit(lambda: ops.hadamard(x, had_k=K_, hadk=hk, n=C, scales=scales, modality=mod,
                                                        want_codes=True, want_f16=False))
                rec(f"hadamard_quant_k{K_}", f"{M}x{C}", sec, best, bytes_=3 * M * C + M)
            del x

    if a.only and "hadsweep" in a.only.split(","):
        # fused Hadamard + INT8 quantiser at M = 16384 for every online-rotation
        # width of the model table; bytes = 2*n_in read + n written + M
        M = 16384
        for n, n_in in ((1024, 1024), (2048, 2048), (4096, 4096), (8192, 8192), (11008, 11008), (13824, 13696),
                        (14336, 14336), (19968, 18944), (5120, 5120)):
            x = (torch.randn(M, n_in, device="cuda") * 3).half()
            mod = (torch.rand(M, device="cuda") < 0.4).to(torch.uint8)
            hk_, K_ = hu.get_hadK(n)
            hk = hu.hadK_device(K_, "cuda") if K_ > 1 else None
            sec, best = timeit(lambda: ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod,
                                                    want_codes=True, want_f16=False))
            rec(f"hadamard_quant_k{K_}", f"{M}x{n_in}->{n}", sec, best, bytes_=(2 * n_in + n) * M + M)
            if K_ > 1 and n // K_ in (256, 512):
                sec, best = timeit(lambda: ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod,
                                                        want_codes=True, want_f16=False, fast_had=True))
                rec(f"hadamard_quant_k{K_}[fast_had]", f"{M}x{n_in}->{n}", sec, best, bytes_=(2 * n_in + n) * M + M)
            if K_ == 1:  # the previous default (mma.sync, warp per token) on the same input
                _lib.lib().mq_debug_force_hadamard_kernel(3)
                sec, best = timeit(lambda: ops.hadamard(x, had_k=K_, hadk=hk, n=n, scales=scales, modality=mod,
                                                        want_codes=True, want_f16=False))
                _lib.lib().mq_debug_force_hadamard_kernel(0)
                rec(f"hadamard_quant_k{K_}[mma.sync]", f"{M}x{n_in}->{n}", sec, best, bytes_=(2 * n_in + n) * M + M)
            del x

    if want("gemm"):
        shapes = [(4096, 4096), (4096, 11008), (11008, 4096)]
        for M in ([16, 4096] if a.quick else [16, 256, 1024, 4096, 16384]):
            for K, N in shapes:
                q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
                codes = torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda")
                w4 = ops.pack_w4(codes)
                s_a = torch.rand(M, device="cuda") * 0.1
                s_w = torch.rand(N, device="cuda") * 0.01
                out = torch.empty(M, N, dtype=torch.float16, device="cuda")
                kname = "default"
                sec, best = timeit(lambda: ops.w4a8_gemm(q, w4, s_a, s_w, N, out=out), flush=True)
                rec(f"w4a8_gemm[{kname}]", f"M{M} K{K} N{N}", sec, best,
                    bytes_=N * K // 2 + M * K + 2 * M * N + 4 * N, ops_=2 * M * N * K)
                del q, codes, w4, out
    if want("gated"):
        # Qwen-VL's MLP input side, w1 | w2 as one GEMM.  fused = a1 * silu(a2) in the
        # epilogue ([M, F] written); unfused = the sibling-fused GEMM ([M, 2F] written) + torch's silu and mul
        K, F = 4096, 11008
        for M in ([4096] if a.quick else [4096, 16384]):
            q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
            c1 = torch.randint(-8, 7, (F, K), dtype=torch.int8, device="cuda")
            c2 = torch.randint(-8, 7, (F, K), dtype=torch.int8, device="cuda")
            s_a = torch.rand(M, device="cuda") * 0.05
            s_w = torch.rand(2 * F, device="cuda") * 0.01
            w4g = ops.pack_w4(ops.interleave_gated(c1, c2))
            w4c = ops.pack_w4(torch.cat([c1, c2]))
            del c1, c2
            out = torch.empty(M, F, dtype=torch.float16, device="cuda")
            full = torch.empty(M, 2 * F, dtype=torch.float16, device="cuda")
            sec, best = timeit(lambda: ops.w4a8_gemm_gated(q, w4g, s_a, s_w, F, out=out))
            rec("w4a8_gemm_gated[fused]", f"M{M} K{K} F{F}", sec, best, bytes_=F * K + M * K + 2 * M * F, ops_=4 * M * F * K)
            sec, best = timeit(lambda: ops.w4a8_gemm(q, w4c, s_a, s_w, 2 * F, out=full))
            rec("w4a8_gemm[w1|w2]", f"M{M} K{K} N{2 * F}", sec, best, bytes_=F * K + M * K + 4 * M * F, ops_=4 * M * F * K)

            def unfused():
                ops.w4a8_gemm(q, w4c, s_a, s_w, 2 * F, out=full)
                return full[:, :F] * torch.nn.functional.silu(full[:, F:])

            sec, best = timeit(unfused)
            rec("w4a8_gemm[w1|w2]+torch silu*mul", f"M{M} K{K} F{F}", sec, best, ops_=4 * M * F * K)
            del q, w4g, w4c, out, full
    if want("gptq"):
        # one add_batch of a 2048-token sample (fp16 X^T X into the fp32 Hessian) against the
        # reference's expression on the same device, and one 128-column block of fasterquant
        import math
        for C in (4096, 11008):
            T = 2048
            x = (torch.randn(T, C, device="cuda") * 2).half()
            H = torch.zeros(C, C, device="cuda")
            sec, best = timeit(lambda: ops.gptq_hessian_update(x, H, 3, 1))
            rec("gptq_hessian_update", f"T{T} C{C}", sec, best, bytes_=8 * C * C + 2 * T * C, ops_=2 * T * C * C)

            def ref_add_batch():  # gptq_utils.py:182-195
                H.mul_(3 / 4)
                inp = math.sqrt(2 / 4) * x.t().float()
                H.add_(inp.matmul(inp.t()))

            sec, best = timeit(ref_add_batch)
            rec("gptq_add_batch[torch fp32]", f"T{T} C{C}", sec, best, ops_=2 * T * C * C)
            del x, H
        nrows, cols = 4096, 4096
        W = torch.randn(nrows, cols, device="cuda") * 0.05
        A = torch.randn(cols, cols, device="cuda")
        Hinv = torch.linalg.cholesky(A @ A.t() / cols + 0.05 * torch.eye(cols, device="cuda"), upper=True).contiguous()
        scale = W.abs().amax(dim=1) / 7
        sec, best = timeit(lambda: ops.gptq_block(W, 0, 128, Hinv, scale, 7))
        rec("gptq_block_f32", f"rows{nrows} x 128 columns", sec, best)
    if want("decode"):
        for K, N in [(4096, 4096), (4096, 11008), (11008, 4096)]:
            codes = torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda")
            w4 = ops.pack_w4(codes)
            s_w = torch.rand(N, device="cuda") * 0.01
            for M in (1, 2, 4, 8, 16):
                q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
                s_a = torch.rand(M, device="cuda") * 0.1
                out = torch.empty(M, N, dtype=torch.float16, device="cuda")
                for name, env in (("w4a8_gemv", None), ("w4a8_gemm_tc", "v2")):
                    _lib.lib().mq_debug_force_gemm_kernel(3 if env else 0)
                    sec, best = timeit(lambda: ops.w4a8_gemm(q, w4, s_a, s_w, N, out=out), flush=True, iters=20)
                    rec(name, f"M{M} K{K} N{N}", sec, best, bytes_=N * K // 2 + M * K + 2 * M * N + 4 * N,
                        ops_=2 * M * N * K)
                _lib.lib().mq_debug_force_gemm_kernel(0)
        # event-timing floor: an empty-ish launch after the L2 flush
        z = torch.zeros(32, device="cuda")
        sec, best = timeit(lambda: z.add_(1.0), flush=True, iters=20)
        rec("timing_floor_tiny_kernel", "32 floats", sec, best, bytes_=256)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
