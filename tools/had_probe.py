"""Bottleneck probes of the tensor-core Hadamard kernels: time one shape under MQ_HAD_DEBUG masks
(results of a probed run are garbage; only the time is meaningful)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mquant_b200 import hadamard_utils as hu, ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--masks", default="0,2,4,8,16,18,20,22,30")
    ap.add_argument("--n", type=int, default=11008)
    ap.add_argument("--M", type=int, default=16384)
    ap.add_argument("--kernel", default="")
    ap.add_argument("--timeline", type=int, default=-1, help="MQ_HAD_DEBUG mask to run with bit 256 and dump stamps")
    ap.add_argument("--timeline3", action="store_true", help="clock64 timeline of CTA 0 of hadk_tc3 (probes build)")
    a = ap.parse_args()
    hk_, K = hu.get_hadK(a.n)
    hk = hu.hadK_device(K, "cuda")
    x = (torch.randn(a.M, a.n, device="cuda") * 3).half()
    mod = (torch.rand(a.M, device="cuda") < 0.4).to(torch.uint8)
    # calibrated-looking scales: fp16 VALUES (MinmaxObserver), the division-free quantiser path
    scales = ops.make_scales(float(torch.tensor(0.1134, dtype=torch.float16)), float(torch.tensor(0.2871, dtype=torch.float16)), "cuda")
    if a.timeline3:
        import ctypes
        import numpy as np
        from mquant_b200 import _lib
        for _ in range(2):
            ops.hadamard(x, had_k=K, hadk=hk, n=a.n, scales=scales, modality=mod, want_codes=True, want_f16=False)
        torch.cuda.synchronize()
        buf = np.zeros(96 * 16, dtype=np.int64)
        fn = ctypes.CDLL(_lib.LIB_PATH).mq_debug_hadk_tc3_stamps if hasattr(_lib, "LIB_PATH") else _lib.lib().mq_debug_hadk_tc3_stamps
        fn.argtypes = [ctypes.c_void_p]
        assert fn(buf.ctypes.data) == 0
        ts = buf.reshape(96, 16)
        t0 = int(ts[0, 0])
        names = ["tma0", "m1_xfull", "m1_d1emp", "m1_done", "m2_tfull", "m2_d2emp", "m2_done", "e2a_get", "e2a_done",
                 "e2b_get", "e2b_done", "e1_temp", "e1_d1full", "e1_tile0", "e1_done"]
        print("tok  " + " ".join(f"{n:>9s}" for n in names))
        for u in list(range(0, 14)) + list(range(60, 72)):
            print(f"{u:4d} " + " ".join(f"{int(ts[u, i]) - t0:9d}" for i in range(16)))
        return
    if a.timeline >= 0:
        os.environ["MQ_HAD_DEBUG"] = str(a.timeline | 256)
        r = ops.hadamard(x, had_k=K, hadk=hk, n=a.n, scales=scales, modality=mod, want_codes=True, want_f16=False)
        torch.cuda.synchronize()
        ts = r["q"].view(-1)[: 64 * 8 * 8].view(torch.int64).view(64, 8).cpu()
        t0 = int(ts[0].min())
        names = ["prod", "xf_first", "mma_tfull", "mma_d2empty", "mma_commit", "epi_d2full", "epi_done", "xf_last"]
        print("unit " + " ".join(f"{n:>11s}" for n in names))
        for u in list(range(0, 12)) + list(range(40, 46)):
            print(f"{u:4d} " + " ".join(f"{int(ts[u, i]) - t0:11d}" for i in range(8)))
        os.environ.pop("MQ_HAD_DEBUG", None)
        return
    for mask in a.masks.split(","):
        os.environ["MQ_HAD_DEBUG"] = mask
        sec, best = timeit(lambda: ops.hadamard(x, had_k=K, hadk=hk, n=a.n, scales=scales, modality=mod,
                                                want_codes=True, want_f16=False))
        print(f"n={a.n} K={K} M={a.M} dbg={mask}: {sec * 1e3:.4f} ms  ({3 * a.M * a.n / sec / 1e9:.0f} GB/s)", flush=True)
    os.environ.pop("MQ_HAD_DEBUG", None)


if __name__ == "__main__":
    main()
