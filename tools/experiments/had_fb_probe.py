"""Experiment: fused pow2 Hadamard + quantiser at 16384 x 4096: marginal cost of the pieces."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mquant_b200 import ops
from tools.microbench import timeit
M, n = 16384, 4096
x = (torch.randn(M, n, device="cuda") * 3).half()
mod = (torch.rand(M, device="cuda") < 0.4).to(torch.uint8)
scales = ops.make_scales(float(torch.tensor(0.1134, dtype=torch.float16)), float(torch.tensor(0.2871, dtype=torch.float16)), "cuda")
for fb in ("0", "1", "2"):
    os.environ["MQ_HAD_FB"] = fb
    for probe, what in ((0, "full"),):
        os.environ["MQ_HAD_PROBE"] = str(probe)
        sec, best = timeit(lambda: ops.hadamard(x, n=n, scales=scales, modality=mod, want_codes=True, want_f16=False))
        print(f"FB={fb} {what:18s}: {sec*1e6:.1f} us (best {best*1e6:.1f}) frac_hbm {3*M*n/sec/1e9/6551.7:.3f}")
