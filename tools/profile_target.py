"""Run each hot kernel a few times at the headline shapes (ncu target).
python tools/profile_target.py [gemm|quant|had|had172|had28|minmax|rmsn|hessian|all] [M]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mquant_b200 import hadamard_utils as hu, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
scales = ops.make_scales(float(torch.tensor(0.1134, dtype=torch.float16)), float(torch.tensor(0.2871, dtype=torch.float16)), "cuda")
mod = (torch.rand(M, device="cuda") < 0.25).to(torch.uint8)
reps = 3
if which in ("gemm", "all"):
    for K, N in ((4096, 11008), (11008, 4096), (4096, 4096)):
        q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
        w4 = ops.pack_w4(torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda"))
        s_a = torch.rand(M, device="cuda") * 0.1
        s_w = torch.rand(N, device="cuda") * 0.01
        for _ in range(reps):
            ops.w4a8_gemm(q, w4, s_a, s_w, N)
        torch.cuda.synchronize()
if which in ("quant", "all"):
    x = (torch.randn(M, 4096, device="cuda") * 3).half()
    for _ in range(reps):
        ops.act_quant_i8(x, scales, mod)
if which in ("minmax", "all"):
    x = (torch.randn(M, 4096, device="cuda") * 3).half()
    st = torch.zeros(2, device="cuda")
    for _ in range(reps):
        ops.observer_minmax(x, st)
if which in ("had", "all"):
    x = (torch.randn(M, 4096, device="cuda") * 3).half()
    for _ in range(reps):
        ops.hadamard(x, n=4096, scales=scales, modality=mod, want_codes=True, want_f16=False)
if which in ("had8192", "all"):  # power-of-two sizes: hadp2_tc_kernel (mq_hadamard_p2.cu)
    x = (torch.randn(M, 8192, device="cuda") * 3).half()
    for _ in range(reps):
        ops.hadamard(x, n=8192, scales=scales, modality=mod, want_codes=True, want_f16=False)
if which in ("had172", "all"):
    x = (torch.randn(M, 11008, device="cuda") * 3).half()
    hk = hu.hadK_device(172, "cuda")
    for _ in range(reps):
        ops.hadamard(x, had_k=172, hadk=hk, n=11008, scales=scales, modality=mod, want_codes=True, want_f16=False)
if which in ("had28", "all"):  # InternVL2-8B feed_forward.w2: 14336 = 28 x 512 (hadk_tc2_kernel)
    x = (torch.randn(M, 14336, device="cuda") * 3).half()
    hk = hu.hadK_device(28, "cuda")
    for _ in range(reps):
        ops.hadamard(x, had_k=28, hadk=hk, n=14336, scales=scales, modality=mod, want_codes=True, want_f16=False)
if which in ("fake", "all"):  # the op of an un-frozen wrapper: INT8 fake quantisation, per-token dynamic quantiser, channel-wise min/max
    x = (torch.randn(M, 4096, device="cuda") * 3).half()
    stc = torch.zeros(2 * 4096, device="cuda")
    for _ in range(reps):
        ops.act_fake_quant(x, scales, mod)
        ops.act_quant_i8_dynamic(x)
        ops.observer_minmax(x, stc, True)
if which in ("rmsn", "all"):
    x = (torch.randn(M, 4096, device="cuda") * 3).half()
    for _ in range(reps):
        ops.rmsn_quant_i8(x, 4096, 1e-6, scales=scales, modality=mod)
if which in ("hessian", "all"):  # GPTQ.add_batch on one 2048-token sample, C = 4096 and 11008
    for C in (4096, 11008):
        x = (torch.randn(2048, C, device="cuda") * 3).half()
        H = torch.zeros(C, C, device="cuda")
        for i in range(reps):
            ops.gptq_hessian_update(x, H, i, 1)
        del H
torch.cuda.synchronize()
print("done", which, M)
