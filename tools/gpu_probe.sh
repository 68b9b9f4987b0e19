#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from mquant_b200 import ops
M,K,N=4096,4096,11008
q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
w4 = ops.pack_w4(torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda"))
s_a = torch.rand(M, device="cuda") * 0.1
s_w = torch.rand(N, device="cuda") * 0.01
for _ in range(4):
    ops.w4a8_gemm(q, w4, s_a, s_w, N)
torch.cuda.synchronize()
PY
MQ_GEMM_2CTA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a8_gemm_2cta -s 2 -c 1 -f -o gpurun_out/prof_gemm2cta python /tmp/one.py > gpurun_out/ncu_2cta.log 2>&1
MQ_GEMM_2CTA=1 MQ_GEMM_DEBUG=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a8_gemm_2cta -s 2 -c 1 -f -o gpurun_out/prof_gemm2cta_dbg9 python /tmp/one.py >> gpurun_out/ncu_2cta.log 2>&1
tail -3 gpurun_out/ncu_2cta.log
