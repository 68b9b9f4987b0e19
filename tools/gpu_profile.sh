#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the hot kernels.
R=${1:-r1}
mkdir -p gpurun_out
echo "=== launch list (bench, 2 layers): our kernels only, warm-up skipped"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:"w4a8_gemm|act_quant_i8_kernel|hadk_tc|hadp2|fwht|rmsn" -s 48 -c 16 --csv \
  --log-file gpurun_out/launches_${R}.csv python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${R}.log 2>&1
tail -1 gpurun_out/launches_${R}.log | cut -c1-200
echo "=== full: gemm (the 4 shapes of the bench step at M=16384)"
cat > /tmp/gemm_shapes.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from mquant_b200 import ops
M = 16384
for K, N in ((4096, 12288), (4096, 4096), (4096, 22016), (11008, 4096), (-4096, 4096), (-11008, 4096)):  # q|k|v, o, w1|w2 (sibling fused), down; K < 0: one 1024-token request (v5 kernel)
    M = 1024 if K < 0 else 16384
    K = abs(K)
    q = torch.randint(-128, 127, (M, K), dtype=torch.int8, device="cuda")
    w4 = ops.pack_w4(torch.randint(-8, 7, (N, K), dtype=torch.int8, device="cuda"))
    s_a = torch.rand(M, device="cuda") * 0.1
    s_w = torch.rand(N, device="cuda") * 0.01
    for _ in range(3):
        ops.w4a8_gemm(q, w4, s_a, s_w, N)
    torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a8_gemm -f -o /tmp/prof_gemm_${R} python /tmp/gemm_shapes.py > gpurun_out/ncu_gemm_${R}.log 2>&1
python tools/summarize_ncu.py /tmp/prof_gemm_${R}.ncu-rep gpurun_out/${R}_gemm_full.csv
echo "=== full: hadamard (tc3 K=172, tc2 K=28, p2 pow2) + quant + minmax + rmsn + gptq hessian"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hadk_tc|hadp2_tc|act_quant_i8_kernel|act_fake_quant_i8|act_quant_i8_dynamic|minmax_tensor_kernel|minmax_channel|fwht_|rmsn_quant|hessian_kernel" -c 48 -f -o /tmp/prof_mem_${R} python tools/profile_target.py all 16384 > gpurun_out/ncu_mem_${R}.log 2>&1
tail -2 gpurun_out/ncu_mem_${R}.log
python tools/summarize_ncu.py /tmp/prof_mem_${R}.ncu-rep gpurun_out/${R}_membound_full.csv
# (the .ncu-rep files stay on the box: gpurun_out/ is limited to 64 MiB; the per-kernel details pages are text)
ncu -i /tmp/prof_mem_${R}.ncu-rep --page details -k regex:"hadk_tc3|hadp2_tc|fwht_mma|rmsn_quant|hessian|w4a8_gemm_v5" 2>/dev/null | grep -v "^    -\+$" > gpurun_out/${R}_membound_details.txt
ncu -i /tmp/prof_gemm_${R}.ncu-rep --page details 2>/dev/null | grep -v "^    -\+$" > gpurun_out/${R}_gemm_details.txt
echo "=== microbench (all)"
timeout 900 python tools/microbench.py --json gpurun_out/microbench_${R}.json > gpurun_out/microbench_${R}.log 2>&1
timeout 600 python tools/microbench.py --only hadsweep,gated,gptq --json gpurun_out/microbench2_${R}.json > gpurun_out/microbench2_${R}.log 2>&1
timeout 400 python tools/had_scaling.py --json gpurun_out/had_scaling_${R}.json > gpurun_out/had_scaling_${R}.log 2>&1
timeout 600 python tools/latency.py --shapes 4096x4096,4096x12288,4096x22016,11008x4096 --ms 256,1024 --kernels default,v2,v4,v5 > gpurun_out/latency_${R}.log 2>&1
tail -3 gpurun_out/microbench_${R}.log | cut -c1-200
echo "=== bench"
timeout 900 python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
cut -c1-300 gpurun_out/bench_${R}.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${R}.json 2>/dev/null
cut -c1-200 gpurun_out/bench_ref_${R}.json
ls -la gpurun_out | tail -14
