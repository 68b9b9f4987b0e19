#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" | tee gpurun_out/r2_tests.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee -a gpurun_out/r2_tests.log
echo "=== bench" | tee gpurun_out/r2_bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -5 | tee -a gpurun_out/r2_bench.log
echo "=== microbench" | tee gpurun_out/r2_micro.log
timeout 900 python tools/microbench.py --json gpurun_out/microbench_r2.json 2>&1 | tail -80 | tee -a gpurun_out/r2_micro.log
echo "=== gemm probes" | tee gpurun_out/r2_probe.log
for d in 0 1 16 2 4 8 3 7; do
  echo "--- MQ_GEMM_DEBUG=$d" | tee -a gpurun_out/r2_probe.log
  MQ_GEMM_DEBUG=$d timeout 300 python tools/microbench.py --only gemm --quick 2>&1 | grep "M4096" | tee -a gpurun_out/r2_probe.log
done
echo "=== ncu gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a8_gemm_kernel -s 3 -c 2 -o gpurun_out/prof_gemm_r1 -f python tools/profile_target.py gemm 8192 > gpurun_out/r2_ncu_gemm.log 2>&1
tail -3 gpurun_out/r2_ncu_gemm.log
echo "=== ncu had/quant"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwht_cta_kernel|act_quant_i8_kernel|minmax_tensor_kernel" -c 6 -o gpurun_out/prof_mem_r1 -f python tools/profile_target.py all 8192 > gpurun_out/r2_ncu_mem.log 2>&1
tail -3 gpurun_out/r2_ncu_mem.log
ls -la gpurun_out
