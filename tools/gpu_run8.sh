#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r8*.log
for k in v2 v1 2cta; do
  echo "=== gemm tests kernel=$k" | tee -a gpurun_out/r8_tests.log
  MQ_GEMM_KERNEL=$k timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -3 | tee -a gpurun_out/r8_tests.log
done
for k in v2 v1; do
  echo "=== microbench kernel=$k" | tee -a gpurun_out/r8_micro.log
  MQ_GEMM_KERNEL=$k timeout 600 python tools/microbench.py --only gemm 2>&1 | grep -E "16384|M4096|M1024|M256|M16 " | cut -c1-330 | tee -a gpurun_out/r8_micro.log
done
for d in 1 3 9 11 16 4; do
  echo "--- v2 MQ_GEMM_DEBUG=$d" | tee -a gpurun_out/r8_micro.log
  MQ_GEMM_KERNEL=v2 MQ_GEMM_DEBUG=$d timeout 300 python tools/microbench.py --only gemm --quick 2>&1 | grep "M4096 K4096 N11008" | cut -c1-200 | tee -a gpurun_out/r8_micro.log
done
echo "=== bench" | tee gpurun_out/r8_bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee -a gpurun_out/r8_bench.log
