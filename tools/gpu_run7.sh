#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests (2cta)" | tee gpurun_out/r7_tests.log
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -5 | tee -a gpurun_out/r7_tests.log
echo "=== microbench" | tee gpurun_out/r7_micro.log
timeout 600 python tools/microbench.py --only gemm 2>&1 | grep -E "16384|M4096|M1024|M256" | cut -c1-330 | tee -a gpurun_out/r7_micro.log
for d in 1 3 9 11 16 4; do
  echo "--- 2CTA MQ_GEMM_DEBUG=$d" | tee -a gpurun_out/r7_micro.log
  MQ_GEMM_DEBUG=$d timeout 300 python tools/microbench.py --only gemm --quick 2>&1 | grep "M4096 K4096 N11008" | cut -c1-200 | tee -a gpurun_out/r7_micro.log
done
echo "=== bench" | tee gpurun_out/r7_bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee -a gpurun_out/r7_bench.log
