#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe8.log
MQ_GEMM_KERNEL=v3 timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -3 | tee -a gpurun_out/probe8.log
MQ_GEMM_KERNEL=v3 timeout 600 python tools/microbench.py --only gemm 2>&1 | grep -E "16384|M4096|M1024|M256|M16 " | cut -c1-330 | tee -a gpurun_out/probe8.log
for d in 1 16 4; do
  echo "--- v3 MQ_GEMM_DEBUG=$d" | tee -a gpurun_out/probe8.log
  MQ_GEMM_KERNEL=v3 MQ_GEMM_DEBUG=$d timeout 300 python tools/microbench.py --only gemm --quick 2>&1 | grep "M4096 K4096 N11008" | cut -c1-200 | tee -a gpurun_out/probe8.log
done
