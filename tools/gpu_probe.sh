#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe6.log
for d in 0 256 9 265 11 267; do
  echo "--- v2 MQ_GEMM_DEBUG=$d" | tee -a gpurun_out/probe6.log
  MQ_GEMM_KERNEL=v2 MQ_GEMM_DEBUG=$d timeout 300 python tools/microbench.py --only gemm --quick 2>&1 | grep "M4096 K4096 N11008" | cut -c1-200 | tee -a gpurun_out/probe6.log
done
