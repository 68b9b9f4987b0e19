#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" | tee gpurun_out/r3_tests.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee -a gpurun_out/r3_tests.log
echo "=== microbench" | tee gpurun_out/r3_micro.log
timeout 900 python tools/microbench.py --only quant,minmax,dyn,had,gemm 2>&1 | grep -E "16384|M4096" | tee -a gpurun_out/r3_micro.log
