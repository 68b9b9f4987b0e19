#!/bin/bash
mkdir -p gpurun_out
echo "=== had tests" | tee gpurun_out/r9.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hadamard" 2>&1 | tail -5 | tee -a gpurun_out/r9.log
timeout 300 python -m pytest tests/test_gpu_layer.py -m gpu -q -x 2>&1 | tail -3 | tee -a gpurun_out/r9.log
timeout 500 python tools/microbench.py --only had 2>&1 | grep -E "k172" | cut -c1-250 | tee -a gpurun_out/r9.log
