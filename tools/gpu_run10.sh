#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "decode or tcgen05_matches" 2>&1 | tail -5 | tee gpurun_out/r10.log
timeout 500 python tools/microbench.py --only gemm 2>&1 | grep -E "M16 |M256" | cut -c1-300 | tee -a gpurun_out/r10.log
