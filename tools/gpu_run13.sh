#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tcgen05_matches and v4" 2>&1 | tail -5 | tee gpurun_out/r13.log
timeout 300 python tools/gemm_probe.py --kernels v4 --masks 0,4,8,12,1 2>&1 | tee -a gpurun_out/r13.log
