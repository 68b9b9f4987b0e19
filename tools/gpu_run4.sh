#!/bin/bash
mkdir -p gpurun_out
echo "=== had tc tests" | tee gpurun_out/r4_tests.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tensor_core_path" 2>&1 | tail -30 | tee -a gpurun_out/r4_tests.log
echo "=== had tc tests with swapped LBO/SBO (MQ_HAD_DEBUG=1)" | tee -a gpurun_out/r4_tests.log
MQ_HAD_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tensor_core_path" 2>&1 | tail -8 | tee -a gpurun_out/r4_tests.log
echo "=== all gpu tests" | tee -a gpurun_out/r4_tests.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee -a gpurun_out/r4_tests.log
echo "=== microbench" | tee gpurun_out/r4_micro.log
timeout 900 python tools/microbench.py --only had,gemm 2>&1 | grep -E "16384|M4096|M1024" | tee -a gpurun_out/r4_micro.log
echo "=== bench" | tee gpurun_out/r4_bench.log
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee -a gpurun_out/r4_bench.log
