#!/bin/bash
# First GPU pass: non-GEMM parity, then the tcgen05 GEMM under its own timeout, smoke, microbench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== non-GEMM parity" | tee gpurun_out/t1.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05 and not extreme" 2>&1 | tail -40 | tee -a gpurun_out/t1.log
echo "=== tcgen05 GEMM parity" | tee gpurun_out/t2.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tcgen05 or extreme" 2>&1 | tail -40 | tee -a gpurun_out/t2.log
echo "exit: $?" | tee -a gpurun_out/t2.log
echo "=== layer tests" | tee gpurun_out/t3.log
timeout 300 python -m pytest tests/test_gpu_layer.py -m gpu -q 2>&1 | tail -40 | tee -a gpurun_out/t3.log
echo "=== smoke" | tee gpurun_out/t4.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -20 | tee -a gpurun_out/t4.log
echo "=== microbench" | tee gpurun_out/t5.log
timeout 600 python tools/microbench.py --quick --json gpurun_out/microbench.json 2>&1 | tail -60 | tee -a gpurun_out/t5.log
