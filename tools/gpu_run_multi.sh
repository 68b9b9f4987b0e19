#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/multi.log
echo "=== multi-gpu test" | tee -a gpurun_out/multi.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8 | tee -a gpurun_out/multi.log
echo "=== bench N=2 (torchrun)" | tee -a gpurun_out/multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -4 | cut -c1-1500 | tee -a gpurun_out/multi.log
echo "=== bench --impl reference N=2" | tee -a gpurun_out/multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -3 | cut -c1-600 | tee -a gpurun_out/multi.log
