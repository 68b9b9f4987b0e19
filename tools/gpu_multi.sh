#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/multi_r2.log
echo "=== multi-gpu test" | tee -a gpurun_out/multi_r2.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 | tee -a gpurun_out/multi_r2.log
echo "=== calibration InternVL2-8B (32 layers, 256 samples) N=2, verified against sequential" | tee -a gpurun_out/multi_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/calib_bench.py --verify 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/multi_r2.log
echo "=== calibration N=1 (no verify)" | tee -a gpurun_out/multi_r2.log
timeout 900 python tools/calib_bench.py 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/multi_r2.log
echo "=== bench N=2 (torchrun)" | tee -a gpurun_out/multi_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-4000 | tee -a gpurun_out/multi_r2.log
echo "=== bench qwen2vl-7b dynamic visual N=2" | tee -a gpurun_out/multi_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --workload qwen2vl-7b --dynamic-visual 2>&1 | tail -1 | cut -c1-700 | tee -a gpurun_out/multi_r2.log
