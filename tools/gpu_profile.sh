#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the hot kernels.
R=${1:-r1}
mkdir -p gpurun_out
echo "=== launch list (bench, 2 layers)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 56 -c 42 --csv \
  --log-file gpurun_out/launches_${R}.csv python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${R}.log 2>&1
tail -2 gpurun_out/launches_${R}.log | cut -c1-300
echo "=== full: gemm"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:w4a8_gemm_kernel -s 3 -c 3 -f -o gpurun_out/prof_gemm_${R} python tools/profile_target.py gemm 16384 > gpurun_out/ncu_gemm_${R}.log 2>&1
echo "=== full: hadamard tc + quant + minmax + fwht"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hadk_tc_kernel|act_quant_i8_kernel|minmax_tensor_kernel|fwht_cta_kernel" -c 12 -f -o gpurun_out/prof_mem_${R} python tools/profile_target.py all 16384 > gpurun_out/ncu_mem_${R}.log 2>&1
tail -2 gpurun_out/ncu_mem_${R}.log
echo "=== gemm probes"
for d in 0 1 16 4; do
  echo "--- MQ_GEMM_DEBUG=$d" | tee -a gpurun_out/probe_${R}.log
  MQ_GEMM_DEBUG=$d timeout 300 python tools/microbench.py --only gemm --quick 2>&1 | grep "M4096" | cut -c1-200 | tee -a gpurun_out/probe_${R}.log
done
ls -la gpurun_out | tail -12
