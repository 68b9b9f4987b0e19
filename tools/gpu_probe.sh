#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/probe9.log
for d in 0 2 4 6 8 16 30; do
  echo "--- hadtc MQ_HAD_DEBUG=$d" | tee -a gpurun_out/probe9.log
  MQ_HAD_DEBUG=$d timeout 300 python tools/microbench.py --only had 2>&1 | grep "k172" | grep 16384 | cut -c1-200 | tee -a gpurun_out/probe9.log
done
