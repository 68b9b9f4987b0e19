#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -3 | tee gpurun_out/r11.log
timeout 500 python tools/microbench.py --only decode 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['kernel'], r['shape'], 'us=%.1f best=%.1f GB/s=%.0f' % (r['ms']*1e3, r['ms_best']*1e3, r.get('GBps',0)))
" | tee -a gpurun_out/r11.log
