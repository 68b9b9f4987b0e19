#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "hadamard or hadU" 2>&1 | tail -15 | tee gpurun_out/r15.log
timeout 500 python tools/microbench.py --only had 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['kernel'], r['shape'], 'ms=%.4f best=%.4f frac_hbm=%.3f' % (r['ms'], r['ms_best'], r.get('frac_hbm',0)))
" | tee -a gpurun_out/r15.log
