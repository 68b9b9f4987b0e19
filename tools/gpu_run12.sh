#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tcgen05_matches" 2>&1 | tail -5 | tee gpurun_out/r12.log
for k in v2 v4; do
MQ_GEMM_KERNEL=$k timeout 500 python tools/microbench.py --only gemm 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    if 'M16 ' in r['shape'] or 'M256 ' in r['shape']: continue
    print(r['kernel'], r['shape'], 'ms=%.4f best=%.4f TOPs=%.0f' % (r['ms'], r['ms_best'], r.get('TOPs',0)))
" | tee -a gpurun_out/r12.log
done
