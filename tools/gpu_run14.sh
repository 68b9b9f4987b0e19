#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r14.log
timeout 500 python tools/microbench.py --only gemm 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['kernel'], r['shape'], 'ms=%.4f best=%.4f TOPs=%.0f' % (r['ms'], r['ms_best'], r.get('TOPs',0)))
" | tee -a gpurun_out/r14.log
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/r14_bench.json
