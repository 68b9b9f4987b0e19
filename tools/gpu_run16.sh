#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm" 2>&1 | tail -4 | tee gpurun_out/r16.log
timeout 600 python tools/sweep.py --dims 4096,11008 --ms 16,32,64,128,256,512,1024,2048 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    if r['kernel'] == 'w4a8_gemm': print('gemm M%-6d K%-6d N%-6d %8.1f us %7.0f TOP/s' % (r['M'], r['K'], r['N'], r['us'], r['TOPs']))
" | tee -a gpurun_out/r16.log
