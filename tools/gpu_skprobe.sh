for s in 2 4 8; do echo "split cap $s"; MQ_SK_SPLIT=$s python tools/sweep.py --dims 4096 --ms 64,256 2>&1 | grep w4a8 | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  M%-5d %6.1f us' % (r['M'], r['us']))"; done
echo "no atomics"; MQ_SK_PROBE=1 python tools/sweep.py --dims 4096 --ms 64,256 2>&1 | grep w4a8 | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  M%-5d %6.1f us' % (r['M'], r['us']))"
echo "v2"; MQ_GEMM_KERNEL=v2 python tools/sweep.py --dims 4096 --ms 64,256 2>&1 | grep w4a8 | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  M%-5d %6.1f us' % (r['M'], r['us']))"
