for i in 1 2; do
python tools/microbench.py --only gemm 2>&1 | grep "M16384\|M4096\|M1024" | cut -c1-120
echo "--- alt"
MQUANT_B200_LIB=$PWD/mquant_b200/libmquant_b200_alt.so python tools/microbench.py --only gemm 2>&1 | grep "M16384\|M4096\|M1024" | cut -c1-120
echo "==="
done
