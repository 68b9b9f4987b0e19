# same-box A/B of the Hadamard sweep: product library vs mquant_b200/libmquant_b200_alt.so
for i in 1 2; do
python tools/microbench.py --only hadsweep 2>&1 | grep "$1" | cut -c1-140
echo "--- alt"
MQUANT_B200_LIB=$PWD/mquant_b200/libmquant_b200_alt.so python tools/microbench.py --only hadsweep 2>&1 | grep "$1" | cut -c1-140
echo "==="
done
