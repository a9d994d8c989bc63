for sc in 256 128 64; do
echo "BTG_SPLIT_COST=$sc"
BTG_SPLIT_COST=$sc BTG_GIBBS_TIMING=1 BIGS=128 timeout 600 python tools/prof_real.py 0.33 2>&1 | grep -E "k_estimate_genotypes|estimateGenotypes reconverge=0" | tail -2
done
