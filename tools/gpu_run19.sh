timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
BTG_GIBBS_TIMING=1 BIGS=128 timeout 600 python tools/prof_real.py 0.33 2>&1 | grep -E "k_estimate_genotypes|estimateGenotypes reconverge=0|estimateNoise"
