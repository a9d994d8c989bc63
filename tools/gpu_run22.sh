timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
for kc in 1 4 8 16; do
echo "BTG_NOISE_CONCURRENCY=$kc"
BTG_NOISE_CONCURRENCY=$kc BIGS=128 timeout 300 python tools/prof_real.py 0.33 2>&1 | grep -E "estimateNoise"
done
