for big in 256 384 512 768 1024; do
  BTG_NOISE_BIG=$big BIGS=$big BTG_NOISE_PHASES=1 timeout 300 python tools/prof_real.py 1.0 2>&1 | grep -E "phases|estimateNoise big" | tail -2
done
