for big in 0 64 384; do
  BTG_NOISE_BIG=$big BIGS=$big BTG_NOISE_PHASES=1 timeout 300 python tools/prof_real.py ${1:-0.04} 2>&1 | grep -E "^clusters|phases|estimateNoise big" | tail -3
done
