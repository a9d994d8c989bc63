set -x
BTG_NOISE_PHASES=1 BIGS=128 timeout 600 python tools/prof_real.py 0.33 2>&1 | grep -v estimateGenotypes
BTG_NOISE_PHASES=1 timeout 300 python tools/prof_noise.py 100000 350 1 2>&1 | tail -6
