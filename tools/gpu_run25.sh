timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
BTG_NOISE_PHASES=1 BIGS=128 timeout 300 python tools/prof_real.py 0.33 2>&1 | grep -E "estimateNoise|phases"
