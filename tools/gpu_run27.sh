python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
BTG_NOISE_PHASES=1 timeout 300 python tools/prof_real.py 1.0 2>&1 | grep -E "clusters|phases|estimateNoise|estimateGenotypes reconverge=0" | head -6
