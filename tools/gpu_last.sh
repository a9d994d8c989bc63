timeout 300 python -m pytest tests/test_gpu_gibbs.py tests/test_gpu_shard.py tests/test_host_cpp.py -m gpu -q -x 2>&1 | tail -3
BTG_NOISE_PHASES=1 BIGS=128 timeout 200 python tools/prof_real.py 0.33 2>&1 | grep -E "phases|estimateNoise"
