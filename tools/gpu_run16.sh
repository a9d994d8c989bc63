set -x
timeout 300 python tools/prof_gibbs.py 30000 8 2>&1 | tail -2
BTG_LIB=$PWD/bayestyper_b200/lib/libbtgpu_outline.so timeout 300 python tools/prof_gibbs.py 30000 8 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_gibbs.py -q -x 2>&1 | tail -3
