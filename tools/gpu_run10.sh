set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r1j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j_pytest.log
tail -4 gpurun_out/r1j_pytest.log
BTG_NOISE_PHASES=1 BIGS=128 timeout 600 python tools/prof_real.py 0.33 2>&1 | grep -v "reconverge=1" | tee gpurun_out/r1j_real.txt
BTG_NOISE_PHASES=1 timeout 300 python tools/prof_noise.py 100000 350 1 2>&1 | tail -6 | tee gpurun_out/r1j_noise.txt
