set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1f_pytest.log
tail -8 gpurun_out/r1f_pytest.log
BTG_NOISE_PHASES=1 BIGS=128 timeout 600 python tools/prof_real.py 0.33 > gpurun_out/r1f_real.txt 2>&1
cat gpurun_out/r1f_real.txt
BTG_NOISE_PHASES=1 timeout 300 python tools/prof_noise.py 100000 350 2 > gpurun_out/r1f_noise.txt 2>&1
cat gpurun_out/r1f_noise.txt
