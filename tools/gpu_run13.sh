set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r1l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1l_pytest.log
tail -3 gpurun_out/r1l_pytest.log
BTG_UPLOAD_TIMING=1 BTG_NOISE_PHASES=1 BIGS=128 timeout 600 python tools/prof_real.py 0.33 2>&1 | grep -v "reconverge=1" | tee gpurun_out/r1l_real.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; echo "bench rc=$?"
