set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1p_pytest.log
tail -4 gpurun_out/r1p_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1p_bench.json 2> gpurun_out/r1p_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1p_bench_reference.json 2> gpurun_out/r1p_bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1p_launches.csv python bench.py --steps 1 --warmup 1 --timed-only > gpurun_out/r1p_launches.log 2>&1
