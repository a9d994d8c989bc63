mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r1q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1q_pytest.log
tail -3 gpurun_out/r1q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err; echo "bench rc=$?"
