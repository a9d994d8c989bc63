mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r1o_bench.err
