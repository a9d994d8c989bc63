mkdir -p gpurun_out; timeout 170 python bench.py --steps 3 --warmup 3 > gpurun_out/r1s_bench.json 2> gpurun_out/r1s_bench.err; echo "bench rc=$?"
