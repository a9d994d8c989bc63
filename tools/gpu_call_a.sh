# one GPU call of the end of round 2: the batched path search first (short timeout), then the whole GPU suite, smoke, the bench line, configs[3] shape both ways
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_paths.py -q > gpurun_out/pytest_paths.log 2>&1; RC=$?; tail -6 gpurun_out/pytest_paths.log; echo "paths rc=$RC"
timeout 400 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_paths.py > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; head -c 300 gpurun_out/bench_n1.json; echo
BTG_STAGE_TIMES=1 BTG_PATHS_BATCH=0 timeout 150 python bench.py --config D --timed-only --steps 1 --warmup 1 > gpurun_out/bench_D_seq.json 2> gpurun_out/bench_D_seq.err; tail -c 700 gpurun_out/bench_D_seq.err; cat gpurun_out/bench_D_seq.json
if [ $RC -eq 0 ]; then
BTG_PATHS_BATCH=1 timeout 200 python bench.py --config D --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_D_batch.json 2> gpurun_out/bench_D_batch.err; tail -c 300 gpurun_out/bench_D_batch.err; head -c 300 gpurun_out/bench_D_batch.json; echo
timeout 70 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_paths.py -q -k "mixed_3s and batch and not first" > gpurun_out/r2_sanitizer_memcheck_paths_batch.log 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck_paths_batch.log
fi
