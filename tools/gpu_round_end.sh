# What the round's GPU evidence was produced with (run under gpurun; outputs land in gpurun_out/):
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json
# launch list of the bench's own step (cold-cache, serialised: shares, not absolutes) and the stream kernel's DRAM traffic for roofline.traffic
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --timed-only
ncu --set full --clock-control none --import-source on -k regex:k_table_add_sample -c 1 -o gpurun_out/r2_stream_bench -f python bench.py --timed-only --steps 1 --warmup 1
ncu --set full --clock-control none --import-source on -k regex:k_find_sample_paths -c 1 -o gpurun_out/r2_paths_bench -f python bench.py --timed-only --steps 1 --warmup 1 --scale 0.25
# sanitizers: the Bloom atomicOr inserts, the hand-rolled grid barrier of the chain kernels, the warp-cooperative path search, the 2-rank mailbox exchange
compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_kmer.py" -q -k "bloom" > gpurun_out/r2_sanitizer_memcheck_bloom.log 2>&1
compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_paths.py" -q -k "snv" > gpurun_out/r2_sanitizer_racecheck_paths.log 2>&1
compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_gibbs.py::test_estimate_noise_matches_oracle" "tests/test_gpu_gibbs_wide.py::test_joint_mode_wide" -q > gpurun_out/r2_sanitizer_memcheck_chain.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_shard.py -q > gpurun_out/r2_sanitizer_memcheck_shard.log 2>&1
# scaling: gpurun --gpus 8 -- 'bash tools/scale_run.sh 8 B --steps 3 --warmup 2; bash tools/scale_run.sh 8 D --steps 2 --warmup 1'
#          (profiles/r2_scale_sharded_*: group split; profiles/r2_scale_chains_B_n{2,4}.json: gpurun --gpus N -- 'bash tools/scale_run.sh N B --steps 2 --warmup 1')
# the last calls of round 2: tools/gpu_call_a.sh (batched path search first, whole suite, smoke, bench line, configs[3] shape both ways, memcheck)
