set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r1b_gpu.txt
nproc >> gpurun_out/r1b_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1b_pytest.log
tail -5 gpurun_out/r1b_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"
timeout 120 python tools/prof_stream.py > gpurun_out/r1b_stream.txt 2>&1
timeout 120 python tools/prof_stream.py 21e6 47e6 shuffled >> gpurun_out/r1b_stream.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_table_add_sample -s 2 -c 1 -o gpurun_out/r1b_stream_full -f python tools/prof_stream.py > gpurun_out/r1b_ncu_stream.log 2>&1
timeout 300 python tools/prof_noise.py 100000 350 2 > gpurun_out/r1b_noise.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_noise_chain -c 1 -o gpurun_out/r1b_noise_full -f python tools/prof_noise.py 100000 40 1 > gpurun_out/r1b_ncu_noise.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_estimate_genotypes -c 1 -o gpurun_out/r1b_gibbs_full -f python tools/prof_gibbs.py 30000 1 > gpurun_out/r1b_ncu_gibbs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 1 --warmup 1 --timed-only > gpurun_out/r1b_launches.log 2>&1
ls -la gpurun_out
