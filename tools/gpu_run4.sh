set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r1e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest.log
tail -8 gpurun_out/r1e_pytest.log
timeout 120 python tools/prof_stream.py > gpurun_out/r1e_stream.txt 2>&1
timeout 120 python tools/prof_stream.py 21e6 47e6 shuffled >> gpurun_out/r1e_stream.txt 2>&1
cat gpurun_out/r1e_stream.txt
BIGS=384,128,64 timeout 600 python tools/prof_real.py 0.33 > gpurun_out/r1e_real.txt 2>&1
cat gpurun_out/r1e_real.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_table_add_sample -s 2 -c 1 -o gpurun_out/r1e_stream_full -f python tools/prof_stream.py > gpurun_out/r1e_ncu_stream.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_noise_chain -c 1 -o gpurun_out/r1e_noise_full -f python tools/prof_noise.py 100000 40 1 > gpurun_out/r1e_ncu_noise.log 2>&1
ls -la gpurun_out | grep r1e
