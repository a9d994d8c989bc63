set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1c_pytest.log
tail -5 gpurun_out/r1c_pytest.log
timeout 120 python tools/prof_stream.py > gpurun_out/r1c_stream.txt 2>&1
timeout 120 python tools/prof_stream.py 21e6 47e6 shuffled >> gpurun_out/r1c_stream.txt 2>&1
cat gpurun_out/r1c_stream.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_table_add_sample -s 2 -c 1 -o gpurun_out/r1c_stream_full -f python tools/prof_stream.py > gpurun_out/r1c_ncu_stream.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 1 --warmup 1 --timed-only > gpurun_out/r1c_launches.log 2>&1
ls -la gpurun_out
