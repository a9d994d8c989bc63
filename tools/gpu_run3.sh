set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest.log
tail -8 gpurun_out/r1d_pytest.log
timeout 120 python tools/prof_stream.py > gpurun_out/r1d_stream.txt 2>&1
timeout 120 python tools/prof_stream.py 21e6 47e6 shuffled >> gpurun_out/r1d_stream.txt 2>&1
cat gpurun_out/r1d_stream.txt
timeout 600 python tools/prof_real.py 0.33 > gpurun_out/r1d_real.txt 2>&1
cat gpurun_out/r1d_real.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_table_add_sample -s 2 -c 1 -o gpurun_out/r1d_stream_full -f python tools/prof_stream.py > gpurun_out/r1d_ncu_stream.log 2>&1
ls -la gpurun_out | grep r1d
