# What the round's GPU evidence was produced with (run under gpurun; outputs land in gpurun_out/):
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --timed-only
ncu --set full --clock-control none --import-source on -k regex:k_table_add_sample -s 2 -c 1 -o gpurun_out/stream_full -f python tools/prof_stream.py
BTG_NOISE_PHASES=1 BTG_GIBBS_TIMING=1 python tools/prof_real.py 0.33
# 2 GPUs: gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_shard.py -q; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/prof_shard.py 0.33'
# staged in round 1 without a GPU: run these first in round 2
#   python tools/e2e_check.py e2e_nested_2s > gpurun_out/e2e_nested.log 2>&1      # nested candidate set end to end
#   python tools/e2e_check.py genome > gpurun_out/e2e_genome.log 2>&1             # several contigs + decoy + haploid chrX (driver_genome)
#   host/btkmc makebloom <kmc prefix> 0.001                                       # makeBloom on the device
