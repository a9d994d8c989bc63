set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -8
timeout 600 python -m pytest tests/test_gpu_shard.py -q > gpurun_out/r1m_shard_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1m_shard_pytest.log
tail -5 gpurun_out/r1m_shard_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/prof_shard.py 0.33 > gpurun_out/r1m_shard2.txt 2>&1; echo "shard rc=$?"
grep -E "sharded|single|rates|Error|error" gpurun_out/r1m_shard2.txt | tail -8
