mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1r_bench_n2.json 2> gpurun_out/r1r_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r1r_bench_n2.json; tail -3 gpurun_out/r1r_bench_n2.err
