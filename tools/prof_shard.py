"""One inference unit sharded over N ranks (torchrun): lock-step estimateNoise with the in-kernel peer exchange over NVLink,
against the single-rank run of the same unit (rank 0 also runs it alone).  Prints time per iteration and whether the rates agree.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 tools/prof_shard.py [scale]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench
from bayestyper_b200 import capi, driver, engine, kmer_pipeline, shard, unit as U

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.33
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("gloo")
torch.cuda.set_device(local)
lib = capi.load()
capi.check(lib.btg_init(local), lib)
dev = torch.device("cuda", local)
opt = driver.Options(random_seed=20190401)
inp = bench.build_batch(lib, 0, scale, dev)          # every rank builds the SAME unit (rank argument 0)
inp.make_resident(lib, opt)
n_paths, mem = driver.find_variant_cluster_paths(lib, inp.graphs, inp.blooms_dev, opt)
pipe = kmer_pipeline.KmerPipeline(inp.graphs, n_paths, mem, 1, inp.genders)
pipe.enumerate_path_kmers()
pipe.scan_buffer(inp.region_buf_dev, 2, 2, False)
kd, cdv = inp.spectra_dev[0]
pipe.add_sample(0, kd, cdv)
whole = pipe.build_unit(multigroup_bloom=None)
nb = driver.estimate_nb_parameters(pipe, inp.region_buf_dev, inp.spectra_dev, inp.genders, opt)
chains = 4
gopts = lambda base: U.default_opts(seed=opt.random_seed, min_frac=U.min_fraction_observed(nb[0], nb[1]), chains=chains, group_base=base)
comm = shard.Comm.torch(world, rank)
sub, base = shard.shard(whole, world, rank)
desc, keep = shard.shard_desc(whole, comm)
cd = engine.CountDistribution(nb[0], nb[1])
eng = engine.InferenceEngine(sub)
dist.barrier()
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
    eng.estimate_noise(cd, gopts(base), want_trace=False, shard=desc)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    tt = torch.tensor([dt], dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("sharded over %d ranks: %d + ... clusters/rank, estimateNoise %.3f s, %.1f us/iteration" % (world, sub.Cn, float(tt), float(tt) / (chains * 350) * 1e6), flush=True)
rates = cd.noise_rates()
allr = [None] * world
dist.all_gather_object(allr, rates)
if rank == 0:
    assert all((r == allr[0]).all() for r in allr), "ranks disagree on the noise rates"
    cd1 = engine.CountDistribution(nb[0], nb[1])
    eng1 = engine.InferenceEngine(whole)
    torch.cuda.synchronize(); t = time.perf_counter()
    eng1.estimate_noise(cd1, gopts(0), want_trace=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("single rank: %d clusters, estimateNoise %.3f s, %.1f us/iteration" % (whole.Cn, dt, dt / (chains * 350) * 1e6))
    print("rates sharded", rates, "single", cd1.noise_rates(), "IDENTICAL" if (rates == cd1.noise_rates()).all() else "DIFFERENT", flush=True)
dist.barrier()
dist.destroy_process_group()
