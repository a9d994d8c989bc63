"""Timing driver on the REAL pipeline's unit (bench.py's batch at a given scale): estimateNoise with several big-cluster
thresholds, estimateGenotypes with / without per-iteration reconvergence.  python tools/prof_real.py [scale]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from bayestyper_b200 import capi, driver, engine, kmer_pipeline, unit as U

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.33
lib = capi.load()
capi.check(lib.btg_init(0), lib)
dev = torch.device("cuda", 0)
opt = driver.Options(random_seed=20190401)
inp = bench.build_batch(lib, 0, scale, dev)
inp.make_resident(lib, opt)
n_paths, mem = driver.find_variant_cluster_paths(lib, inp.graphs, inp.blooms_dev, opt)
pipe = kmer_pipeline.KmerPipeline(inp.graphs, n_paths, mem, 1, inp.genders)
pipe.enumerate_path_kmers()
pipe.scan_buffer(inp.region_buf_dev, 2, 2, False)
kd, cdv = inp.spectra_dev[0]
pipe.add_sample(0, kd, cdv)
unit = pipe.build_unit(multigroup_bloom=None)
H = unit.a["cl_nhap"]; K = np.diff(unit.a["cl_kmer_off"].astype(np.int64)); nu = np.diff(unit.a["cl_uniq_off"].astype(np.int64))
cost = (H.astype(np.int64) * (H + 1) // 2) * (nu // 10 + 1)
print("clusters", unit.Cn, "variants", unit.n_variants, "H hist", np.bincount(H)[:12], "K mean", K.mean(), "uniq mean", nu.mean())
print("fill cost quantiles", np.quantile(cost, [0.5, 0.9, 0.99, 0.999, 1.0]), "frac > 64/128/384:", (cost > 64).mean(), (cost > 128).mean(), (cost > 384).mean())
nb = driver.estimate_nb_parameters(pipe, inp.region_buf_dev, inp.spectra_dev, inp.genders, opt)
cd = engine.CountDistribution(nb[0], nb[1])
eng = engine.InferenceEngine(unit)
gopts = U.default_opts(seed=opt.random_seed, min_frac=U.min_fraction_observed(nb[0], nb[1]), chains=4)
for big in os.environ.get("BIGS", "384,128,64,32").split(","):
    os.environ["BTG_NOISE_BIG"] = big
    torch.cuda.synchronize(); t = time.perf_counter()
    eng.estimate_noise(cd, gopts, want_trace=False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("estimateNoise big>%s: %.3f s, %.1f us/iteration, rates %s" % (big, dt, dt / (4 * 350) * 1e6, cd.noise_rates()), flush=True)
gopts = U.default_opts(seed=opt.random_seed, min_frac=U.min_fraction_observed(nb[0], nb[1]))
for sync in ("0", "1", "0", "1"):
    os.environ["BTG_GIBBS_SYNC"] = sync
    torch.cuda.synchronize(); t = time.perf_counter()
    r = eng.estimate_genotypes(cd, gopts)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("estimateGenotypes reconverge=%s: %.3f s  %.0f clusters/s  gpp checksum %.6f" % (sync, dt, unit.Cn / dt, float(r["gpp"].astype(np.float64).sum())), flush=True)
