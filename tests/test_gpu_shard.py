"""Sharded lock-step modes on the device: two ranks = two processes, each holding half of the unit's
groups, add up the per-sample noise statistics once per iteration through peer mailboxes inside the chain kernel
(csrc/comm.cuh).  Noise trace, final rates and (joint mode) every genotype field must equal the single-rank run bit
for bit.  On a one-GPU box both ranks run on cuda:0 (the driver time-slices the two persistent kernels; the exchange
is the same CUDA-IPC mapping as across NVLink), with more GPUs visible each rank takes its own."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

from bayestyper_b200 import engine
from tests._fixtures import GibbsFixture

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CH, BURN, SAMPLES = 2, 6, 10


def _worker(rank, world, port, out_dir, name, joint):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from bayestyper_b200 import capi, engine as E, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("BTG_PEER_TIMEOUT_MS", "60000")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    lib = capi.load()
    capi.check(lib.btg_init(dev), lib)
    fx = GibbsFixture(name)
    comm = shard.Comm.torch(world, rank)
    sub, base = shard.shard(fx.unit, world, rank)
    desc, keep = shard.shard_desc(fx.unit, comm)
    cd = E.CountDistribution(fx.nb_p, fx.nb_size)
    opts = fx.opts(chains=CH, burn=BURN, samples=SAMPLES, group_base=base)
    eng = E.InferenceEngine(sub)
    out = {}
    if joint:
        res, trace = eng.estimate_noise_and_genotypes(cd, opts, shard=desc)
        out.update({k: res[k] for k in shard.RESULT_KEYS})
    else:
        trace = eng.estimate_noise(cd, opts, shard=desc)
    out["trace"] = trace
    out["rates"] = cd.noise_rates()
    parts = [None] * world
    dist.all_gather_object(parts, out)
    if rank == 0:
        np.savez(Path(out_dir) / "ranks.npz", **{f"{k}_{r}": v for r, p in enumerate(parts) for k, v in p.items()})
    dist.barrier()
    eng.close(); cd.close(); comm.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name,joint", [("gibbs_mixed_3s", False), ("gibbs_chrx_2s", True)])
def test_two_ranks_equal_one(btg, tmp_path, name, joint):
    from bayestyper_b200 import shard
    fx = GibbsFixture(name)
    cd = engine.CountDistribution(fx.nb_p, fx.nb_size)
    opts = fx.opts(chains=CH, burn=BURN, samples=SAMPLES)
    eng = engine.InferenceEngine(fx.unit)
    if joint:
        want, wtrace = eng.estimate_noise_and_genotypes(cd, opts)
    else:
        want, wtrace = None, eng.estimate_noise(cd, opts)
    wrates = cd.noise_rates()
    # a shard descriptor without a communicator (one rank holding everything) is the same computation
    cd1 = engine.CountDistribution(fx.nb_p, fx.nb_size)
    desc, keep = shard.shard_desc(fx.unit, None)
    t1 = eng.estimate_noise_and_genotypes(cd1, opts, shard=desc)[1] if joint else eng.estimate_noise(cd1, opts, shard=desc)
    assert (t1 == wtrace).all()
    eng.close(); cd.close(); cd1.close()
    mp.spawn(_worker, args=(2, 29533 + int(joint), str(tmp_path), name, joint), nprocs=2, join=True)
    got = np.load(tmp_path / "ranks.npz")
    for r in range(2):
        assert (got[f"trace_{r}"] == wtrace).all(), f"rank {r}: noise trace differs from the single-rank run"
        assert (got[f"rates_{r}"] == wrates).all()
    if joint:
        for k in shard.RESULT_KEYS:
            assert (np.concatenate([got[f"{k}_0"], got[f"{k}_1"]]) == want[k]).all(), k


# ---- the whole step of a sharded unit (driver.genotype with a Shard): what bench.py --gpus N runs ----------------------------------

def _small_batch(lib, dev):
    import bench
    return bench.build_batch(lib, 0, 0.01, dev)          # every rank builds the SAME 3,000-variant batch


def _driver_worker(rank, world, port, out_dir, noise_split):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from bayestyper_b200 import capi, driver, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("BTG_PEER_TIMEOUT_MS", "60000")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    lib = capi.load()
    capi.check(lib.btg_init(dev), lib)
    opt = driver.Options(random_seed=77, n_chains=3, gibbs_burn_in=8, gibbs_samples=12, noise_split=noise_split)
    inp = _small_batch(lib, torch.device("cuda", dev))
    inp.make_resident(lib, opt)

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    ctx = driver.Shard(world, rank, allgather, shard.Comm.torch(world, rank))
    _, _, res, info = driver.genotype(inp, opt, resident=True, shard=ctx)
    G = len(inp.graphs["group_cluster_off"]) - 1
    parts = allgather({"res": {k: res[k] for k in shard.RESULT_KEYS}, "groups": ctx.my_groups(G), "rates": info["noise_rates"], "n": info["n_clusters"]})
    if rank == 0:
        import pickle
        with open(Path(out_dir) / "parts.pkl", "wb") as f:
            pickle.dump(parts, f)
    dist.barrier()
    inp.free(lib)
    ctx.comm.close()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("noise_split", ["chains", "groups"])
def test_sharded_step_equals_the_single_rank_step(btg, tmp_path, noise_split):
    """Two ranks share ONE unit (strided groups): own path search + best-path all-gather, estimateNoise split by chains (no exchange while the
    chains run) or by groups (in-kernel mailbox exchange), estimateGenotypes on the own groups.  Noise rates and every result row must equal
    the single-rank step's, bit for bit."""
    import pickle
    import torch
    from bayestyper_b200 import capi, driver, shard
    lib = capi.load()
    opt = driver.Options(random_seed=77, n_chains=3, gibbs_burn_in=8, gibbs_samples=12)
    inp = _small_batch(lib, torch.device("cuda", 0))
    inp.make_resident(lib, opt)
    graphs, _, want, winfo = driver.genotype(inp, opt, resident=True)
    mp.spawn(_driver_worker, args=(2, 29541 + (noise_split == "groups"), str(tmp_path), noise_split), nprocs=2, join=True)
    with open(tmp_path / "parts.pkl", "rb") as f:
        parts = pickle.load(f)
    gco = np.asarray(graphs["group_cluster_off"], np.int64)
    cvo = np.asarray(graphs["cl_var_off"], np.int64)
    assert sum(p["n"] for p in parts) == winfo["n_clusters"]
    nv, S = int(cvo[-1]), 1
    level_off = {"v2": np.arange(nv + 1) * 2 * S, "v1": np.arange(nv + 1) * S, "v": np.arange(nv + 1),
                 "geno": np.asarray(want["geno_off"], np.int64), "allele": np.asarray(want["allele_off"], np.int64), "valt": np.asarray(want["valt_off"], np.int64)}
    level = {"gt": "v2", "gq": "v1", "ploidy": "v1", "gpp": "geno", "app": "allele", "nak": "allele", "fak": "allele", "mac": "allele", "saf": "allele",
             "an": "v", "hc": "v", "ac": "valt", "af": "valt", "acp": "valt", "anc": "valt"}
    for p in parts:
        assert (p["rates"] == winfo["noise_rates"]).all()
        clusters = np.concatenate([np.arange(gco[g], gco[g + 1]) for g in p["groups"]])
        variants = np.concatenate([np.arange(cvo[c], cvo[c + 1]) for c in clusters])
        for k in shard.RESULT_KEYS:
            _, rows = driver._take_csr(level_off[level[k]], variants)
            got = np.asarray(p["res"][k])
            assert len(got) == len(rows), k
            assert (got == np.asarray(want[k])[rows]).all(), k
    inp.free(lib)
