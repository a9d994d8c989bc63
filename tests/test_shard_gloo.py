"""The N>1 path on CPU (gloo, world_size 2): group sharding keeps every group's random streams (group_index_base),
so the concatenation of the per-rank results equals the single-process result bit for bit; the noise rates travel
by broadcast.  The per-rank compute stand-in is oracle-P (no GPU in this container)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from bayestyper_b200 import shard
    from tests import _oracle as O
    from tests._fixtures import GibbsFixture
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = GibbsFixture("gibbs_mixed_3s")
    rates = torch.zeros(fx.S, dtype=torch.float64)
    if rank == 0:
        rates[:] = torch.from_numpy(fx.tab["noise_rates"])
    dist.broadcast(rates, 0)
    sub, base = shard.shard(fx.unit, world, rank)
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    cd.set_noise_rates(rates.numpy())
    res = O.oracle_estimate_genotypes(sub, cd, fx.opts(chains=2, burn=10, samples=30, group_base=base))
    parts = [None] * world
    dist.all_gather_object(parts, {k: res[k] for k in shard.RESULT_KEYS})
    if rank == 0:
        np.savez(Path(out_dir) / "sharded.npz", **shard.concat_results(parts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    sys.path.insert(0, str(ROOT))
    from bayestyper_b200 import shard
    from tests import _oracle as O
    from tests._fixtures import GibbsFixture
    fx = GibbsFixture("gibbs_mixed_3s")
    parts = shard.partition(fx.unit, 2)
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == fx.unit.G
    cost = shard.group_costs(fx.unit)
    c0, c1 = cost[parts[0][0]:parts[0][1]].sum(), cost[parts[1][0]:parts[1][1]].sum()
    assert abs(c0 - c1) / (c0 + c1) < 0.25
    O.load()          # build the oracle library once here: two workers running `make` at the same time would race
    mp.spawn(_worker, args=(2, 29517, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    cd.set_noise_rates(fx.tab["noise_rates"])
    want = O.oracle_estimate_genotypes(fx.unit, cd, fx.opts(chains=2, burn=10, samples=30))
    for k in shard.RESULT_KEYS:
        assert (got[k] == want[k]).all(), k


def test_partition_edge_cases():
    sys.path.insert(0, str(ROOT))
    from bayestyper_b200 import shard
    from tests._fixtures import GibbsFixture
    fx = GibbsFixture("gibbs_snv_1s")
    for world in (1, 3, 8):
        p = shard.partition(fx.unit, world)
        assert p[0][0] == 0 and p[-1][1] == fx.unit.G
        assert all(p[i][1] == p[i + 1][0] for i in range(world - 1))
    empty = fx.unit.subset_groups(np.zeros(0, np.int64))
    assert shard.partition(empty, 4) == [(0, 0)] * 4


def test_group_tables_describe_the_whole_unit():
    """What every rank hands to the lock-step entry points (btg_shard_desc): clusters and variants per group."""
    sys.path.insert(0, str(ROOT))
    from bayestyper_b200 import shard
    from tests._fixtures import GibbsFixture
    for name in ("gibbs_mixed_3s", "gibbs_nested_2s"):
        fx = GibbsFixture(name)
        n_cl, n_var = shard.group_tables(fx.unit)
        a = fx.unit.a
        gco, cvo = a["group_cluster_off"].astype(np.int64), a["cl_var_off"].astype(np.int64)
        assert len(n_cl) == fx.unit.G and n_cl.sum() == fx.unit.Cn and n_var.sum() == cvo[-1]
        for g in (0, fx.unit.G // 2, fx.unit.G - 1):
            assert n_cl[g] == gco[g + 1] - gco[g]
            assert n_var[g] == cvo[gco[g + 1]] - cvo[gco[g]]
        # the shards tile the group axis, so group_index_base + local index addresses these tables
        parts = shard.partition(fx.unit, 3)
        assert sum(hi - lo for lo, hi in parts) == fx.unit.G


# ---- the k-mer path's one exchange: best paths of every cluster (driver.Shard / subset_graphs_for_paths / merge_best_paths) ----------

def _paths_worker(rank, world, port, out_dir, name):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from bayestyper_b200 import btd, driver
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    sh = driver.Shard(world, rank, allgather)
    d = btd.read(ROOT / "tests" / "golden" / f"{name}.btd")
    g = {k[2:]: v for k, v in d.items() if k.startswith("g.")}
    G = len(g["group_cluster_off"]) - 1
    mine = sh.my_groups(G)
    sub, clusters = driver.subset_graphs_for_paths(g, mine)
    # stand-in for this rank's path search (no GPU here): the golden best paths of its own clusters, in the sub-graph's order
    V = np.diff(g["cl_vertex_off"].astype(np.int64))
    po = g["cl_path_off"].astype(np.int64)
    n_paths = (np.diff(po) // np.maximum(V, 1))[clusters]
    mem = np.concatenate([g["path_bits"][po[c]:po[c + 1]] for c in clusters]) if len(clusters) else np.zeros(0, np.uint8)
    assert (np.diff(sub["cl_vertex_off"].astype(np.int64)) == V[clusters]).all()
    parts = sh.allgather((n_paths, mem, clusters))
    merged_n, merged_mem = driver.merge_best_paths(g, [(p[0], p[1]) for p in parts], [p[2] for p in parts])
    np.savez(Path(out_dir) / f"merged_{rank}.npz", n=merged_n, mem=merged_mem)
    dist.barrier()
    dist.destroy_process_group()


def test_best_paths_exchange_two_ranks(tmp_path):
    """Every rank searches the paths of its own (strided) groups; after the all-gather every rank holds the best paths of the
    whole unit, identical to the single-rank result (here: the reference's golden best paths)."""
    sys.path.insert(0, str(ROOT))
    from bayestyper_b200 import btd
    name = "paths_nested_2s"
    mp.spawn(_paths_worker, args=(2, 29531, str(tmp_path), name), nprocs=2, join=True)
    d = btd.read(ROOT / "tests" / "golden" / f"{name}.btd")
    V = np.diff(d["g.cl_vertex_off"].astype(np.int64))
    want_n = np.diff(d["g.cl_path_off"].astype(np.int64)) // np.maximum(V, 1)
    for r in range(2):
        got = np.load(tmp_path / f"merged_{r}.npz")
        assert (got["n"] == want_n).all()
        assert (got["mem"] == d["g.path_bits"]).all()


def test_subset_graphs_keeps_vertices_edges_and_group_indices():
    sys.path.insert(0, str(ROOT))
    from bayestyper_b200 import btd, driver
    d = btd.read(ROOT / "tests" / "golden" / "paths_nested_2s.btd")
    g = {k[2:]: v for k, v in d.items() if k.startswith("g.")}
    G = len(g["group_cluster_off"]) - 1
    groups = np.arange(1, G, 3, dtype=np.int64)
    sub, clusters = driver.subset_graphs_for_paths(g, groups)
    gco = g["group_cluster_off"].astype(np.int64)
    assert (clusters == np.concatenate([np.arange(gco[x], gco[x + 1]) for x in groups])).all()
    assert (sub["cl_group_global"] == np.repeat(groups, np.diff(gco)[groups])).all()      # the seed of a cluster's path search is its group's index in the WHOLE unit
    assert (sub["cluster_idx"] == g["cluster_idx"][clusters]).all()
    cvo, vso, vio = (g[k].astype(np.int64) for k in ("cl_vertex_off", "v_seq_off", "v_in_off"))
    s_cvo, s_vso, s_vio = (sub[k].astype(np.int64) for k in ("cl_vertex_off", "v_seq_off", "v_in_off"))
    for i, c in enumerate(clusters):
        assert s_cvo[i + 1] - s_cvo[i] == cvo[c + 1] - cvo[c]
        for j in range(cvo[c + 1] - cvo[c]):
            v, sv = cvo[c] + j, s_cvo[i] + j
            assert (sub["seq"][s_vso[sv]:s_vso[sv + 1]] == g["seq"][vso[v]:vso[v + 1]]).all()
            assert sub["v_flags"][sv] == g["v_flags"][v]
            assert (sub["v_in_src"][s_vio[sv]:s_vio[sv + 1]] == g["v_in_src"][vio[v]:vio[v + 1]]).all()     # edge sources are vertex indices inside the cluster
    empty, cl = driver.subset_graphs_for_paths(g, np.zeros(0, np.int64))
    assert len(cl) == 0 and len(empty["cl_vertex_off"]) == 1


def _chains_worker(rank, world, port, out_dir):
    """estimateNoise split by CHAINS: this rank runs the chains rank, rank + world, ... of the WHOLE unit — here the
    oracle stands in for btg_estimate_noise_chains (its chains are the same independent streams: the rows of the other ranks' chains are
    simply not used) — and returns the [n_chains][S] post-burn-in rate sums of its own chains, zeros elsewhere; the exchange and the finish
    are the host code of driver.genotype."""
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from tests import _oracle as O
    from tests._fixtures import GibbsFixture
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = GibbsFixture("gibbs_snv_1s")
    CH, BURN, SAMPLES = 5, 6, 9
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    trace = O.oracle_estimate_noise(fx.unit, cd, fx.opts(chains=CH, burn=BURN, samples=SAMPLES))
    rows = trace[:-1].reshape(CH, BURN + SAMPLES + 1, 2 + fx.S)
    sums = np.zeros((CH, fx.S))
    for c in range(rank, CH, world):
        sums[c] = rows[c, 1 + BURN:, 2:].sum(axis=0)
    parts = [None] * world
    dist.all_gather_object(parts, sums)
    total = np.zeros_like(sums)
    for p in parts:                      # rank order; every row is non-zero on exactly one rank
        total += p
    assert all((np.count_nonzero([p[c].any() for p in parts]) == 1) for c in range(CH))
    rates = total.sum(axis=0) / (CH * SAMPLES)          # btg_count_dist_finish_noise
    np.save(Path(out_dir) / f"rates_{rank}.npy", rates)
    dist.barrier()
    dist.destroy_process_group()


def test_noise_chain_sums_exchange_two_ranks(tmp_path):
    """Both ranks end with the noise rates of the single-process run, and with the SAME bits as each other (the sums are added in rank order on
    every rank)."""
    sys.path.insert(0, str(ROOT))
    from tests import _oracle as O
    from tests._fixtures import GibbsFixture
    O.load()
    mp.spawn(_chains_worker, args=(2, 29523, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rates_0.npy"), np.load(tmp_path / "rates_1.npy")
    assert (r0 == r1).all()
    fx = GibbsFixture("gibbs_snv_1s")
    cd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    O.oracle_estimate_noise(fx.unit, cd, fx.opts(chains=5, burn=6, samples=9))
    assert np.abs(r0 / cd.noise_rates() - 1).max() < 1e-12
