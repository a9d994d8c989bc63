"""Smoke checks for the later pipeline stages: one small Gibbs invocation (default mode + estimateNoise) on a committed fixture,
checked against oracle-P, and the sample-stream kernel against a numpy lookup."""
import numpy as np


def run(lib) -> None:
    import torch
    from . import capi, engine, synth
    from tests import _oracle as O
    from tests._fixtures import GibbsFixture
    # Gibbs path: identical diplotype tallies / noise trace vs the CPU restatement consuming the same streams
    fx = GibbsFixture("gibbs_snv_1s")
    opts = fx.opts(chains=3, burn=10, samples=20)
    ocd = O.OracleCountDist(fx.nb_p, fx.nb_size)
    gcd = engine.CountDistribution(fx.nb_p, fx.nb_size)
    eng = engine.InferenceEngine(fx.unit)
    otrace = O.oracle_estimate_noise(fx.unit, ocd, opts)
    gtrace = eng.estimate_noise(gcd, opts)
    assert np.abs(gtrace[:, 2:] - otrace[:, 2:]).max() <= 1e-9 * np.abs(otrace[:, 2:]).max(), "noise trace mismatch"
    gcd.set_noise_rates(ocd.noise_rates())      # the same rates to the last bit on both sides before genotyping
    ores, otally = O.oracle_estimate_genotypes(fx.unit, ocd, opts, want_tally=True)
    gres = eng.estimate_genotypes(gcd, opts)
    toff = fx.unit.tally_offsets()
    for c in range(fx.unit.Cn):
        assert (eng.cluster_tally(c).reshape(-1) == otally[int(toff[c]):int(toff[c + 1])]).all(), f"tallies of cluster {c} differ"
    assert np.abs(gres["gpp"] - ores["gpp"]).max() <= 1e-4 and (gres["gt"] == ores["gt"]).all(), "posterior mismatch"
    eng.close(); gcd.close()
    # sample stream -> exact table (KMC-ordered records)
    rng = np.random.default_rng(5)
    km = rng.integers(0, 2**64, size=(4000, 2), dtype=np.uint64)
    km[:, 1] &= np.uint64((1 << 46) - 1)
    km = km[synth.kmc_order(km)]
    kd = torch.from_numpy(km.view(np.int64)).cuda()
    lo = torch.empty(len(km), dtype=torch.int64, device="cuda"); hi = torch.empty_like(lo)
    torch.cuda.synchronize()
    capi.check(lib.btg_table_keys_from_kmers_dev(kd.data_ptr(), len(km), lo.data_ptr(), hi.data_ptr(), None), lib)
    recs = kd[::3].contiguous()
    cts = torch.full((recs.shape[0],), 7, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(len(km), dtype=torch.uint8, device="cuda"); has = torch.zeros(len(km), dtype=torch.uint8, device="cuda")
    capi.check(lib.btg_table_set_index_dev(None, 0), lib)
    torch.cuda.synchronize()
    capi.check(lib.btg_table_add_sample_kmers_dev(lo.data_ptr(), hi.data_ptr(), len(km), recs.data_ptr(), cts.data_ptr(), recs.shape[0], 1, 0, counts.data_ptr(),
                                                  has.data_ptr(), None), lib)
    capi.check(lib.btg_kmer_hash(capi.ptr(np.zeros((1, 2), np.uint64)), 1, capi.ptr(np.zeros(1, np.uint64))), lib)   # library-stream sync
    exp = np.zeros(len(km), np.uint8); exp[::3] = 7
    assert (counts.cpu().numpy() == exp).all() and (has.cpu().numpy() == (exp > 0)).all(), "sample stream mismatch"
