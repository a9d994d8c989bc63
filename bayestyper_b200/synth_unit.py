"""Synthetic inference units of the shapes BASELINE.json names, built straight from a
Workload's truth (variants + sample genotypes) — the bench's "synthetic batch".

What the k-mer stages (path search -> classify -> getHaplotypeCandidates) hand to the
Gibbs sampler is, per cluster, a (k-mers x haplotype-candidates) multiplicity matrix,
per-sample k-mer counts and the per-variant coverage bitmaps.  For whole-genome /
30-sample shapes the k-mer sets cannot be materialised (3 x 45 GB per sample,
SURVEY.md §8d rows C/D), so this generator derives those descriptors directly:

  candidates  = the distinct allele combinations carried by the samples' haplotypes in
                the cluster (+ the all-reference combination with probability p_ref)
  rows        = distinct canonical 55-mers of the candidates over the cluster span
  counts      = NB(mean*m, var*m) for the sample's true diplotype multiplicity m > 0,
                Poisson(noise) for m = 0, saturated at 255
  coverage    = a k-mer covers variant v on candidate h when its window overlaps v's allele

It is input synthesis (like drawing a random token batch), not an implementation of the
reference's k-mer stages.
"""
from __future__ import annotations

import numpy as np

from . import synth
from .unit import Unit

K = synth.K


def _aligned_kmers(seq: bytes):
    """canonical k-mer starting at every position (n = len-K+1 rows) + validity mask."""
    c = synth._CODE[np.frombuffer(seq, np.uint8)]
    L = len(c)
    n = L - K + 1
    bad = np.concatenate([[0], np.cumsum(c > 3)])
    valid = (bad[K:] - bad[:-K]) == 0
    clean = np.where(c > 3, 0, c).astype(np.uint8)
    km = synth.canonical_kmers(synth._ACGT[clean].tobytes())
    assert len(km) == n
    return km, valid


def clusters_of(variants, k: int = K):
    """Index ranges [i0, i1) of variants that fall in one cluster (gap to the running end < k)."""
    out = []
    i0 = 0
    end = -10**18
    for i, v in enumerate(variants):
        if i > i0 and v.pos - end >= k:
            out.append((i0, i))
            i0 = i
            end = -10**18
        end = max(end, v.pos + len(v.ref) - 1)
    if variants:
        out.append((i0, len(variants)))
    return out


def build_unit(w: synth.Workload, seed: int = 5, noise_rate: float = 0.02, p_ref: float = 0.3, max_cluster_variants: int = 6,
               max_clusters: int | None = None) -> Unit:
    rng = np.random.default_rng(seed)
    S = w.genotypes.shape[0]
    variants = w.variants
    nv = len(variants)
    pos = np.array([v.pos for v in variants], np.int64)
    reflen = np.array([len(v.ref) for v in variants], np.int64)
    altlen = np.array([len(v.alts[0]) for v in variants], np.int64)
    # haplotypes: index 0 = reference, then sample haplotypes
    hap_alleles = [np.zeros(nv, np.uint8)]
    hap_owner = []
    for s in range(S):
        n_hap = 1 if (w.genders[s] == "M" and w.chrom.lower() in ("x", "chrx")) else 2
        for h in range(n_hap):
            hap_alleles.append(w.genotypes[s, :, h].astype(np.uint8))
            hap_owner.append(s)
    seqs = [w.reference] + [synth.apply_variants(w.reference, variants, a) for a in hap_alleles[1:]]
    kms = [_aligned_kmers(q) for q in seqs]
    # shift of haplotype coordinates after each variant
    shifts = [np.concatenate([[0], np.cumsum(np.where(a > 0, altlen - reflen, 0))]) for a in hap_alleles]
    sample_haps = [[i + 1 for i, o in enumerate(hap_owner) if o == s] for s in range(S)]

    cl_ranges = clusters_of(variants)
    # split over-long clusters so that the candidate count stays bounded (the reference caps at 32 per sample)
    ranges = []
    for i0, i1 in cl_ranges:
        while i1 - i0 > max_cluster_variants:
            ranges.append((i0, i0 + max_cluster_variants)); i0 += max_cluster_variants
        ranges.append((i0, i1))
    if max_clusters is not None:
        ranges = ranges[:max_clusters]

    A = {k: [] for k in ("cl_nhap", "mult", "k_has_counts", "k_counts", "k_ic", "uniq_idx", "vh_var", "vh_bits", "hap_alleles", "var_nalleles")}
    cl_kmer_off, cl_var_off, cl_mult_off, cl_uniq_off, cl_hapvar_off = [0], [0], [0], [0], [0]
    kmer_vh_cnt, vh_bits_len = [], []
    p_nb = w.depth_mean / w.depth_var
    size_nb = w.depth_mean ** 2 / (w.depth_var - w.depth_mean)
    view_dt = np.dtype([("a", np.uint64), ("b", np.uint64)])
    for (i0, i1) in ranges:
        n = i1 - i0
        lo_ref = int(pos[i0]) - (K - 1)
        hi_ref = int((pos[i0:i1] + reflen[i0:i1] - 1).max()) + (K - 1)
        combos = {}
        for hi_, a in enumerate(hap_alleles):
            key = a[i0:i1].tobytes()
            if key not in combos and (hi_ > 0 or rng.random() < p_ref):
                combos[key] = hi_
        cand = list(combos.values())
        H = len(cand)
        rows_k, rows_c, rows_w = [], [], []
        cover = []                       # per candidate: (n_windows, n) bool
        for ci, hi_ in enumerate(cand):
            sh = shifts[hi_]
            a0 = lo_ref + int(sh[i0]); a1 = hi_ref + int(sh[i1])       # haplotype coords (inclusive)
            nw = a1 - a0 + 1 - (K - 1)
            km, valid = kms[hi_]
            sl = slice(a0, a0 + nw)
            ok = valid[sl]
            rows_k.append(km[sl][ok]); rows_c.append(np.full(int(ok.sum()), ci)); rows_w.append(np.arange(nw)[ok])
            al = hap_alleles[hi_][i0:i1]
            vs = pos[i0:i1] + sh[i0:i1] - a0                             # allele start in window coords
            ln = np.where(al > 0, altlen[i0:i1], reflen[i0:i1])
            ve = vs + ln - 1
            wst = np.arange(nw)[:, None]
            cover.append(((wst <= ve[None, :]) & (wst + K - 1 >= vs[None, :]))[ok])
        allk = np.concatenate(rows_k)
        allc = np.concatenate(rows_c)
        allcov = np.concatenate(cover)
        uk, inv = np.unique(np.ascontiguousarray(allk).view(view_dt).reshape(-1), return_inverse=True)
        Kc = len(uk)
        M = np.zeros((Kc, H), np.uint8)
        np.add.at(M, (inv, allc), 1)
        # coverage bitmaps
        cov = np.zeros((Kc, n, H), bool)
        r, v = np.nonzero(allcov)
        cov[inv[r], v, allc[r]] = True
        # counts
        counts = np.zeros((Kc, S), np.uint8)
        cidx = {h: i for i, h in enumerate(cand)}
        for s in range(S):
            m_true = np.zeros(Kc, np.int64)
            for hi_ in sample_haps[s]:
                key = hap_alleles[hi_][i0:i1].tobytes()
                m_true += M[:, cidx[combos[key]]]
            nbv = rng.negative_binomial(np.maximum(size_nb * m_true, 1e-9), p_nb)
            noise = rng.poisson(noise_rate, Kc)
            counts[:, s] = np.minimum(np.where(m_true > 0, nbv, noise), 255)
        has = counts.any(axis=1)
        A["cl_nhap"].append(H)
        A["mult"].append(M.reshape(-1))
        A["k_has_counts"].append(has.astype(np.uint8))
        A["k_counts"].append(counts.reshape(-1))
        A["k_ic"].append(np.zeros(Kc * 2, np.uint8))
        A["uniq_idx"].append(np.arange(Kc, dtype=np.uint32))
        anyv = cov.any(axis=2)                                           # (Kc, n)
        rr, vv = np.nonzero(anyv)
        A["vh_var"].append(vv.astype(np.uint16))
        A["vh_bits"].append(cov[rr, vv].astype(np.uint8).reshape(-1))
        kmer_vh_cnt.append(anyv.sum(axis=1))
        vh_bits_len.append(np.full(len(rr), H, np.int64))
        A["hap_alleles"].append(np.stack([hap_alleles[h][i0:i1] for h in cand]).astype(np.uint16).reshape(-1))
        A["var_nalleles"].append(np.full(n, 2, np.uint16))
        cl_kmer_off.append(cl_kmer_off[-1] + Kc); cl_var_off.append(cl_var_off[-1] + n)
        cl_mult_off.append(cl_mult_off[-1] + Kc * H); cl_uniq_off.append(cl_uniq_off[-1] + Kc)
        cl_hapvar_off.append(cl_hapvar_off[-1] + H * n)
    C = len(ranges)
    cat = lambda k, dt: np.concatenate(A[k]).astype(dt) if A[k] else np.zeros(0, dt)
    a = {
        "sample_gender": np.array([0 if g == "F" else 1 for g in w.genders], np.uint8),
        "group_ploidy": np.tile(np.array([1 if (g == "M" and w.chrom.lower() in ("x", "chrx")) else 2 for g in w.genders], np.uint8), C),
        "group_cluster_off": np.arange(C + 1, dtype=np.uint64),
        "group_src_off": np.arange(C + 1, dtype=np.uint64), "group_src": np.zeros(C, np.uint32),
        "group_edge_off": np.zeros(C + 1, np.uint64), "group_edge_src": np.zeros(0, np.uint32), "group_edge_dst": np.zeros(0, np.uint32),
        "cluster_idx": np.zeros(C, np.uint32),
        "cl_nhap": np.array(A["cl_nhap"], np.uint32),
        "cl_kmer_off": np.array(cl_kmer_off, np.uint64), "cl_var_off": np.array(cl_var_off, np.uint64), "cl_mult_off": np.array(cl_mult_off, np.uint64),
        "mult": cat("mult", np.uint8), "k_has_counts": cat("k_has_counts", np.uint8), "k_counts": cat("k_counts", np.uint8), "k_ic": cat("k_ic", np.uint8),
        "k_shared": np.full(cl_kmer_off[-1], 0xFFFFFFFF, np.uint32),
        "cl_uniq_off": np.array(cl_uniq_off, np.uint64), "uniq_idx": cat("uniq_idx", np.uint32),
        "cl_multi_off": np.zeros(C + 1, np.uint64), "multi_idx": np.zeros(0, np.uint32),
        "kmer_vh_off": np.concatenate([[0], np.cumsum(np.concatenate(kmer_vh_cnt))]).astype(np.uint64) if C else np.zeros(1, np.uint64),
        "vh_var": cat("vh_var", np.uint16),
        "vh_bits_off": np.concatenate([[0], np.cumsum(np.concatenate(vh_bits_len))]).astype(np.uint64) if C else np.zeros(1, np.uint64),
        "vh_bits": cat("vh_bits", np.uint8),
        "cl_hapvar_off": np.array(cl_hapvar_off, np.uint64), "hap_alleles": cat("hap_alleles", np.uint16),
        "var_nalleles": cat("var_nalleles", np.uint16), "var_dep": np.zeros(cl_var_off[-1], np.uint8),
        "hap_nested_off": np.zeros(int(np.sum(A["cl_nhap"])) + 1, np.uint64), "hap_nested": np.zeros(0, np.uint32),
        "cl_dep_off": np.zeros(C + 1, np.uint64), "dep_cluster": np.zeros(0, np.uint32),
        "dep_var_off": np.zeros(1, np.uint64), "dep_var": np.zeros(0, np.uint16),
    }
    return Unit(a, S)


def tile_unit(u: Unit, times: int, seed: int = 9, depth=(15.0, 25.0), noise_rate: float = 0.02) -> Unit:
    """`times` copies of every group with independently redrawn counts (same structure):
    scales a unit to whole-chromosome cluster counts without re-enumerating k-mers."""
    if times <= 1:
        return u
    rng = np.random.default_rng(seed)
    a = u.a
    out = {}
    C = u.Cn

    def rep_off(off):
        lens = np.diff(off.astype(np.int64))
        return np.concatenate([[0], np.cumsum(np.tile(lens, times))]).astype(np.uint64)

    for k in ("group_cluster_off", "group_src_off", "group_edge_off", "cl_kmer_off", "cl_var_off", "cl_mult_off", "cl_uniq_off",
              "cl_multi_off", "kmer_vh_off", "vh_bits_off", "cl_hapvar_off", "hap_nested_off", "cl_dep_off", "dep_var_off"):
        out[k] = rep_off(a[k])
    for k in ("group_ploidy", "group_src", "group_edge_src", "group_edge_dst", "cluster_idx", "cl_nhap", "mult", "k_has_counts", "k_ic", "k_shared",
              "uniq_idx", "multi_idx", "vh_var", "vh_bits", "hap_alleles", "var_nalleles", "var_dep", "hap_nested", "dep_cluster", "dep_var"):
        out[k] = np.tile(a[k], times)
    out["sample_gender"] = a["sample_gender"]
    # redraw counts: keep zero/non-zero structure, resample the positive ones around their value
    base = a["k_counts"].astype(np.int64)
    reps = [a["k_counts"]]
    for _ in range(times - 1):
        jitter = rng.poisson(np.maximum(base, 0.05))
        reps.append(np.minimum(np.where(base > 0, np.maximum(jitter, 1), (rng.random(len(base)) < noise_rate).astype(np.int64)), 255).astype(np.uint8))
    out["k_counts"] = np.concatenate(reps)
    S = u.S
    out["k_has_counts"] = out["k_counts"].reshape(-1, S).any(axis=1).astype(np.uint8)
    return Unit(out, S)
