"""Multi-GPU sharding of an inference unit: variant-cluster GROUPS are independent, so each rank
genotypes a contiguous, cost-balanced block of groups with its global group indices (seeds unchanged) and the
results are concatenated in rank order.  No data-path collective in the default mode; the S noise rates estimated
on rank 0 are broadcast (S doubles)."""
from __future__ import annotations

import numpy as np

from .unit import Unit


def group_costs(unit: Unit) -> np.ndarray:
    """Relative Gibbs cost per group: sum over its clusters of S * D + k-mers (same model as the kernel's cost order)."""
    a = unit.a
    H = a["cl_nhap"].astype(np.int64)
    K = np.diff(a["cl_kmer_off"].astype(np.int64))
    c = unit.S * (H * (H + 1) // 2) * 8 + K
    gco = a["group_cluster_off"].astype(np.int64)
    return np.add.reduceat(c, gco[:-1]) if len(c) else np.zeros(0, np.int64)


def partition(unit: Unit, world: int):
    """Contiguous blocks of groups with (nearly) equal total cost: [(first_group, last_group_exclusive)] per rank."""
    cost = group_costs(unit).astype(np.float64)
    G = len(cost)
    if G == 0:
        return [(0, 0)] * world
    cum = np.cumsum(cost)
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(cum, cum[-1] * r / world, side="left")))
    bounds.append(G)
    bounds = np.maximum.accumulate(np.array(bounds))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def shard(unit: Unit, world: int, rank: int):
    """(sub-unit of this rank, index of its first group in the whole unit)."""
    lo, hi = partition(unit, world)[rank]
    return unit.subset_groups(np.arange(lo, hi)), lo


RESULT_KEYS = ("gt", "gq", "gpp", "app", "nak", "fak", "mac", "saf", "ploidy", "an", "ac", "af", "acp", "anc", "hc")


def concat_results(parts):
    """Concatenate per-rank result dicts (rank order == group order)."""
    return {k: np.concatenate([p[k] for p in parts]) for k in RESULT_KEYS}
