"""Multi-GPU sharding of an inference unit: variant-cluster GROUPS are independent, so each rank
genotypes a contiguous, cost-balanced block of groups with its global group indices (seeds unchanged) and the
results are concatenated in rank order.  The default mode (estimateGenotypes) has no data-path exchange.  The
lock-step modes (estimateNoise, estimateNoiseAndGenotypes) exchange the per-sample (n_obs, sum) of the noise counts
once per iteration inside the chain kernel, over peer mailboxes (csrc/comm.cuh); this module only carries the
64-byte mailbox handles between the ranks (torch.distributed, any backend) and describes the whole unit to the library."""
from __future__ import annotations

import ctypes as C

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


HANDLE_BYTES = 64


class ShardDesc(C.Structure):
    _fields_ = [("comm", C.c_void_p), ("n_groups_total", C.c_uint64), ("group_n_clusters", C.c_void_p), ("group_n_variants", C.c_void_p)]


class Comm:
    """btg_comm: this rank's mailbox + the peers' mappings.  `allgather(bytes) -> [bytes per rank]` is the host transport
    for the handles; `Comm.torch(world, rank)` uses torch.distributed (gloo or nccl)."""

    def __init__(self, world: int, rank: int, allgather=None):
        from . import capi
        self.lib = capi.load()
        self.world, self.rank = world, rank
        mine = np.zeros(HANDLE_BYTES, np.uint8)
        self.h = capi.check(self.lib.btg_comm_create(world, rank, capi.ptr(mine)), self.lib)
        if world > 1:
            handles = allgather(mine.tobytes())
            assert len(handles) == world and all(len(b) == HANDLE_BYTES for b in handles)
            buf = np.frombuffer(b"".join(handles), np.uint8).copy()
            capi.check(self.lib.btg_comm_connect(self.h, capi.ptr(buf)), self.lib)

    @classmethod
    def torch(cls, world: int, rank: int):
        import torch.distributed as dist

        def allgather(b: bytes):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        return cls(world, rank, allgather)

    def close(self):
        if self.h:
            self.lib.btg_comm_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def group_tables(unit: Unit):
    """(clusters per group, variants per group) of a unit — what every rank must know about the WHOLE unit."""
    a = unit.a
    gco = np.asarray(a["group_cluster_off"], np.int64)
    nvar = np.diff(np.asarray(a["cl_var_off"], np.int64))
    n_cl = np.diff(gco).astype(np.uint32)
    csum = np.concatenate([[0], np.cumsum(nvar)])
    return n_cl, (csum[gco[1:]] - csum[gco[:-1]]).astype(np.uint32)


def shard_desc(whole: Unit, comm: Comm | None):
    """btg_shard_desc of the whole unit for the lock-step entry points; returns (struct, keep-alive tuple)."""
    n_cl, n_var = group_tables(whole)
    n_cl, n_var = np.ascontiguousarray(n_cl), np.ascontiguousarray(n_var)
    d = ShardDesc(comm.h if comm is not None else None, len(n_cl), n_cl.ctypes.data, n_var.ctypes.data)
    return d, (n_cl, n_var, comm)
