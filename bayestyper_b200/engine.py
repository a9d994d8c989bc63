"""Host-side mirror of the reference's stage objects over the C ABI: CountDistribution and
InferenceEngine (include/bayesTyper/CountDistribution.hpp, InferenceEngine.hpp).  Thin: every
method is one libbtgpu call; there is no CPU implementation behind it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .unit import Unit, GibbsOpts


class CountDistribution:
    """CountDistribution(samples, options) + setGenomicCountDistributions (CountDistribution.cpp:51-141)."""

    def __init__(self, nb_p, nb_size, noise_rate_prior=(1.0, 0.01)):
        self.lib = capi.load()
        self.S = len(nb_p)
        p = np.ascontiguousarray(nb_p, np.float64)
        sz = np.ascontiguousarray(nb_size, np.float64)
        self.h = capi.check(self.lib.btg_count_dist_create(self.S, capi.ptr(p), capi.ptr(sz), noise_rate_prior[0], noise_rate_prior[1]), self.lib)

    def set_noise_rates(self, rates):
        r = np.ascontiguousarray(rates, np.float64)
        assert len(r) == self.S
        capi.check(self.lib.btg_count_dist_set_noise_rates(self.h, capi.ptr(r)), self.lib)

    def finish_noise(self, chain_sums, gibbs_samples: int):
        """The end of estimateNoise from the per-chain rate sums of all ranks (btg_estimate_noise_chains): mean -> setNoiseRates."""
        cs = np.ascontiguousarray(chain_sums, np.float64)
        assert cs.ndim == 2 and cs.shape[1] == self.S
        capi.check(self.lib.btg_count_dist_finish_noise(self.h, capi.ptr(cs), cs.shape[0], gibbs_samples), self.lib)

    def noise_rates(self):
        out = np.zeros(self.S)
        capi.check(self.lib.btg_count_dist_get_noise_rates(self.h, capi.ptr(out)), self.lib)
        return out

    def tables(self):
        g = np.zeros((self.S, 256, 256)); n = np.zeros((self.S, 256))
        capi.check(self.lib.btg_count_dist_tables(self.h, capi.ptr(g), capi.ptr(n)), self.lib)
        return g, n

    def close(self):
        if self.h:
            self.lib.btg_count_dist_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class InferenceEngine:
    """InferenceEngine over one uploaded unit (InferenceEngine.hpp:56-98)."""

    def __init__(self, unit: Unit):
        self.lib = capi.load()
        self.unit = unit
        self._desc = unit.desc()
        dd = unit.dev_desc() if getattr(unit, "dev", None) else None
        if dd is None:
            self.h = capi.check(self.lib.btg_unit_upload(C.addressof(self._desc)), self.lib)
        else:   # row-level arrays stay in HBM: device-to-device copies instead of a host round trip
            self._dev_desc, n_vh, n_vh_bits = dd
            self.h = capi.check(self.lib.btg_unit_upload_dev(C.addressof(self._desc), C.addressof(self._dev_desc), n_vh, n_vh_bits), self.lib)

    @classmethod
    def from_handle(cls, unit: Unit, handle):
        """Wraps a btg_unit that already lives in HBM (btg_counter_build_unit); `unit` only sizes the result arrays."""
        self = cls.__new__(cls)
        self.lib = capi.load()
        self.unit = unit
        self.h = handle
        return self

    def estimate_genotypes(self, cd: CountDistribution, opts: GibbsOpts) -> dict:
        res, arrays = self.unit.alloc_result()
        capi.check(self.lib.btg_estimate_genotypes(self.h, cd.h, C.addressof(opts), C.addressof(res)), self.lib)
        return arrays

    def estimate_noise(self, cd: CountDistribution, opts: GibbsOpts, want_trace: bool = True, shard=None):
        """InferenceEngine::estimateNoise.  shard = btg_shard_desc (shard.shard_desc) when this engine holds one rank's
        groups of a larger unit: the selection and the noise draws are then those of the whole unit."""
        rows = opts.n_chains * (opts.gibbs_burn_in + opts.gibbs_samples + 1) + 1
        trace = np.zeros((rows, 2 + self.unit.S)) if want_trace else None
        tp = capi.ptr(trace) if want_trace else None
        if shard is None:
            capi.check(self.lib.btg_estimate_noise(self.h, cd.h, C.addressof(opts), tp), self.lib)
        else:
            capi.check(self.lib.btg_estimate_noise_sharded(self.h, cd.h, C.addressof(opts), C.addressof(shard), tp), self.lib)
        return trace

    def estimate_noise_chains(self, cd: CountDistribution, opts: GibbsOpts, chain_first: int, chain_stride: int, want_trace: bool = False):
        """The chains c = chain_first (mod chain_stride) of estimateNoise on this (whole) unit: returns ([n_chains][S] post-burn-in rate sums,
        rows of the other chains 0; trace or None).  cd's rates are set by CountDistribution.finish_noise once the ranks' sums are added up."""
        sums = np.zeros((opts.n_chains, self.unit.S))
        rows = opts.n_chains * (opts.gibbs_burn_in + opts.gibbs_samples + 1) + 1
        trace = np.zeros((rows, 2 + self.unit.S)) if want_trace else None
        capi.check(self.lib.btg_estimate_noise_chains(self.h, cd.h, C.addressof(opts), chain_first, chain_stride, capi.ptr(sums),
                                                      capi.ptr(trace) if want_trace else None), self.lib)
        return sums, trace

    def estimate_noise_and_genotypes(self, cd: CountDistribution, opts: GibbsOpts, want_trace: bool = True, shard=None):
        """InferenceEngine::estimateNoiseAndGenotypes (--noise-genotyping)."""
        res, arrays = self.unit.alloc_result()
        rows = opts.n_chains * (opts.gibbs_burn_in + opts.gibbs_samples + 1)
        trace = np.zeros((rows, 2 + self.unit.S)) if want_trace else None
        tp = capi.ptr(trace) if want_trace else None
        if shard is None:
            capi.check(self.lib.btg_estimate_noise_and_genotypes(self.h, cd.h, C.addressof(opts), C.addressof(res), tp), self.lib)
        else:
            capi.check(self.lib.btg_estimate_noise_and_genotypes_sharded(self.h, cd.h, C.addressof(opts), C.addressof(shard), C.addressof(res), tp), self.lib)
        return arrays, trace

    def cluster_tally(self, cluster: int) -> np.ndarray:
        H = int(self.unit.a["cl_nhap"][cluster])
        n = (H + 1) * (H + 2) // 2
        out = np.zeros((n, self.unit.S), np.uint32)
        capi.check(self.lib.btg_unit_cluster_tally(self.h, cluster, capi.ptr(out), out.size), self.lib)
        return out

    def close(self):
        if self.h:
            self.lib.btg_unit_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
