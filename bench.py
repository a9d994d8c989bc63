#!/usr/bin/env python
"""bench.py — headline benchmark: variant clusters genotyped / second (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our arm (libbtgpu, B200)
  python bench.py --impl reference ...                   the reference's own CPU code (oracle/_ref/btref)
  python bench.py --config D ...                         configs[3] shape: 30 samples, --noise-genotyping joint mode

One "step" = one pass of both hot paths over one synthetic batch of the named shape (default configs[1]: 1 sample, chr22-like
SNV+indel candidate set, ~300k variants, k=55), bayestyper_b200/driver.py:
  k-mer match   findVariantClusterPaths (a7) -> path k-mer enumeration + exact table (a9, a12) -> genome scan (a10)
                -> sample k-mer stream (a11) -> classify + haplotype candidates (a13, a14) -> NB fit (a18)
  Gibbs         estimateNoise + estimateGenotypes (or estimateNoiseAndGenotypes): 20 chains x (100 + 250) iterations (a15-a23)
`value` times the step with the sample k-mer sets, sample Bloom filters and reference already in HBM (CUDA events on the library
stream); `e2e` runs the same call with every input in host memory (H2D inside the timed region) and the results read back.

N > 1 (torchrun), default `--mode sharded`: ONE unit of the named size is sharded over the N ranks (strong scaling) — rank r
takes every N-th group of the size-sorted unit, searches the paths of its groups, the ranks exchange their best paths (the one real
exchange of the k-mer path), and the lock-step Gibbs chains add up their noise counts once per iteration INSIDE the chain kernel
over NVLink peer mailboxes (csrc/comm.cuh; 7000 exchanges per step).  `--mode replicas`: every rank runs its own unit of the named
size with no data-path exchange (weak scaling, the round-1 line).  Time = max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "variant_clusters_genotyped_per_sec"
UNIT = "clusters/s"
WORKLOADS = {
    "B": "configs[1]: 1 sample, chr22-like SNV+indel candidate VCF (~300k variants), k=55, default Gibbs 20x(100+250)",
    "D": "configs[3] shape: 30 samples, --noise-genotyping joint mode, --max-number-of-sample-haplotypes 32, k=55, 20x(100+250); a 2.7 Mb / 20k-variant "
         "slice of the whole-genome set (population allele frequencies ~ Beta(0.2,0.8), Hardy-Weinberg genotypes), SURVEY.md section 8d",
}
WORKLOAD = WORKLOADS["B"]


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks + throttle reasons sampled DURING the timed region, through NVML in-process (nvidia_ml_py): spawning nvidia-smi every
    100 ms (round 1) takes the driver lock ~10 times a second and slowed the step it was watching by 15 %; even NVML at 5 Hz slowed the
    persistent lock-step kernels progressively (2.1 -> 3.9 s per step over four steps), so the rate is 1 Hz.  Falls back to nvidia-smi
    when NVML cannot be loaded."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []          # (sm_mhz, sm_max_mhz, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap)
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        self._max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                idx = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        if self._max is None:
            self._max = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = self._max
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bit = lambda name_new, name_old: getattr(n, name_new, getattr(n, name_old, 0))
        self.rows.append((sm, mx,
                          bool(r & bit("nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown")),
                          bool(r & bit("nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown")),
                          bool(r & bit("nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown")),
                          bool(r & bit("nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            if f[0].isdigit():
                self.rows.append((int(f[0]), int(f[1]) if f[1].isdigit() else None, *[x.lower().startswith("active") for x in f[2:6]]))

    def _run(self):
        if os.environ.get("BTG_BENCH_NO_CLOCKS"):
            return
        while not self._stop.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(1.0)     # 1 Hz: every query of the driver slows the lock-step chain kernel it is watching (5 Hz cost the step up to 1.8x)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# synthetic batch
# ------------------------------------------------------------------------------------------------
def device_spectrum(lib, haplotypes, mean, var, seed, n_errors, dev):
    """The sample's KMC-like k-mer spectrum, synthesised on the device (setup, untimed): canonical 55-mers of the
    haplotypes (btg_scan_sequence_dev), distinct k-mers with copy number, NB(mean*copies, var*copies) counts."""
    import torch
    from bayestyper_b200 import capi
    chunks = []
    for h in haplotypes:
        t = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(dev)
        km = torch.empty((t.numel(), 2), dtype=torch.int64, device=dev)
        valid = torch.empty(t.numel(), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        capi.check(lib.btg_scan_sequence_dev(t.data_ptr(), t.numel(), km.data_ptr(), valid.data_ptr(), None), lib)
        torch.cuda.synchronize()
        chunks.append(km[valid.to(torch.bool)])
        del km, valid, t
    km = torch.cat(chunks)
    del chunks
    o = torch.sort(km[:, 0], stable=True).indices
    o = o[torch.sort(km[o, 1], stable=True).indices]
    km = km[o]
    del o
    new = torch.ones(len(km), dtype=torch.bool, device=dev)
    new[1:] = (km[1:] != km[:-1]).any(1)
    first = torch.nonzero(new).squeeze(1)
    copies = torch.diff(torch.cat([first, torch.tensor([len(km)], device=dev)])).to(torch.float64)
    keys = km[new].contiguous()
    del km, new
    g = torch.Generator(device=dev).manual_seed(seed)
    p = mean / var
    size = mean * mean / (var - mean)
    lam = torch._standard_gamma(size * copies, generator=g) * ((1 - p) / p)
    counts = torch.clamp(torch.poisson(lam, generator=g), max=255).to(torch.uint8)
    keep = counts > 0
    keys, counts = keys[keep], counts[keep]
    if n_errors:
        err = torch.empty((n_errors, 2), dtype=torch.int64, device=dev).random_(generator=g)
        err[:, 1] &= (1 << 46) - 1
        keys = torch.cat([keys, err]); counts = torch.cat([counts, torch.ones(n_errors, dtype=torch.uint8, device=dev)])
    # a KMC database lists its records in lexicographic k-mer order (kmc_file.cpp:428-515): put the spectrum in that order
    keys = keys.contiguous()
    k_lo = torch.empty(len(keys), dtype=torch.int64, device=dev); k_hi = torch.empty_like(k_lo)
    torch.cuda.synchronize()
    capi.check(lib.btg_table_keys_from_kmers_dev(keys.data_ptr(), len(keys), k_lo.data_ptr(), k_hi.data_ptr(), None), lib)
    capi.check(lib.btg_kmer_hash(capi.ptr(np.zeros((1, 2), np.uint64)), 1, capi.ptr(np.zeros(1, np.uint64))), lib)   # library-stream sync
    o = torch.sort(k_lo, stable=True).indices
    o = o[torch.sort(k_hi[o], stable=True).indices]
    return keys[o].contiguous(), counts[o].contiguous()


def build_batch_d(lib, rank: int, scale: float, dev):
    """configs[3] shape on one contig: 30 samples (alternating genders, diploid chr1), 20k SNV/indel candidates on 2.7 Mb, population
    allele frequencies ~ Beta(0.2, 0.8) (seed 31), Hardy-Weinberg genotypes (seed 32), 30x spectra synthesised on the device."""
    from bayestyper_b200 import driver, synth
    S = 30
    n_var = max(60, int(20_000 * scale))
    ref = synth.random_reference(n_var * 136, 11 + 1000 * rank)
    var = synth.make_variants(ref, n_var, 12 + 1000 * rank, 0.075, 0.075)
    af = np.random.default_rng(31 + 1000 * rank).beta(0.2, 0.8, size=len(var))
    g = synth.make_genotypes(len(var), S, 32 + 1000 * rank, allele_freq=af)
    spectra = []
    for s_ in range(S):
        haps = [synth.apply_variants(ref, var, g[s_, :, h]) for h in range(2)]
        spectra.append(device_spectrum(lib, haps, 15.0, 25.0, 14 + 1000 * rank + 7 * s_, int(50_000 * scale), dev))
    inp = driver.Inputs("chr1", ref, var, ["F" if i % 2 == 0 else "M" for i in range(S)], spectra=None)
    inp.spectra_dev = spectra
    inp.truth = g
    inp.prepare()
    return inp


def build_batch(lib, rank: int, scale: float, dev):
    """configs[1]: chr22-like reference (10 Mb N + 40.8 Mb), ~300k SNV/indel candidates, 1 female sample at 30x."""
    from bayestyper_b200 import driver, synth
    n_var = max(200, int(300_000 * scale))
    n_prefix = int(10_000_000 * scale)
    length = int(40_800_000 * scale) + n_prefix
    ref = synth.random_reference(length, 11 + 1000 * rank, n_prefix)
    var = synth.make_variants(ref, n_var, 12 + 1000 * rank, 0.075, 0.075, lo=n_prefix + 55)
    g = synth.make_genotypes(len(var), 1, 13 + 1000 * rank)
    haps = [synth.apply_variants(ref, var, g[0, :, h]) for h in range(2)]
    keys, counts = device_spectrum(lib, haps, 15.0, 25.0, 14 + 1000 * rank, int(500_000 * scale), dev)
    inp = driver.Inputs("chr22", ref, var, ["F"], spectra=None)
    inp.spectra_dev = [(keys, counts)]
    inp.truth = g
    inp.prepare()
    return inp


def truth_agreement(inp, res, S):
    """Correctness of the FULL-SIZE step, printed in the bench line (`parity_at_size`): the genotypes called by the last timed step against
    the genotypes the sample spectra were synthesised from (alt-allele count per variant and sample; tests/test_gpu_e2e.py asserts the same
    quantity on the small fixtures, where the reference's own calls are available too).  Not a substitute for the oracle parity tests: a
    property the domain offers at a size the CPU reference needs minutes for."""
    g = getattr(inp, "truth", None)
    if g is None or res is None:
        return None
    try:
        truth = np.asarray(g).transpose(1, 0, 2).sum(axis=2)                 # (variants, S) alt-allele count
        pos = np.array([v.pos + 1 for v in inp.variants], np.int64)
        order = np.argsort(pos, kind="stable")
        vp = np.asarray(inp.graphs["var_pos"], np.int64)
        at = np.searchsorted(pos[order], vp)
        if len(vp) != len(res["gt"]) // (2 * S) or (pos[order][np.minimum(at, len(pos) - 1)] != vp).any():
            return {"error": "result rows do not line up with the candidate variants"}
        t = truth[order[at]]
        gt = np.asarray(res["gt"]).reshape(-1, S, 2)
        called = gt[..., 0] != 0xFFFF
        alt = gt.astype(np.int64).sum(axis=2)
        return {"what": "called genotype (alt-allele count) == genotype the spectra were drawn from, over called (variant, sample) pairs of the last timed step",
                "variant_sample_pairs": int(called.size), "called_frac": float(called.mean()),
                "gt_matches_truth_frac": float((alt[called] == t[called]).mean()) if called.any() else None,
                "hom_ref_truth_frac": float((t == 0).mean())}
    except Exception as e:          # the check must never cost the bench line
        return {"error": f"{type(e).__name__}: {e}"}


def run_ours(args):
    import torch
    from bayestyper_b200 import capi, driver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line (NCCL prints its version banner there otherwise)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    lib = capi.load()
    capi.check(lib.btg_init(local_rank), lib)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(lib.btg_get_stream(), device=dev)
    opt = driver.Options(random_seed=20190401, noise_genotyping=args.config == "D", noise_split=args.noise_split, gibbs_samples=args.gibbs_samples)
    sharded = world > 1 and args.mode in ("auto", "sharded")
    shard_ctx = None
    if sharded:   # one unit over all ranks: every rank builds the SAME batch; mailbox handles + best paths travel over torch.distributed
        from bayestyper_b200 import shard as shard_mod

        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        shard_ctx = driver.Shard(world, rank, allgather, shard_mod.Comm.torch(world, rank))

    t0 = time.time()
    inp = (build_batch_d if args.config == "D" else build_batch)(lib, 0 if sharded else rank, args.scale, dev)
    inp.make_resident(lib, opt)
    setup_s = time.time() - t0
    # the synthetic batch is ~1.5 M long-lived Python objects (300k Variant records, the graph arrays' wrappers): a full collection of the cyclic
    # GC walks all of them (75 ms for the Variant records alone on this container) whenever it triggers inside a step — the suspected cause of the one step in three
    # that took 150-300 ms longer in profiles/r2_bench_n1_v{1,2}.json (not re-measured: the round's GPU minutes were spent).  Park them in the
    # permanent generation; the steps' own garbage is still collected.
    import gc
    gc.collect()
    gc.freeze()
    n_sample = int(sum(k.shape[0] for k, _ in inp.spectra_dev))
    S = len(inp.spectra_dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM -------------------------------------------------------------
    info = None
    for _ in range(args.warmup):
        _, _, _, info = driver.genotype(inp, opt, resident=True, shard=shard_ctx)
    n_clusters = info["n_clusters"] if info else None
    barrier()
    lib.btg_launch_count_reset()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_wall = []
        for _ in range(args.steps):
            t_step = time.perf_counter()
            _, _, res, info = driver.genotype(inp, opt, resident=True, shard=shard_ctx)     # returns with the results on the host: the step has ended
            step_wall.append((time.perf_counter() - t_step) * 1e3)
            if info.get("stage_wall_ms"):
                print("step stages (BTG_STAGE_TIMES):", {k: round(v, 1) for k, v in info["stage_wall_ms"].items()}, file=sys.stderr)
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
    launches = int(lib.btg_launch_count())
    n_clusters = info["n_clusters"]
    parity_at_size = None if sharded else truth_agreement(inp, res, S)      # a shard's result rows cover its own groups only
    n_total = info["n_clusters_total"] if sharded else world * n_clusters      # clusters genotyped by ALL ranks in one step
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_total / (ms_per_step / 1e3)

    if args.timed_only:          # profiler runs (ncu launch list): the warm-up + timed steps only; not a bench line
        if rank == 0:
            emit({"timed_only": True, "ms_per_step": ms_per_step, "value": value, "gpu_launches": launches, "n_clusters": n_clusters, "step_wall_ms": step_wall})
        inp.free(lib)
        return
    # ---- e2e: every input crosses the boundary from host memory inside the call --------------------------
    h_spectra, h_blooms, h2d, bloom_total = [], [], 0, 0
    nk, nb, nh = C.c_uint64(), C.c_uint64(), C.c_uint32()
    for (keys_d, counts_d), bl in zip(inp.spectra_dev, inp.blooms_dev):
        h_keys = torch.empty(keys_d.shape, dtype=torch.int64, pin_memory=True); h_keys.copy_(keys_d)
        h_counts = torch.empty(counts_d.shape, dtype=torch.uint8, pin_memory=True); h_counts.copy_(counts_d)
        lib.btg_bloom_info(bl, C.byref(nk), C.byref(nb), C.byref(nh))
        bloom_bytes = np.zeros((nb.value + 7) // 8, np.uint8)
        capi.check(lib.btg_bloom_download(bl, capi.ptr(bloom_bytes), bloom_bytes.size), lib)
        h_spectra.append((h_keys, h_counts)); h_blooms.append((bloom_bytes, nk.value, nb.value))
        h2d += h_keys.numel() * 8 + h_counts.numel() + bloom_bytes.size
        bloom_total += nb.value
    host_inp = driver.Inputs(inp.chrom, inp.reference, inp.variants, inp.genders, spectra=h_spectra, blooms=h_blooms, graphs=inp.graphs, regions=inp.regions)
    h2d += int(inp.region_buf_dev.numel())
    torch.cuda.synchronize()
    _, unit, res, _ = driver.genotype(host_inp, opt, resident=False, want_unit=not sharded, shard=shard_ctx)      # warm-up; also sizes the unit traffic
    from bayestyper_b200.unit import Unit as _Unit
    small = 0 if unit is None else sum(v.nbytes for k, v in unit.a.items() if k not in _Unit.DEVICE_FIELDS)   # per-cluster / per-group descriptors cross (down, up);
    h2d += small                                                                          # the row-level arrays stay in HBM (btg_unit_upload_dev)
    d2h = sum(v.nbytes for v in res.values()) + small
    barrier()
    e2e_steps = args.steps                                                                # the same number of steps as `value`
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        driver.genotype(host_inp, opt, resident=False, shard=shard_ctx)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t1) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_total / float(t.item())

    # ---- stage breakdown + roofline of the k-mer-match stream kernel (measured live, CUDA events) ---------
    stage_ms, roof = stage_breakdown(lib, inp, opt, stream, dev, shard_ctx)
    paths_roof = stage_ms.pop("__paths_roofline", None)

    peak, peak_src = measured_peaks()
    roof.update({"peak": peak, "unit": "GB/s", "frac": roof["achieved"] / peak, "peak_source": peak_src, "bound": "hbm", "traffic": stream_traffic(roof)})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "f64 (Gibbs log-likelihoods) / u64 (k-mer hashing)", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "clusters": n_total, "clusters_per_gpu": n_clusters, "variants": len(inp.variants) * (1 if sharded else world),
                   "samples": S, "haplotype_candidates_per_cluster": info.get("haplotype_candidates"), "reference_nt": len(inp.reference), "sample_kmers": n_sample, "path_kmers": info["n_path_kmers"],
                   "step": "findVariantClusterPaths -> path k-mer table -> genome scan -> sample k-mer stream -> classify/haplotype candidates -> NB fit -> "
                           + ("estimateNoiseAndGenotypes" if args.config == "D" else "estimateNoise -> estimateGenotypes"),
                   "gibbs": "20 chains x (100 burn-in + %d samples), k-mer subsampling 0.1" % opt.gibbs_samples,
                   "l2": "inputs exceed L2 (sample k-mer streams %.2f GB, sample Bloom filters %.0f MB)" % (n_sample * 17 / 1e9, bloom_total / 8e6),
                   "parallelism": ("1 GPU" if world == 1 else
                                   ("ONE unit sharded over %d ranks (every %d-th group of the size-sorted unit): path search + estimateGenotypes per rank, best paths all-gathered; "
                                    % (world, world)
                                    + ("estimateNoise: rank r runs the chains r, r + %d, ... of the whole unit (chains are independent streams), per-chain sums all-gathered" % world
                                       if args.config == "B" and args.noise_split == "chains" else
                                       "noise counts of the lock-step chains added up inside the chain kernel over NVLink peer mailboxes (%d exchanges per step)" % (20 * (100 + opt.gibbs_samples))))
                                   if sharded else "%d replicas, one unit of the named size each, no data-path exchange" % world),
                   "scale": args.scale},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roof,
        "roofline_paths": None if paths_roof is None else {**paths_roof, "peak": peak, "unit": "GB/s", "frac": paths_roof["achieved"] / peak, "peak_source": peak_src},
        "parity_at_size": parity_at_size,
        "stage_ms": stage_ms,
        "step_wall_ms": step_wall,
        "setup_s": setup_s,
    }
    if rank == 0:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args.config, args.cpu_variants or (3000 if args.config == "B" else 120), os.cpu_count() or 1, args.gibbs_samples)
        emit(line)
    inp.free(lib)
    if shard_ctx is not None:
        shard_ctx.comm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def stream_traffic(roof):
    """DRAM bytes per launch of the stream kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture of
    THIS bench's launch (profiles/r2_stream_traffic.json, written by tools/ncu_stream_traffic.py from an `ncu --set full` capture of
    `bench.py --timed-only`); null when the capture is of another launch shape."""
    p = ROOT / "profiles" / "r2_stream_traffic.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text())
    return d["dram_bytes"] if (d.get("records"), d.get("table_keys")) == (roof["records"], roof["table_keys"]) else None


def stage_breakdown(lib, inp, opt, stream, dev, shard_ctx=None):
    """One extra (untimed for `value`) pass with a synchronisation after every stage (a sharded run reports rank 0's stages; the
    exchanges are inside them), and the isolated timing of the stream kernel k_table_add_sample for the roofline object."""
    import torch
    from bayestyper_b200 import capi, driver, engine, kmer_pipeline, unit as U
    out = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t) * 1e3
        return r

    S = len(inp.spectra_dev)
    sharded = shard_ctx is not None
    G = len(inp.graphs["group_cluster_off"]) - 1
    if sharded:
        mine = shard_ctx.my_groups(G)
        sub, my_clusters = driver.subset_graphs_for_paths(inp.graphs, mine)
        part = timed("findVariantClusterPaths (own groups)", lambda: driver.find_variant_cluster_paths(lib, sub, inp.blooms_dev, opt))
        parts = timed("best paths all-gather", lambda: shard_ctx.allgather((part[0], part[1], my_clusters)))
        n_paths, mem = driver.merge_best_paths(inp.graphs, [(p[0], p[1]) for p in parts], [p[2] for p in parts])
    else:
        n_paths, mem = timed("findVariantClusterPaths", lambda: driver.find_variant_cluster_paths(lib, inp.graphs, inp.blooms_dev, opt))
    kd, cdv = inp.spectra_dev[0]
    sdesc = keep = None
    if sharded:   # a sharded unit subsets the unit on the device: the torch-glue mirror of the k-mer stages (driver.genotype does the same)
        pipe = kmer_pipeline.KmerPipeline(inp.graphs, n_paths, mem, S, inp.genders)
        timed("countPathKmers(enumerate+sort)", pipe.enumerate_path_kmers)
        timed("countInterclusterKmers(scan)", lambda: pipe.scan_buffer(inp.region_buf_dev, 2, 2, False))
        timed("parseSampleKmers(stream, all samples)", lambda: [pipe.add_sample(i, kd_, cd_) for i, (kd_, cd_) in enumerate(inp.spectra_dev)])
        unit = timed("classify+getHaplotypeCandidates", lambda: pipe.build_unit(multigroup_bloom=None, device_resident=True))
        nb = timed("NB fit (parameter k-mers)", lambda: driver.estimate_nb_parameters(pipe, inp.region_buf_dev, inp.spectra_dev, inp.genders, opt))
        from bayestyper_b200 import shard as shard_mod
        chain_rates = None
        if not opt.noise_genotyping and opt.noise_split == "chains":      # estimateNoise by chains on the whole unit (driver.genotype does the same)
            whole = timed("whole-unit upload (noise chains)", lambda: engine.InferenceEngine(unit))
            cd_w = engine.CountDistribution(nb[0], nb[1])
            wopts = U.default_opts(seed=opt.random_seed, min_frac=U.min_fraction_observed(nb[0], nb[1]))
            sums = timed("estimateNoise (chains rank::world of the whole unit)", lambda: whole.estimate_noise_chains(cd_w, wopts, shard_ctx.rank, shard_ctx.world)[0])
            whole.close()
            total = timed("noise chain sums all-gather", lambda: sum(shard_ctx.allgather(sums)))
            cd_w.finish_noise(total, opt.gibbs_samples)
            chain_rates = cd_w.noise_rates()
            cd_w.close()
        sdesc, keep = shard_mod.shard_desc(unit, shard_ctx.comm)
        unit = timed("unit subset (own groups, on the device)", lambda: unit.subset_groups(mine))
        eng = timed("unit upload", lambda: engine.InferenceEngine(unit))
    else:         # the product path of one GPU: every stage through a handle of the C ABI (csrc/counter.cu)
        from bayestyper_b200 import counter
        kc = counter.KmerCounter(inp.graphs, n_paths, mem, S, inp.genders)
        timed("countPathKmers(enumerate+sort)", kc.count_path_kmers)
        timed("countInterclusterKmers(scan)", lambda: kc.count_intercluster_kmers(inp.region_buf_dev.data_ptr(), inp.region_buf_dev.numel(), 2, 2, False))
        timed("parseSampleKmers(stream, all samples)", lambda: [kc.parse_sample_kmers(i, kd_.data_ptr(), cd_.data_ptr(), cd_.numel()) for i, (kd_, cd_) in enumerate(inp.spectra_dev)])
        handle = timed("classify+getHaplotypeCandidates+unit", lambda: kc.build_unit(np.full(G * S, 2, np.uint8)))
        nb = timed("NB fit (parameter k-mers)", lambda: kc.fit_nb(inp.region_buf_dev.data_ptr(), inp.region_buf_dev.numel(), inp.spectra_dev, (2, 2), None, opt.random_seed, opt.max_parameter_kmers))
        sizes = U.Unit({**{k_: np.zeros(0, dt) for k_, dt in U._DESC_FIELDS}, "group_cluster_off": inp.graphs["group_cluster_off"], "cl_kmer_off": np.zeros(kc.Cn + 1, np.uint64),
                        "cl_var_off": inp.graphs["cl_var_off"], "var_nalleles": kc._keep["var_nalleles"], "cl_nhap": np.asarray(n_paths, np.uint32)}, S)
        eng = engine.InferenceEngine.from_handle(sizes, handle)
        kc.close()
        pipe = kmer_pipeline.KmerPipeline(inp.graphs, n_paths, mem, S, inp.genders)      # the same table once more, for the isolated timing of the stream kernel below
        pipe.enumerate_path_kmers()
    cd = engine.CountDistribution(nb[0], nb[1])
    gopts = U.default_opts(seed=opt.random_seed, min_frac=U.min_fraction_observed(nb[0], nb[1]),
                           group_base=shard_ctx.rank if sharded else 0, group_stride=shard_ctx.world if sharded else 1)
    if opt.noise_genotyping:
        timed("estimateNoiseAndGenotypes", lambda: eng.estimate_noise_and_genotypes(cd, gopts, want_trace=False, shard=sdesc))
    else:
        if sharded and chain_rates is not None:
            cd.set_noise_rates(chain_rates)
        else:
            timed("estimateNoise", lambda: eng.estimate_noise(cd, gopts, want_trace=False, shard=sdesc))
        timed("estimateGenotypes", lambda: eng.estimate_genotypes(cd, gopts))
    eng.close(); cd.close()
    # roofline of the k-mer-match kernel that costs the time: k_find_sample_paths, one sample's pass over the (rank's) graphs, CUDA events on the library stream
    import ctypes as _C
    from bayestyper_b200.driver import GraphsDesc
    gsrc = sub if sharded else inp.graphs
    gco_ = gsrc["group_cluster_off"]
    keep_ = {"cl_vertex_off": np.ascontiguousarray(gsrc["cl_vertex_off"], np.uint64), "v_seq_off": np.ascontiguousarray(gsrc["v_seq_off"], np.uint64),
             "seq": np.ascontiguousarray(gsrc["seq"], np.uint8), "v_flags": np.ascontiguousarray(gsrc["v_flags"], np.uint8),
             "v_in_off": np.ascontiguousarray(gsrc["v_in_off"], np.uint64), "v_in_src": np.ascontiguousarray(gsrc["v_in_src"], np.uint32),
             "cl_group": np.ascontiguousarray(gsrc["cl_group_global"], np.uint32) if "cl_group_global" in gsrc
                         else np.repeat(np.arange(len(gco_) - 1, dtype=np.uint32), np.diff(np.asarray(gco_, np.int64))),
             "cl_idx": np.ascontiguousarray(gsrc["cluster_idx"], np.uint32)}
    gd_ = GraphsDesc(); gd_.n_clusters = len(keep_["cl_vertex_off"]) - 1
    for k_, v_ in keep_.items():
        setattr(gd_, k_, v_.ctypes.data)
    gr_ = capi.check(lib.btg_graphs_upload(_C.addressof(gd_), 1, opt.max_sample_haplotypes), lib)
    capi.check(lib.btg_find_sample_paths(gr_, inp.blooms_dev[0], 0, opt.random_seed, opt.max_sample_haplotypes), lib)      # warm-up
    capi.check(lib.btg_graphs_reset(gr_), lib)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    capi.check(lib.btg_find_sample_paths(gr_, inp.blooms_dev[0], 0, opt.random_seed, opt.max_sample_haplotypes), lib)
    p1.record(stream)
    stream.synchronize()
    pst = np.zeros(3, np.uint64)
    capi.check(lib.btg_graphs_path_stats(gr_, pst.ctypes.data), lib)
    lib.btg_graphs_free(gr_)
    p_ms = p0.elapsed_time(p1)
    p_alg = int(pst[1]) * 32 + int(pst[2]) // 4
    out["__paths_roofline"] = {"kernel": "k_find_sample_paths (findSamplePaths: one sample's Bloom filter probed for every k-mer of every candidate path, a7)",
                               "achieved": p_alg / (p_ms / 1e3) / 1e9, "algorithmic_bytes_per_launch": p_alg, "ms_per_launch": p_ms, "kmer_lookups": int(pst[0]),
                               "bloom_probes_executed": int(pst[1]), "nucleotides": int(pst[2]),
                               "bytes_model": "32 B per Bloom probe executed under the reference's early-exit order + 0.25 B per nucleotide walked (SURVEY.md section 8d)",
                               "bound": "latency of the sequential vertex DP per cluster (lane 0) between the lane-parallel probe rounds; the sample's filter is L2-resident"}
    # roofline: the sample k-mer stream probing the exact path-k-mer table (sample 0)
    pipe.use_index()
    counts = torch.zeros_like(pipe.counts); rec = torch.zeros_like(pipe.has_record)
    torch.cuda.synchronize()
    n = kd.shape[0]
    for _ in range(3):
        capi.check(lib.btg_table_add_sample_kmers_dev(pipe.kw0.data_ptr(), pipe.kw1.data_ptr(), pipe.n_keys, kd.data_ptr(), cdv.data_ptr(), n, S, 0, counts.data_ptr(), rec.data_ptr(), None), lib)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        capi.check(lib.btg_table_add_sample_kmers_dev(pipe.kw0.data_ptr(), pipe.kw1.data_ptr(), pipe.n_keys, kd.data_ptr(), cdv.data_ptr(), n, S, 0, counts.data_ptr(), rec.data_ptr(), None), lib)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    hits = int(rec.sum())
    alg = n * 17 + pipe.n_keys * 16 + hits * 1
    roof = {"kernel": "k_table_add_sample (parseSampleKmers: sample k-mer stream probing the exact path-k-mer table, a11)",
            "achieved": alg / (ms / 1e3) / 1e9, "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms,
            "records": n, "table_keys": pipe.n_keys, "hits": hits,
            "bytes_model": "17 B per record + 16 B per path k-mer + 1 B per hit (SURVEY.md section 8d, merge-join form)"}
    return out, roof


# ------------------------------------------------------------------------------------------------
# the reference's own CPU path (oracle-R), bounded sample
# ------------------------------------------------------------------------------------------------
# configs[1] as build_batch() makes it at --scale 1 (seeded: the same candidate set on every run; the numbers are the ones our arm's line reports);
# InferenceEngine.cpp:50 noise_variants_batch_size
FULL_B = {"variants": 299_455, "clusters": 200_730, "noise_cap_variants": 100_000}


def cpu_reference(config: str, n_variants: int, threads: int, gibbs_samples: int = 250):
    """Runs the reference's translation units (oracle/_ref/btref: cluster + genotype stage order) on a bounded sample of the same
    workload shape, all host threads.

    configs[1] (B): the sample has the full config's per-cluster shape but the reference's estimateNoise works on at most 100,000
    variants (InferenceEngine.cpp:50): in the full 300k-variant config it touches one third of the unit, in a small sample all of it.
    `value` is therefore the FULL-CONFIG rate projected from the sample's measured stage times — k-mer stages and estimateGenotypes
    scale with the clusters, estimateNoise with min(variants, 100k) — i.e. full_clusters / (t_kmer * r + t_geno * r + t_noise * r_noise),
    r = 300k / sample variants, r_noise = 100k / sample variants; the unprojected sample rate is reported beside it.
    configs[3] shape (D): the joint mode touches every cluster in every iteration, so the sample rate is the rate."""
    from bayestyper_b200 import synth
    btref = ROOT / "oracle" / "_ref" / "btref"
    if not btref.exists():
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref/btref not built"}
    S = 30 if config == "D" else 1
    length = int(n_variants * 136)
    ref = synth.random_reference(length, 11)
    var = synth.make_variants(ref, n_variants, 12, 0.075, 0.075)
    if config == "D":
        af = np.random.default_rng(31).beta(0.2, 0.8, size=len(var))
        g = synth.make_genotypes(len(var), S, 32, allele_freq=af)
        w = synth.Workload("D-sample", "chr1", ref, var, g, ["F" if i % 2 == 0 else "M" for i in range(S)])
    else:
        g = synth.make_genotypes(len(var), 1, 13)
        w = synth.Workload("B-sample", "chr22", ref, var, g, ["F"])
    with tempfile.TemporaryDirectory() as td:
        synth.write_workdir(w, td, n_errors=max(1000, int(50_000 * n_variants / 3000)) if config == "B" else 2000)
        t0 = time.time()
        subprocess.check_call([str(btref), "run", "--workdir", td, "--threads", str(threads), "--seed", "20190401", "--gibbs-samples", str(gibbs_samples)] + (["--noise-genotyping"] if config == "D" else []),
                              stdout=subprocess.DEVNULL)
        wall = time.time() - t0
        tj = json.loads((Path(td) / "ref_out" / "timings.json").read_text())
    kmer_s = sum(tj.get(k, 0.0) for k in ("findVariantClusterPaths", "countPathMultigroupKmers", "countPathKmers", "countInterclusterKmers", "parseSampleKmers", "classifyPathKmers"))
    geno_s, noise_s = tj.get("estimateGenotypes", 0.0) + tj.get("estimateNoiseAndGenotypes", 0.0), tj.get("estimateNoise", 0.0)
    total_s = geno_s + noise_s + kmer_s
    sample_rate = tj["clusters_genotyped"] / total_s
    out = {"unit": UNIT, "cores": threads, "kind": "reference", "step_s": total_s, "clusters": tj["clusters_genotyped"], "sample_variants": len(var),
           "measured_sample_value": sample_rate, "estimateGenotypes_s": geno_s, "estimateNoise_s": noise_s, "kmer_stages_s": kmer_s}
    try:    # the one-off measurement of the FULL config through the reference (tools/reference_full_config.py, build container): quoted beside the projection
        fm = json.loads((ROOT / "profiles" / f"r2_reference_full_config{config}.json").read_text())
        out["full_config_measured"] = {"clusters_per_s": fm["clusters_per_s"], "step_s": fm["step_s"], "clusters": fm["clusters_genotyped"], "variants": fm["variants"],
                                       "threads": fm["threads"], "host": fm["host"], "profile": f"profiles/r2_reference_full_config{config}.json"}
    except Exception:
        pass
    if config == "B":
        r = FULL_B["variants"] / len(var)
        r_noise = min(FULL_B["variants"], FULL_B["noise_cap_variants"]) / min(len(var), FULL_B["noise_cap_variants"])
        full_s = (kmer_s + geno_s) * r + noise_s * r_noise
        out.update({"value": FULL_B["clusters"] / full_s, "projected_full_step_s": full_s,
                    "sample": f"{len(var)} variants / {tj['num_clusters']} clusters of the same chr22-like shape through the reference's own stages, {threads} threads "
                              f"(estimateGenotypes {geno_s:.2f} s, estimateNoise {noise_s:.2f} s, k-mer stages {kmer_s:.2f} s, wall {wall:.1f} s = {sample_rate:.0f} clusters/s on the sample); "
                              f"value = rate of the full config ({FULL_B['variants']} variants / {FULL_B['clusters']} clusters) projected from these stage times (estimateNoise is capped at 100k variants, InferenceEngine.cpp:50: "
                              f"x{r_noise:.1f}; the other stages x{r:.1f})"})
    else:
        out.update({"value": sample_rate,
                    "sample": f"{len(var)} variants / {tj['num_clusters']} clusters x {S} samples of the same shape through the reference's own stages, {threads} threads "
                              f"(estimateNoiseAndGenotypes {geno_s:.2f} s, k-mer stages {kmer_s:.2f} s, wall {wall:.1f} s); the joint mode touches every cluster in every "
                              f"iteration, so the sample rate is the rate of the shape"})
    return out


def reference_sample_variants(config: str, steps: int, warmup: int) -> int:
    """Variants per sample run of the reference arm: (one warm-up run when W > 0) + K timed runs must end within a few minutes."""
    runs = steps + (1 if warmup > 0 else 0)
    if config == "B":
        return int(min(12_000, max(2_000, 60_000 // max(1, runs))))
    return int(min(120, max(40, 720 // max(1, runs))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # every step is one run of the reference on a bounded sample; the sample is sized so that the W > 0 warm-up run plus the K timed runs end
    # within a few minutes on a 16-thread host (measured: ~3.5 ms of wall per sample variant for B, ~0.25 s for the 30-sample D shape):
    # K = 1..4 -> 12,000 variants per run, the driver's K = 20 -> 2,857.  The projection to the full config (cpu_reference) is the same
    # for every sample size; the size is printed in config.sample.
    n_var = args.cpu_variants if args.cpu_variants else reference_sample_variants(args.config, args.steps, args.warmup)
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue            # a CPU process has no clocks or caches to warm beyond the first run: one warm-up run stands for all W
        last = cpu_reference(args.config, n_var, threads, args.gibbs_samples)
        if last["value"] is None:
            emit({"impl": "reference", "unavailable": last["sample"]})
            return
        if i >= args.warmup:
            vals.append(last)
    value = float(np.mean([v["value"] for v in vals]))
    ms = float(np.mean([v["step_s"] * 1e3 for v in vals]))            # MEASURED: the reference's stages on one bounded sample (what the K timed steps actually took)
    proj = [v["projected_full_step_s"] * 1e3 for v in vals if "projected_full_step_s" in v]
    cb = dict(last); cb["value"] = value
    full_clusters = FULL_B["clusters"] if args.config == "B" else last["clusters"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 and args.mode != "replicas" else "weak", "vs_baseline": None,
            "dtype": "f64 (Gibbs log-likelihoods) / u64 (k-mer hashing)", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "clusters": full_clusters, "variants": FULL_B["variants"] if args.config == "B" else last["sample_variants"],
                       "samples": 30 if args.config == "D" else 1, "gibbs": "20 chains x (100 burn-in + %d samples), k-mer subsampling 0.1" % args.gibbs_samples,
                       "step": "the reference's own cluster + genotype stages on the host cores", "sample": last["sample"], "threads": threads},
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if proj:   # `value` is the full-config rate projected from the measured stage times (cpu_reference); the full step itself is never run
        line["projected_full_step_ms"] = float(np.mean(proj))
        line["measured_sample_value"] = float(np.mean([v["measured_sample_value"] for v in vals]))
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: everything else that libraries print there (NCCL's version banner at communicator creation,
    whatever NCCL_DEBUG_FILE says) is sent to stderr by pointing fd 1 at fd 2; emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the chr22-sized batch (tests use small values)")
    ap.add_argument("--config", default="B", choices=["B", "D"], help="B = configs[1] (default, the metric's config); D = configs[3] shape (30 samples, joint mode)")
    ap.add_argument("--noise-split", default="chains", choices=["chains", "groups"],
                    help="sharded estimateNoise: every rank runs its share of the independent chains on the whole unit (default), or all chains on its own groups "
                         "with the per-iteration noise counts added up inside the chain kernel over NVLink mailboxes (what the joint mode of --config D always does)")
    ap.add_argument("--gibbs-samples", type=int, default=250, help="--gibbs-samples of the reference (post-burn-in iterations per chain; default 250): the sweep axis of BASELINE.json configs[4] "
                                                                    "(250 / 500 / 1000 / 2000), e.g. --config D --gibbs-samples 1000")
    ap.add_argument("--mode", default="auto", choices=["auto", "sharded", "replicas"], help="N > 1: one unit sharded over the ranks (default) or one unit per rank")
    ap.add_argument("--cpu-variants", type=int, default=0, help="size of the bounded CPU sample (0: the reference arm sizes it from --steps so that the run ends in minutes — 12000 variants per step for B up to 4 steps, 2857 at 20 steps; 3000 for the cpu_baseline leg; at most 120 for D)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timed-only", action="store_true", help="warm-up + timed steps only (for ncu launch lists); prints no bench line")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
