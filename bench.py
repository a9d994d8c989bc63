#!/usr/bin/env python
"""bench.py — headline benchmark: variant clusters genotyped / second (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our arm (libbtgpu, B200)
  python bench.py --impl reference ...                   the reference's own CPU code (oracle/_ref/btref)

One "step" = one pass of both hot paths over one synthetic batch of the named shape
(configs[1]: 1 sample, chr22-like SNV+indel candidate set, ~300k variants, k=55), bayestyper_b200/driver.py:
  k-mer match   findVariantClusterPaths (a7) -> path k-mer enumeration + exact table (a9, a12) -> genome scan (a10)
                -> sample k-mer stream (a11) -> classify + haplotype candidates (a13, a14) -> NB fit (a18)
  Gibbs         estimateNoise + estimateGenotypes: 20 chains x (100 + 250) iterations per cluster (a15-a23)
`value` times the step with the sample k-mer set, sample Bloom and reference already in HBM (CUDA events on the
library stream); `e2e` runs the same call with every input in host memory (H2D inside the timed region) and the
results read back.
N > 1 (torchrun): groups are independent -> every rank runs its own shard of the same size, no
data-path collective ("weak"); time = max over ranks.
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
WORKLOAD = "configs[1]: 1 sample, chr22-like SNV+indel candidate VCF (~300k variants), k=55, default Gibbs 20x(100+250)"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


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
    inp.prepare()
    return inp


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
    opt = driver.Options(random_seed=20190401)

    t0 = time.time()
    inp = build_batch(lib, rank, args.scale, dev)
    inp.make_resident(lib, opt)
    setup_s = time.time() - t0
    n_sample = int(inp.spectra_dev[0][0].shape[0])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM -------------------------------------------------------------
    info = None
    for _ in range(args.warmup):
        _, _, _, info = driver.genotype(inp, opt, resident=True)
    n_clusters = info["n_clusters"] if info else None
    barrier()
    lib.btg_launch_count_reset()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            _, _, res, info = driver.genotype(inp, opt, resident=True)
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
    launches = int(lib.btg_launch_count())
    n_clusters = info["n_clusters"]
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * n_clusters / (ms_per_step / 1e3)

    if args.timed_only:          # profiler runs (ncu launch list): the warm-up + timed steps only; not a bench line
        if rank == 0:
            print(json.dumps({"timed_only": True, "ms_per_step": ms_per_step, "value": value, "gpu_launches": launches, "n_clusters": n_clusters}), flush=True)
        inp.free(lib)
        return
    # ---- e2e: every input crosses the boundary from host memory inside the call --------------------------
    keys_d, counts_d = inp.spectra_dev[0]
    h_keys = torch.empty(keys_d.shape, dtype=torch.int64, pin_memory=True); h_keys.copy_(keys_d)
    h_counts = torch.empty(counts_d.shape, dtype=torch.uint8, pin_memory=True); h_counts.copy_(counts_d)
    nk, nb, nh = C.c_uint64(), C.c_uint64(), C.c_uint32()
    lib.btg_bloom_info(inp.blooms_dev[0], C.byref(nk), C.byref(nb), C.byref(nh))
    bloom_bytes = np.zeros((nb.value + 7) // 8, np.uint8)
    capi.check(lib.btg_bloom_download(inp.blooms_dev[0], capi.ptr(bloom_bytes), bloom_bytes.size), lib)
    host_inp = driver.Inputs(inp.chrom, inp.reference, inp.variants, inp.genders, spectra=[(h_keys, h_counts)],
                             blooms=[(bloom_bytes, nk.value, nb.value)], graphs=inp.graphs, regions=inp.regions)
    region_bytes = int(inp.region_buf_dev.numel())
    h2d = h_keys.numel() * 8 + h_counts.numel() + bloom_bytes.size + region_bytes
    torch.cuda.synchronize()
    _, unit, res, _ = driver.genotype(host_inp, opt, resident=False, want_unit=True)      # warm-up; also sizes the unit traffic
    from bayestyper_b200.unit import Unit as _Unit
    small = sum(v.nbytes for k, v in unit.a.items() if k not in _Unit.DEVICE_FIELDS)      # per-cluster / per-group descriptors cross (down, up);
    h2d += small                                                                          # the row-level arrays stay in HBM (btg_unit_upload_dev)
    d2h = sum(v.nbytes for v in res.values()) + small
    barrier()
    e2e_steps = max(1, min(args.steps, 2))
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        driver.genotype(host_inp, opt, resident=False)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t1) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_clusters / float(t.item())

    # ---- stage breakdown + roofline of the k-mer-match stream kernel (measured live, CUDA events) ---------
    stage_ms, roof = stage_breakdown(lib, inp, opt, stream, dev)

    peak, peak_src = measured_peaks()
    roof.update({"peak": peak, "unit": "GB/s", "frac": roof["achieved"] / peak, "peak_source": peak_src, "bound": "hbm", "traffic": None,
                 # no ncu capture of THIS launch; the same kernel on tools/prof_stream.py's 40.6 M records / 21.0 M keys moved 1.15 GB of DRAM
                 # traffic for 1.04 GB of algorithmic bytes (profiles/r1_stream_tiled_ncu_full.txt): no wasted re-reads
                 "traffic_reference": {"profile": "profiles/r1_stream_tiled_ncu_full.txt", "dram_bytes": 1150482000, "algorithmic_bytes": 1038973086,
                                       "workload": "tools/prof_stream.py: 40,620,727 records, 21,000,000 keys"}})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (Gibbs log-likelihoods) / u64 (k-mer hashing)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clusters_per_gpu": n_clusters, "variants_per_gpu": len(inp.variants), "samples": 1,
                   "reference_nt": len(inp.reference), "sample_kmers": n_sample, "path_kmers": info["n_path_kmers"],
                   "step": "findVariantClusterPaths -> path k-mer table -> genome scan -> sample k-mer stream -> classify/haplotype candidates -> NB fit -> estimateNoise -> estimateGenotypes",
                   "gibbs": "20 chains x (100 burn-in + 250 samples), k-mer subsampling 0.1",
                   "l2": "inputs exceed L2 (sample k-mer stream %.2f GB, sample Bloom %.0f MB)" % (n_sample * 17 / 1e9, nb.value / 8e6),
                   "parallelism": "groups sharded across ranks, no collective" if world > 1 else "1 GPU", "scale": args.scale},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roof,
        "stage_ms": stage_ms,
        "setup_s": setup_s,
    }
    if rank == 0:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args.cpu_variants, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    inp.free(lib)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def stage_breakdown(lib, inp, opt, stream, dev):
    """One extra (untimed for `value`) pass with a synchronisation after every stage, and the isolated timing of the
    stream kernel k_table_add_sample for the roofline object."""
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

    n_paths, mem = timed("findVariantClusterPaths", lambda: driver.find_variant_cluster_paths(lib, inp.graphs, inp.blooms_dev, opt))
    pipe = kmer_pipeline.KmerPipeline(inp.graphs, n_paths, mem, 1, inp.genders)
    timed("countPathKmers(enumerate+sort)", pipe.enumerate_path_kmers)
    timed("countInterclusterKmers(scan)", lambda: pipe.scan_buffer(inp.region_buf_dev, 2, 2, False))
    kd, cdv = inp.spectra_dev[0]
    timed("parseSampleKmers(stream)", lambda: pipe.add_sample(0, kd, cdv))
    unit = timed("classify+getHaplotypeCandidates", lambda: pipe.build_unit(multigroup_bloom=None, device_resident=True))
    nb = timed("NB fit (parameter k-mers)", lambda: driver.estimate_nb_parameters(pipe, inp.region_buf_dev, inp.spectra_dev, inp.genders, opt))
    cd = engine.CountDistribution(nb[0], nb[1])
    eng = timed("unit upload", lambda: engine.InferenceEngine(unit))
    gopts = U.default_opts(seed=opt.random_seed, min_frac=U.min_fraction_observed(nb[0], nb[1]))
    timed("estimateNoise", lambda: eng.estimate_noise(cd, gopts, want_trace=False))
    timed("estimateGenotypes", lambda: eng.estimate_genotypes(cd, gopts))
    eng.close(); cd.close()
    # roofline: the sample k-mer stream probing the exact path-k-mer table
    pipe.use_index()
    counts = torch.zeros_like(pipe.counts); rec = torch.zeros_like(pipe.has_record)
    torch.cuda.synchronize()
    n = kd.shape[0]
    for _ in range(3):
        capi.check(lib.btg_table_add_sample_kmers_dev(pipe.kw0.data_ptr(), pipe.kw1.data_ptr(), pipe.n_keys, kd.data_ptr(), cdv.data_ptr(), n, 1, 0, counts.data_ptr(), rec.data_ptr(), None), lib)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        capi.check(lib.btg_table_add_sample_kmers_dev(pipe.kw0.data_ptr(), pipe.kw1.data_ptr(), pipe.n_keys, kd.data_ptr(), cdv.data_ptr(), n, 1, 0, counts.data_ptr(), rec.data_ptr(), None), lib)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / reps
    hits = int(rec.sum())
    alg = n * 17 + pipe.n_keys * 16 + hits * 1
    roof = {"kernel": "k_table_add_sample (parseSampleKmers: sample k-mer stream probing the exact path-k-mer table, a11)",
            "achieved": alg / (ms / 1e3) / 1e9, "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms,
            "records": n, "table_keys": pipe.n_keys, "hits": hits,
            "bytes_model": "17 B per record + 16 B per path k-mer + S B per hit (SURVEY.md section 8d, merge-join form)"}
    return out, roof


# ------------------------------------------------------------------------------------------------
# the reference's own CPU path (oracle-R), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference(n_variants: int, threads: int):
    """Runs the reference's translation units (oracle/_ref/btref: cluster + genotype stage order) on a
    bounded sample of the same workload shape and reports its estimateGenotypes throughput."""
    from bayestyper_b200 import synth
    btref = ROOT / "oracle" / "_ref" / "btref"
    if not btref.exists():
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref/btref not built"}
    length = int(n_variants * 136)
    ref = synth.random_reference(length, 11)
    var = synth.make_variants(ref, n_variants, 12, 0.075, 0.075)
    g = synth.make_genotypes(len(var), 1, 13)
    w = synth.Workload("B-sample", "chr22", ref, var, g, ["F"])
    with tempfile.TemporaryDirectory() as td:
        synth.write_workdir(w, td, n_errors=50_000)
        t0 = time.time()
        subprocess.check_call([str(btref), "run", "--workdir", td, "--threads", str(threads), "--seed", "20190401"], stdout=subprocess.DEVNULL)
        wall = time.time() - t0
        tj = json.loads((Path(td) / "ref_out" / "timings.json").read_text())
    kmer_s = sum(tj.get(k, 0.0) for k in ("findVariantClusterPaths", "countPathMultigroupKmers", "countPathKmers", "countInterclusterKmers", "parseSampleKmers", "classifyPathKmers"))
    total_s = tj["estimateGenotypes"] + tj.get("estimateNoise", 0.0) + kmer_s
    return {"value": tj["clusters_genotyped"] / total_s, "unit": UNIT, "cores": threads, "kind": "reference", "step_s": total_s,
            "sample": f"{len(var)} variants / {tj['num_clusters']} clusters of the same chr22-like shape through the reference's own stages "
                      f"(estimateGenotypes {tj['estimateGenotypes']:.2f} s, estimateNoise {tj.get('estimateNoise', 0):.2f} s, k-mer stages {kmer_s:.2f} s, wall {wall:.1f} s)",
            "clusters": tj["clusters_genotyped"], "estimateGenotypes_s": tj["estimateGenotypes"], "kmer_stages_s": kmer_s,
            "clusters_per_s_gibbs_only": tj["clusters_genotyped"] / tj["estimateGenotypes"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_reference(args.cpu_variants, threads)
        if last["value"] is None:
            print(json.dumps({"impl": "reference", "unavailable": last["sample"]}))
            return
        if i >= args.warmup:
            vals.append(last)
    value = float(np.mean([v["clusters"] / v["step_s"] for v in vals]))
    ms = float(np.mean([v["step_s"] * 1e3 for v in vals]))
    cb = dict(last); cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 / u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": last["sample"], "threads": threads},
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the chr22-sized batch (tests use small values)")
    ap.add_argument("--cpu-variants", type=int, default=3000, help="size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timed-only", action="store_true", help="warm-up + timed steps only (for ncu launch lists); prints no bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
