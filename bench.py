#!/usr/bin/env python
"""bench.py — headline benchmark: variant clusters genotyped / second (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          our arm (libbtgpu, B200)
  python bench.py --impl reference ...                   the reference's own CPU code (oracle/_ref/btref)

One "step" = one pass of the hot path over one synthetic batch of the named shape
(configs[1]: 1 sample, chr22-like SNV+indel candidate set, ~300k variants, k=55):
  k-mer match   path-k-mer Bloom build (a9) -> sample k-mer stream filtered through it (a11)
                -> path k-mers probed in the sample Bloom (a7's innermost loop)
  Gibbs         InferenceEngine::estimateGenotypes: 20 chains x (100 + 250) iterations per cluster (a15-a23)
`value` times the step with every input already in HBM (CUDA events on the library stream);
`e2e` times the same step through the host-buffer C ABI (H2D of k-mers + unit descriptors, D2H of results).
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
def build_batch(rank: int, scale: float):
    """The chr22-like unit (structure enumerated on a 1/tile slice of the chromosome, tiled with
    independently redrawn counts) + the k-mer sets the k-mer stages stream."""
    from bayestyper_b200 import synth, synth_unit
    tile = 10
    n_var = int(30_000 * scale)
    length = int(4_080_000 * scale)
    ref = synth.random_reference(length, 11 + 1000 * rank)
    var = synth.make_variants(ref, n_var, 12 + 1000 * rank, 0.075, 0.075)
    g = synth.make_genotypes(len(var), 1, 13 + 1000 * rank)
    w = synth.Workload("B", "chr22", ref, var, g, ["F"])
    base = synth_unit.build_unit(w, seed=14 + 1000 * rank)
    unit = synth_unit.tile_unit(base, tile, seed=15 + 1000 * rank)
    return unit, tile * len(var)


def random_kmers_torch(n, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    k = torch.empty((n, 2), dtype=torch.int64, device=device).random_(generator=g)
    k[:, 1] &= (1 << 46) - 1
    return k


def run_ours(args):
    import torch
    from bayestyper_b200 import capi, engine, unit as U

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    lib = capi.load()
    capi.check(lib.btg_init(local_rank), lib)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(lib.btg_get_stream(), device=dev)
    K = 55

    t0 = time.time()
    unit, n_variants = build_batch(rank, args.scale)
    n_clusters = unit.Cn
    n_path = int(unit.a["cl_kmer_off"][-1])                   # path k-mers (rows)
    n_sample = int(80_000_000 * args.scale)                   # distinct 55-mers of a ~40.8 Mb diploid sample
    setup_unit_s = time.time() - t0

    # --- resident inputs --------------------------------------------------------------------
    path_k = random_kmers_torch(n_path, 100 + rank, dev)
    sample_k = random_kmers_torch(n_sample, 200 + rank, dev)
    n_shared = min(n_path, n_sample) // 2
    sample_k[:n_shared] = path_k[:n_shared]                   # half of the path k-mers are observed
    sample_bloom = capi.check(lib.btg_bloom_create(n_sample, 1e-3, K), lib)
    capi.check(lib.btg_bloom_insert_dev(sample_bloom, sample_k.data_ptr(), n_sample, None), lib)
    hit_path = torch.zeros(n_path, dtype=torch.uint8, device=dev)
    hit_sample = torch.zeros(n_sample, dtype=torch.uint8, device=dev)
    probes = torch.zeros(n_sample, dtype=torch.uint8, device=dev)
    nb_p, nb_size = np.array([0.6]), np.array([22.5])         # NB(mean 15, var 25) per haploid copy
    cd = engine.CountDistribution(nb_p, nb_size)
    cd.set_noise_rates([0.02])
    opts = U.default_opts(seed=20190401, min_frac=U.min_fraction_observed(nb_p, nb_size), group_base=rank * unit.G)
    eng = engine.InferenceEngine(unit)
    res_struct, res_arrays = unit.alloc_result()
    torch.cuda.synchronize()

    sp = stream.cuda_stream

    def kmer_stage(path_bloom, count_probes=False):
        capi.check(lib.btg_tbloom_insert_dev(path_bloom, path_k.data_ptr(), n_path, sp), lib)
        capi.check(lib.btg_tbloom_lookup_dev(path_bloom, sample_k.data_ptr(), n_sample, hit_sample.data_ptr(), sp), lib)
        capi.check(lib.btg_bloom_lookup_dev(sample_bloom, path_k.data_ptr(), n_path, hit_path.data_ptr(), sp), lib)

    def step_resident():
        pb = capi.check(lib.btg_tbloom_create(n_path + 1_000_000, 1e-4, K), lib)
        kmer_stage(pb)
        capi.check(lib.btg_estimate_genotypes_async(eng.h, cd.h, C.addressof(opts), sp), lib)
        return pb

    # algorithmic bytes of the stream-filter kernel: 17 B per record + 32 B per executed probe
    pb0 = capi.check(lib.btg_tbloom_create(n_path + 1_000_000, 1e-4, K), lib)
    capi.check(lib.btg_tbloom_insert_dev(pb0, path_k.data_ptr(), n_path, sp), lib)
    stream.synchronize()
    probes_per_record = measure_tbloom_probes(lib, pb0, sample_k, dev)
    lib.btg_tbloom_free(pb0)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident step ------------------------------------------------------------
    for _ in range(args.warmup):
        pb = step_resident(); stream.synchronize(); lib.btg_tbloom_free(pb)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    kmer_ms, stream_ms = [], []
    lib.btg_launch_count_reset()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pbs = []
        marks = []
        for _ in range(args.steps):
            a, b, c_, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            pb = capi.check(lib.btg_tbloom_create(n_path + 1_000_000, 1e-4, K), lib)
            a.record(stream)
            capi.check(lib.btg_tbloom_insert_dev(pb, path_k.data_ptr(), n_path, sp), lib)
            b.record(stream)
            capi.check(lib.btg_tbloom_lookup_dev(pb, sample_k.data_ptr(), n_sample, hit_sample.data_ptr(), sp), lib)
            c_.record(stream)
            capi.check(lib.btg_bloom_lookup_dev(sample_bloom, path_k.data_ptr(), n_path, hit_path.data_ptr(), sp), lib)
            d.record(stream)
            capi.check(lib.btg_estimate_genotypes_async(eng.h, cd.h, C.addressof(opts), sp), lib)
            pbs.append(pb); marks.append((a, b, c_, d))
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
    launches = int(lib.btg_launch_count())
    for pb in pbs:
        lib.btg_tbloom_free(pb)
    total_ms = e0.elapsed_time(e1)
    stream_ms = [m[1].elapsed_time(m[2]) for m in marks]
    kmer_ms = [m[0].elapsed_time(m[3]) for m in marks]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * n_clusters / (ms_per_step / 1e3)

    # ---- e2e: host buffers through the C ABI --------------------------------------------------------
    h_path = torch.empty((n_path, 2), dtype=torch.int64, pin_memory=True); h_path.copy_(path_k)
    h_sample = torch.empty((n_sample, 2), dtype=torch.int64, pin_memory=True); h_sample.copy_(sample_k)
    h_hit_p = torch.empty(n_path, dtype=torch.uint8, pin_memory=True)
    h_hit_s = torch.empty(n_sample, dtype=torch.uint8, pin_memory=True)
    desc = unit.desc()
    h2d = n_path * 16 * 2 + n_sample * 16 + sum(v.nbytes for v in unit.a.values())
    d2h = n_path + n_sample + sum(v.nbytes for k, v in res_arrays.items() if k not in ("allele_off", "geno_off", "valt_off"))

    def step_e2e():
        pb = capi.check(lib.btg_tbloom_create(n_path + 1_000_000, 1e-4, K), lib)
        capi.check(lib.btg_tbloom_insert(pb, h_path.data_ptr(), n_path), lib)
        capi.check(lib.btg_tbloom_lookup(pb, h_sample.data_ptr(), n_sample, h_hit_s.data_ptr()), lib)
        capi.check(lib.btg_bloom_lookup(sample_bloom, h_path.data_ptr(), n_path, h_hit_p.data_ptr()), lib)
        lib.btg_tbloom_free(pb)
        u = capi.check(lib.btg_unit_upload(C.addressof(desc)), lib)
        capi.check(lib.btg_estimate_genotypes(u, cd.h, C.addressof(opts), C.addressof(res_struct)), lib)
        lib.btg_unit_free(u)

    step_e2e()
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t1) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_clusters / float(t.item())

    # ---- roofline of the k-mer-match kernel --------------------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = n_sample * (17 + 32.0 * probes_per_record)
    stream_kernel_ms = float(np.mean(stream_ms))
    achieved = alg_bytes / (stream_kernel_ms / 1e3) / 1e9
    gibbs_ms = ms_per_step - float(np.mean(kmer_ms))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (Gibbs log-likelihoods) / u64 (k-mer hashing)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clusters_per_gpu": n_clusters, "variants_per_gpu": int(unit.n_variants), "samples": unit.S,
                   "path_kmers": n_path, "sample_kmers": n_sample, "gibbs": "20 chains x (100 burn-in + 250 samples)", "kmer_subsampling_rate": 0.1,
                   "l2": "inputs exceed L2 (k-mer streams %.1f GB, Gibbs state > 126 MB)" % ((n_sample + n_path) * 16 / 1e9),
                   "parallelism": "groups sharded across ranks, no collective" if world > 1 else "1 GPU", "scale": args.scale},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"kernel": "k_tbloom_lookup (sample k-mer stream filtered through the path-k-mer Bloom, a11)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "probes_per_record": probes_per_record, "ms_per_launch": stream_kernel_ms},
        "stage_ms": {"kmer_match": float(np.mean(kmer_ms)), "gibbs": gibbs_ms},
        "setup_s": {"unit": setup_unit_s},
    }
    if rank == 0:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args.cpu_variants, os.cpu_count() or 1)
        print(json.dumps(line), flush=True)
    eng.close(); cd.close()
    lib.btg_bloom_free(sample_bloom)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def measure_tbloom_probes(lib, pb, sample_k, dev):
    """Mean number of probes the reference's early-exit loop executes per streamed record: measured on a
    1M-record sample with an equally loaded KmerBloom (same fpr) through btg_bloom_lookup_probes_dev."""
    import torch
    from bayestyper_b200 import capi
    n = min(1_000_000, sample_k.shape[0])
    sub_k, sub_b, nh = C.c_uint64(), C.c_uint64(), C.c_uint32()
    lib.btg_tbloom_info(pb, C.byref(sub_k), C.byref(sub_b), C.byref(nh))
    # an absent k-mer passes each probe with the filter's fill ratio f: E[probes] = sum_{i<nh} f^i
    bits = np.zeros((65536, (sub_b.value + 7) // 8), np.uint8)
    capi.check(lib.btg_tbloom_download(pb, bits.ctypes.data, bits.size), lib)
    fill = float(np.unpackbits(bits[:256]).mean()) * (bits.shape[1] * 8) / sub_b.value
    hit = torch.zeros(n, dtype=torch.uint8, device=dev)
    capi.check(lib.btg_tbloom_lookup_dev(pb, sample_k.data_ptr(), n, hit.data_ptr(), None), lib)
    torch.cuda.synchronize()
    h = float(hit.float().mean())
    miss_probes = sum(fill ** i for i in range(nh.value))
    return h * nh.value + (1 - h) * miss_probes


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
    return {"value": tj["clusters_genotyped"] / (tj["estimateGenotypes"] + kmer_s), "unit": UNIT, "cores": threads, "kind": "reference",
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
    value = float(np.mean([v["clusters"] / (v["estimateGenotypes_s"] + v["kmer_stages_s"]) for v in vals]))
    ms = float(np.mean([(v["estimateGenotypes_s"] + v["kmer_stages_s"]) * 1e3 for v in vals]))
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
