"""Times the variants of the sample-stream kernel (BTG_STREAM_VARIANT, table.cu) on a synthetic table/stream of the
bench's configs[1] shape (21 M path k-mers, 47 M records of which 19.7 M hit) and checks every variant's table against
variant 0's.  Run on the GPU box:  python tools/stream_variants.py [variants...]   (one subprocess per variant: the
library reads the variable once)."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(variant: int, n_keys: int, n_rec: int, n_hit: int, S: int):
    import torch
    from bayestyper_b200 import capi
    lib = capi.load()
    capi.check(lib.btg_init(0), lib)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(7)

    def lexsort(hi, lo):
        o = torch.sort(lo, stable=True).indices
        o = o[torch.sort(hi[o], stable=True).indices]
        return hi[o].contiguous(), lo[o].contiguous()

    def rand_keys(n):
        hi = torch.randint(0, 1 << 46, (n,), generator=g, device=dev, dtype=torch.int64)
        lo = torch.randint(-(1 << 62), 1 << 62, (n,), generator=g, device=dev, dtype=torch.int64) * 2 + torch.randint(0, 2, (n,), generator=g, device=dev, dtype=torch.int64)
        return hi, lo

    khi, klo = lexsort(*rand_keys(n_keys))
    # a few thousand keys sharing key_hi with their predecessor (alleles that differ late in the k-mer)
    tie = torch.randint(1, n_keys, (n_keys // 500,), generator=g, device=dev)
    khi[tie] = khi[tie - 1]
    khi, klo = lexsort(khi, klo)
    sel = torch.randperm(n_keys, generator=g, device=dev)[:n_hit]
    ohi, olo = rand_keys(n_rec - n_hit)
    rhi, rlo = lexsort(torch.cat([khi[sel], ohi]), torch.cat([klo[sel], olo]))
    rec_kmers = torch.empty((n_rec, 2), dtype=torch.int64, device=dev)
    capi.check(lib.btg_table_keys_to_kmers_dev(rlo.data_ptr(), rhi.data_ptr(), n_rec, rec_kmers.data_ptr(), None), lib)
    rec_counts = torch.randint(1, 256, (n_rec,), generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
    lut_bits = 24
    cnt = torch.bincount(khi >> (46 - lut_bits), minlength=1 << lut_bits)
    lut = torch.zeros((1 << lut_bits) + 1, dtype=torch.int64, device=dev)
    lut[1:] = torch.cumsum(cnt, 0)
    capi.check(lib.btg_table_set_index_dev(lut.data_ptr(), lut_bits), lib)
    counts = torch.zeros((n_keys, S), dtype=torch.uint8, device=dev)
    has = torch.zeros(n_keys, dtype=torch.uint8, device=dev)

    def launch(sample=0):
        capi.check(lib.btg_table_add_sample_kmers_dev(klo.data_ptr(), khi.data_ptr(), n_keys, rec_kmers.data_ptr(), rec_counts.data_ptr(), n_rec, S, sample,
                                                      counts.data_ptr(), has.data_ptr(), None), lib)

    launch(S - 1)
    torch.cuda.synchronize()
    digest = hashlib.sha256(counts.cpu().numpy().tobytes() + has.cpu().numpy().tobytes()).hexdigest()[:16]
    hits = int(has.sum())
    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(lib.btg_get_stream(), device=dev)   # the library launches on its own stream
    torch.cuda.synchronize()
    best = []
    for _ in range(3):
        e0.record(st)
        for _ in range(5):
            launch()
        e1.record(st)
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / 5)
    alg = n_rec * 17 + n_keys * 16 + hits
    ms = min(best)
    print(json.dumps({"variant": variant, "S": S, "ms": ms, "all_ms": best, "GBps": alg / ms / 1e6, "hits": hits, "digest": digest}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(*[int(x) for x in sys.argv[2:7]])
    else:
        variants = [int(x) for x in sys.argv[1:]] or [0, 1, 2]
        for S in (1, 3):
            for v in variants:
                env = dict(os.environ, BTG_STREAM_VARIANT=str(v))
                subprocess.run([sys.executable, __file__, "--one", str(v), "20987263", "47373149", "19721812", str(S)], env=env, check=False)
