"""`bayesTyperTools makeBloom` as a device build (host/btkmc makebloom): the KMC database is read on the
host (include/btgpu_kmc.hpp), the k-mers are inserted by k_bloom_insert, and the <prefix>.bloomMeta / .bloomData it writes must be
what the reference's KmerBloom writes for the same k-mers: byte for byte the oracle's filter (kmer_oracle.c is pinned to the
reference's BloomFilter by tests/test_oracle_kmer.py), and loadable by btg_bloom_load."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import capi, kmcio
from tests import _oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("writer", ["kmc1", "kmc2"])
def test_makebloom_on_the_device(btg, tmp_path, writer):
    exe = ROOT / "host" / "btkmc"
    if not exe.exists():
        pytest.skip("host/btkmc not built")
    raw = O.random_kmers(6000, 5)
    canon = np.zeros_like(raw)
    L = O.load()
    for i in range(len(raw)):                                                  # a KMC database holds canonical k-mers (both strands counted)
        L.bto_canonical(raw[i], 55, canon[i])
    km = np.unique(canon, axis=0)
    counts = (np.arange(len(km)) % 7 + 1).astype(np.uint32)
    prefix = tmp_path / "sample"
    (kmcio.write_kmc1 if writer == "kmc1" else kmcio.write_kmc2)(prefix, km, counts)
    out = subprocess.run([str(exe), "makebloom", str(prefix), "0.001"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    n, bits, k = (int(x) for x in Path(str(prefix) + ".bloomMeta").read_text().split())
    assert (n, k) == (len(km), 55)
    data = np.fromfile(str(prefix) + ".bloomData", np.uint8)
    assert data.size == (bits + 7) // 8
    nh = L.bto_bloom_num_hashes(bits, n)                                       # recomputed from the meta (KmerBloom.cpp:140-146)
    listed, _, _ = kmcio.read_kmc(prefix)
    want = O.bloom_build(np.ascontiguousarray(listed), bits, nh)
    assert (data == want).all()
    lib = btg
    b = capi.check(lib.btg_bloom_load(str(prefix).encode(), 55), lib)
    hit = np.zeros(len(km), np.uint8)
    capi.check(lib.btg_bloom_lookup(b, capi.ptr(np.ascontiguousarray(km)), len(km), capi.ptr(hit)), lib)
    assert hit.all()
    lib.btg_bloom_free(b)
