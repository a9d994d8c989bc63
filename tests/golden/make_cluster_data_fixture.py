"""`<out>_cluster_data/` text files as the REFERENCE's `cluster` stage wrote them (oracle-R: oracle/_ref/btref run --cluster-only) for the
pipe_mixed_3s workload: intercluster_regions and the first 400 parameter k-mers.  oracle-R's iostreams shim does not compress, so the
plain text it leaves is gzip-compressed here, which is what the real reference writes (boost gzip_compressor).
Run in the build container (needs /root/reference):  python tests/golden/make_cluster_data_fixture.py"""
import gzip
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import synth  # noqa: E402
from tests.golden.make_fixtures import PIPE_WORKLOADS  # noqa: E402


def _text(p: Path) -> bytes:
    raw = p.read_bytes()
    return gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw


def main():
    w = PIPE_WORKLOADS["pipe_mixed_3s"]()
    with tempfile.TemporaryDirectory() as td:
        wd = synth.write_workdir(w, td, spectra=synth.sample_spectra(w, 4, 3000))
        subprocess.check_call([str(ROOT / "oracle" / "_ref" / "btref"), "run", "--workdir", str(wd), "--threads", "2", "--seed", "20190401", "--cluster-only"],
                              stdout=subprocess.DEVNULL)
        cd = Path(wd) / "ref_out" / "bayestyper_cluster_data"
        regions = _text(cd / "intercluster_regions.txt.gz")
        kmers = b"".join(_text(cd / "parameter_kmers.fa.gz").splitlines(keepends=True)[:401])
    out = ROOT / "tests" / "golden"
    with gzip.GzipFile(out / "cluster_data_mixed_3s.intercluster_regions.txt.gz", "wb", mtime=0) as f:
        f.write(regions)
    with gzip.GzipFile(out / "cluster_data_mixed_3s.parameter_kmers.fa.gz", "wb", mtime=0) as f:
        f.write(kmers)
    print("regions", regions.count(b"\n"), "parameter k-mers", kmers.count(b"\n") - 1)


if __name__ == "__main__":
    main()
