"""Golden VCFs written by the REFERENCE's own GenotypeWriter (oracle-R: oracle/_ref/btref compiles src/bayesTyper/GenotypeWriter.cpp
in place) on small seeded workloads: diploid and chrX (haploid males).  (A chrY-only genome makes the reference abort: a female sample has no genomic
k-mers to fit, CountDistribution.cpp:115; the no-genotype sample format is covered by a synthetic case in tests/test_vcf_writer.py.)  Run in the build
container (needs /root/reference):  python tests/golden/make_vcf_fixtures.py"""
import gzip
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import synth  # noqa: E402

def _nested_with_origins():
    """Deletions with variants inside; every third variant carries ids / ACO call-set origins of its own."""
    w = synth.nested_sv(5, 14_000, 2, seed=83, n_background=40, sv_len=(150, 500))
    for i, v in enumerate(w.variants):
        if i % 3 == 0:
            v.id = f"rs{1000 + i}"
            v.aco = [("gatk:platypus", "manta", "gatk")[(i // 3 + a) % 3] for a in range(len(v.alts))]
    return w


VCF_WORKLOADS = {
    "vcf_mixed_3s": lambda: synth.small_mixed(160, 16_000, 3, seed=81),
    "vcf_chrx_2s": lambda: synth.small_mixed(90, 9_000, 2, seed=82, chrom="chrX"),
    # deletions with variants inside: groups of several clusters (VCGS > 1, VCGR != VCR) and '*' alleles on dependent variants
    "vcf_nested_2s": _nested_with_origins,
}


def main(only=None):
    btref = ROOT / "oracle" / "_ref" / "btref"
    for name, mk in VCF_WORKLOADS.items():
        if only and name not in only:
            continue
        w = mk()
        with tempfile.TemporaryDirectory() as td:
            wd = synth.write_workdir(w, td, n_errors=2000)
            subprocess.check_call([str(btref), "run", "--workdir", str(wd), "--threads", "4", "--seed", "20190401"], stdout=subprocess.DEVNULL)
            txt = (Path(wd) / "ref_out" / "bayestyper.vcf").read_text()
            if name == "vcf_chrx_2s":   # the two parameter files of the same run (tests/test_vcf_writer.py::test_parameter_files_reproduce_the_reference)
                (ROOT / "tests" / "golden" / "params_chrx_2s_genomic.txt").write_text((Path(wd) / "ref_out" / "bayestyper_genomic_parameters.txt").read_text())
                with gzip.GzipFile(ROOT / "tests" / "golden" / "params_chrx_2s_noise.txt.gz", "wb", mtime=0) as f:
                    f.write((Path(wd) / "ref_out" / "bayestyper_noise_parameters.txt").read_bytes())
        txt = txt.replace(str(td), "/WORKDIR")          # the temporary directory appears in ##reference and the option lines
        out = ROOT / "tests" / "golden" / f"{name}.vcf.gz"
        with gzip.GzipFile(out, "wb", mtime=0) as f:
            f.write(txt.encode())
        body = [l for l in txt.splitlines() if not l.startswith("#")]
        print(name, "genders", w.genders, "records", len(body), "bytes", out.stat().st_size, "ploidy-0 samples", sum(l.count("\t:.:.:.:.:.:.") for l in body),
              "haploid GT", sum(bool(re.search(r"\t\d:", l)) for l in body))


if __name__ == "__main__":
    main(sys.argv[1:])
