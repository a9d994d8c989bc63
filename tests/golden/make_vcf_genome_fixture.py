"""Golden VCF written by the REFERENCE's GenotypeWriter (oracle-R) for a genome of several contigs: chr2 listed before chr1 in the
FASTA (records follow genome order, not name order), chrX (haploid male), a contig without variants, and a decoy contig.
Run in the build container (needs /root/reference):  python tests/golden/make_vcf_genome_fixture.py"""
import gzip
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bayestyper_b200 import synth  # noqa: E402

GENDERS = ["F", "M"]


def genome_workload():
    """(genome: contig -> Workload in FASTA order, contig without variants, decoys: name -> sequence)."""
    parts = {
        "chr2": synth.small_mixed(70, 8_000, 2, seed=91, chrom="chr2"),
        "chr1": synth.nested_sv(3, 9_000, 2, seed=92, n_background=30, sv_len=(150, 400), chrom="chr1"),
        "chrX": synth.small_mixed(50, 6_000, 2, seed=93, chrom="chrX"),
    }
    empty = ("chrUn", synth.random_reference(900, 94))
    decoys = {"decoy1": synth.random_reference(1_500, 95)}
    return parts, empty, decoys


def sample_spectra(parts, decoys, n_errors=1500):
    out = []
    for s, gender in enumerate(GENDERS):
        haps = []
        for name, w in parts.items():
            n_hap = 1 if (gender == "M" and name == "chrX") else 2
            haps += [synth.apply_variants(w.reference, w.variants, w.genotypes[s, :, h]) for h in range(n_hap)]
        for seq in decoys.values():
            haps += [seq, seq]
        out.append(synth.sample_kmer_counts(haps, 400 + 17 * s, 15.0, 25.0, n_errors))
    return out


def write_workdir(wd: Path, parts, empty, decoys, spectra):
    with open(wd / "genome.fa", "wb") as f:
        for name, seq in [(n, w.reference) for n, w in parts.items()] + [empty]:
            f.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + b"\n")
    with open(wd / "decoy.fa", "wb") as f:
        for name, seq in decoys.items():
            f.write(b">" + name.encode() + b"\n" + seq + b"\n")
    with open(wd / "variants.vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for name, w in parts.items():
            for i, v in enumerate(w.variants):
                f.write(f"{name}\t{v.pos + 1}\t{name}_{i}\t{v.ref.decode()}\t{','.join(a.decode() for a in v.alts)}\t.\t.\t.\n")
    with open(wd / "samples.tsv", "w") as f:
        for s, (km, ct) in enumerate(spectra):
            f.write(f"S{s + 1}\t{GENDERS[s]}\t{wd / f'S{s + 1}'}\n")
            synth.write_kmer_file(str(wd / f"S{s + 1}.kmers.bin"), km, ct)


def main():
    parts, empty, decoys = genome_workload()
    spectra = sample_spectra(parts, decoys)
    with tempfile.TemporaryDirectory() as td:
        wd = Path(td)
        write_workdir(wd, parts, empty, decoys, spectra)
        subprocess.check_call([str(ROOT / "oracle" / "_ref" / "btref"), "run", "--workdir", str(wd), "--threads", "4", "--seed", "20190401",
                               "--decoy-file", str(wd / "decoy.fa")], stdout=subprocess.DEVNULL)
        txt = (wd / "ref_out" / "bayestyper.vcf").read_text().replace(str(td), "/WORKDIR")
    with gzip.GzipFile(ROOT / "tests" / "golden" / "vcf_genome_2s.vcf.gz", "wb", mtime=0) as f:
        f.write(txt.encode())
    body = [l.split("\t") for l in txt.splitlines() if not l.startswith("#")]
    print("records", len(body), "contig order", list(dict.fromkeys(t[0] for t in body)), "header contigs", [l for l in txt.splitlines() if l.startswith("##contig")])


if __name__ == "__main__":
    main()
