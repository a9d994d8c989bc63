"""Readers of the files either side of the path: the candidate-variant VCF and genome FASTA that `bayesTyper cluster` takes, and
BayesTyper's genotype VCFs (FORMAT GT:GQ:GPP:APP:NAK:FAK:MAC:SAF)."""
from __future__ import annotations

import dataclasses
import gzip


@dataclasses.dataclass
class Candidate:
    """One line of the candidate set, as VariantFileParser keeps it (VariantFileParser.cpp:140-167,296-322)."""
    pos: int                    # 0-based position of REF[0]
    ref: bytes
    alts: list                  # list[bytes], a trailing b"*" kept
    id: str = "."
    aco: list = None            # per alternative allele call-set origin (INFO ACO=...), or None


def read_candidates(path) -> dict:
    """Candidate variants per contig, in file order (contig -> list[Candidate]).  Like the reference, only CHROM POS ID REF ALT and
    the INFO column are looked at, `.vcf` and `.vcf.gz` are accepted and the column header line is required
    (VariantFileParser.cpp:67-167); ACO is the only INFO attribute read (getInfoAttributeString, :547-561)."""
    path = str(path)
    if not (path.endswith(".vcf") or path.endswith(".vcf.gz")):
        raise ValueError("variant file needs to end in .vcf or .vcf.gz")
    out: dict = {}
    seen_header = False
    with (gzip.open if path.endswith(".gz") else open)(path, "rb") as f:
        for line in f:
            if line.startswith(b"#"):
                if line.startswith(b"#CHROM"):
                    if line.count(b"\t") + 1 < 8:
                        raise ValueError("variant file header has fewer than 8 columns")
                    seen_header = True
                continue
            if not seen_header:
                raise ValueError("variant file has no #CHROM header line")
            t = line.rstrip(b"\r\n").split(b"\t")
            if len(t) < 8:
                raise ValueError(f"variant line with fewer than 8 columns: {line[:60]!r}")
            if b"," in t[3]:
                raise ValueError("REF holds several alleles")
            alts = t[4].split(b",")
            aco = None
            for kv in t[7].split(b";"):
                if kv.startswith(b"ACO="):
                    aco = kv[4:].decode().split(",")
                    if len(aco) != len(alts):
                        raise ValueError(f"ACO lists {len(aco)} origins for {len(alts)} alternative alleles at {t[0].decode()}:{t[1].decode()}")
                    break
            out.setdefault(t[0].decode(), []).append(Candidate(int(t[1]) - 1, t[3], alts, t[2].decode(), aco))
    return out


def read_fasta(path) -> dict:
    """Contig name (up to the first blank or tab) -> sequence bytes, in file order (Chromosomes::parseFasta, Chromosomes.cpp:72-117)."""
    out: dict = {}
    name = None
    parts: list = []
    with (gzip.open if str(path).endswith(".gz") else open)(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    out[name] = b"".join(parts)
                name = line[1:].replace(b"\t", b" ").split(b" ")[0].decode()
                if not name or name in out:
                    raise ValueError(f"empty or repeated contig name in {path}")
                parts = []
            else:
                if name is None:
                    raise ValueError(f"{path} does not start with a '>' line")
                parts.append(line)
    if name is not None:
        out[name] = b"".join(parts)
    return out


def read_vcf(path):
    op = gzip.open if str(path).endswith(".gz") else open
    rows = []
    samples = []
    with op(path, "rt") as f:
        for line in f:
            if line.startswith("##"):
                continue
            t = line.rstrip("\n").split("\t")
            if line.startswith("#CHROM"):
                samples = t[9:]
                continue
            info = dict(kv.split("=", 1) if "=" in kv else (kv, "") for kv in t[7].split(";"))
            fmt = t[8].split(":")
            srec = []
            for s in t[9:]:
                d = dict(zip(fmt, s.split(":")))
                rec = {"GT": d["GT"]}
                if "GQ" in d and d["GQ"] != ".":
                    rec["GQ"] = int(d["GQ"])
                for k in ("GPP", "APP", "NAK", "FAK", "MAC"):
                    if k in d and d[k] != ".":
                        rec[k] = [float(x) for x in d[k].split(",")]
                if "SAF" in d and d["SAF"] != ".":
                    rec["SAF"] = [int(x) for x in d["SAF"].split(",")]
                srec.append(rec)
            rows.append({"chrom": t[0], "pos": int(t[1]), "id": t[2], "ref": t[3], "alt": t[4].split(","), "qual": t[5],
                         "filter": t[6], "info": info, "samples": srec})
    return samples, rows
