"""Minimal reader of BayesTyper genotype VCFs (FORMAT GT:GQ:GPP:APP:NAK:FAK:MAC:SAF)."""
from __future__ import annotations

import gzip


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
