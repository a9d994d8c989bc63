"""Samples file and chromosome ploidy, the two small host-side inputs of `bayesTyper genotype` that decide which count
distribution a (sample, contig) pair uses on the device.

`read_samples` follows Sample::Sample (src/bayesTyper/Sample.cpp:38-70); `ChromosomePloidy` follows
src/bayesTyper/ChromosomePloidy.cpp:40-185: without a ploidy file every contig is diploid except X / chrX (males haploid) and
Y / chrY (females 0, males haploid), names compared case-insensitively; with `--chromosome-ploidy-file` every non-decoy contig
needs a `<contig> \\t <female> \\t <male>` line with values 0-2 (the name is matched exactly there).
"""
from __future__ import annotations


def read_samples(path) -> list:
    """[(sample id, 'F' | 'M', KMC output prefix)] in file order."""
    out = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            t = line.split("\t")
            if len(t) != 3:
                raise ValueError(f'line "{line}" in the samples file should contain three tab-separated columns '
                                 "(<Sample ID>, <Gender> & <KMC Output Prefix>)")
            if t[1] in ("F", "Female"):
                gender = "F"
            elif t[1] in ("M", "Male"):
                gender = "M"
            else:
                raise ValueError(f'gender (column two) in line "{line}" in the samples file should be either "F" (Female) or "M" (Male)')
            out.append((t[0], gender, t[2]))
    return out


def _is_female(g) -> bool:
    return g in ("F", "Female", 0)


class ChromosomePloidy:
    """(female, male) ploidy per contig and the per-sample ploidy that follows from the genders."""

    def __init__(self, contigs, genders, ploidy_file=None, decoys=()):
        self.genders = list(genders)
        self._gender_ploidy = {}
        decoys = set(decoys)
        if not ploidy_file:
            for name in contigs:
                if name in decoys:
                    continue
                low = name.lower()
                if low in ("x", "chrx"):
                    self._gender_ploidy[name] = (2, 1)
                elif low in ("y", "chry"):
                    self._gender_ploidy[name] = (0, 1)
                else:
                    self._gender_ploidy[name] = (2, 2)
            return
        table = {}
        with open(ploidy_file) as f:
            for line in f:
                line = line.rstrip("\n")
                t = line.split("\t")
                if len(t) != 3:
                    raise ValueError(f'line "{line}" in the chromosome ploidy file should contain three tab-separated columns '
                                     "(<Chromosome name>, <Female Ploidy> & <Male Ploidy>)")
                female, male = int(t[1]), int(t[2])
                if t[0] in table:
                    raise ValueError(f'chromosome (column one) in line "{line}" appears multiple times in the chromosome ploidy file')
                if not 0 <= female <= 2 or not 0 <= male <= 2:
                    raise ValueError(f'ploidy in line "{line}" in the chromosome ploidy file should be between zero and two')
                table[t[0]] = (female, male)
        for name in contigs:
            if name in decoys:
                continue
            if name not in table:
                raise ValueError(f'chromosome "{name}" in reference genome does not appear in the chromosome ploidy file')
            self._gender_ploidy[name] = table[name]

    def gender_ploidy(self, contig: str):
        """ChromosomePloidy::getGenderPloidy: (female, male)."""
        return self._gender_ploidy[contig]

    def sample_ploidy(self, contig: str) -> list:
        """ChromosomePloidy::getSamplePloidy: one ploidy per sample."""
        female, male = self._gender_ploidy[contig]
        return [female if _is_female(g) else male for g in self.genders]
