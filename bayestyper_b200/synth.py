"""Seeded synthetic inputs of the shapes BASELINE.json names.

Produces what the two commands consume: a reference FASTA, a candidate VCF,
samples.tsv and, per sample, the (canonical 55-mer, count) set a KMC database
would hold (flat binary `<prefix>.kmers.bin`: u64 n, n x 2 x u64 packed k-mers in
the boundary layout, n x u8 counts).  Pure numpy — independent of both the CUDA
kernels and the oracle, so it can feed either side.
"""
from __future__ import annotations

import dataclasses
from pathlib import Path

import numpy as np

K = 55
_ACGT = np.frombuffer(b"ACGT", np.uint8)
_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i


@dataclasses.dataclass
class Variant:
    pos: int          # 0-based position of REF[0]
    ref: bytes
    alts: list        # list[bytes]
    id: str = None    # VCF ID (write_workdir writes v<index> when unset)
    aco: list = None  # per alternative allele call-set origin (INFO ACO), optional


@dataclasses.dataclass
class Workload:
    name: str
    chrom: str
    reference: bytes
    variants: list            # list[Variant], sorted, non-overlapping
    genotypes: np.ndarray     # (S, n_variants, 2) allele index per haplotype
    genders: list             # 'F'/'M' per sample
    depth_mean: float = 15.0  # haploid k-mer count mean
    depth_var: float = 25.0


def random_reference(length: int, seed: int, n_prefix: int = 0) -> bytes:
    rng = np.random.default_rng(seed)
    s = _ACGT[rng.integers(0, 4, length)]
    if n_prefix:
        s[:n_prefix] = ord("N")
    return s.tobytes()


def make_variants(reference: bytes, n: int, seed: int, frac_del: float = 0.0, frac_ins: float = 0.0,
                  max_indel: int = 50, lo: int | None = None, hi: int | None = None) -> list:
    """n candidate variants at distinct positions; SNV ALT = uniform other base;
    indel lengths ~ Geometric(0.3) capped at max_indel, VCF-style anchor base.
    Variants overlapping an earlier deletion's span (or an N) are dropped."""
    rng = np.random.default_rng(seed)
    ref = np.frombuffer(reference, np.uint8)
    L = len(ref)
    lo = K if lo is None else lo
    hi = L - K if hi is None else hi
    pos = np.sort(rng.choice(np.arange(lo, hi), size=n, replace=False))
    kind = rng.random(n)
    lens = np.minimum(rng.geometric(0.3, n), max_indel)
    out = []
    blocked_until = -1
    for p, u, ln in zip(pos.tolist(), kind.tolist(), lens.tolist()):
        if p <= blocked_until:
            continue
        if u < frac_del:
            if p + 1 + ln + K >= L or (_CODE[ref[p:p + 1 + ln]] > 3).any():
                continue
            out.append(Variant(p, ref[p:p + 1 + ln].tobytes(), [ref[p:p + 1].tobytes()]))
            blocked_until = p + ln
        elif u < frac_del + frac_ins:
            if _CODE[ref[p]] > 3:
                continue
            ins = _ACGT[rng.integers(0, 4, ln)].tobytes()
            out.append(Variant(p, ref[p:p + 1].tobytes(), [ref[p:p + 1].tobytes() + ins]))
            blocked_until = p
        else:
            if _CODE[ref[p]] > 3:
                continue
            alt = _ACGT[(int(_CODE[ref[p]]) + int(rng.integers(1, 4))) % 4]
            out.append(Variant(p, ref[p:p + 1].tobytes(), [bytes([alt])]))
            blocked_until = p
    return out


def make_genotypes(n_variants: int, n_samples: int, seed: int, probs=(0.4, 0.4, 0.2), allele_freq=None) -> np.ndarray:
    """(S, n_variants, 2) alt-allele indicator per haplotype.  Either fixed
    hom-ref/het/hom-alt probabilities or Hardy-Weinberg draws from allele_freq."""
    rng = np.random.default_rng(seed)
    g = np.zeros((n_samples, n_variants, 2), np.uint8)
    if allele_freq is not None:
        g[:] = rng.random((n_samples, n_variants, 2)) < allele_freq[None, :, None]
        return g
    u = rng.random((n_samples, n_variants))
    het = (u >= probs[0]) & (u < probs[0] + probs[1])
    hom = u >= probs[0] + probs[1]
    side = rng.integers(0, 2, (n_samples, n_variants))
    g[..., 0] = hom | (het & (side == 0))
    g[..., 1] = hom | (het & (side == 1))
    return g


def apply_variants(reference: bytes, variants: list, alleles: np.ndarray) -> bytes:
    """Haplotype sequence with allele index alleles[i] (0 = REF) at variants[i]."""
    parts = []
    cur = 0
    for v, a in zip(variants, alleles.tolist()):
        if a == 0 or v.pos < cur:      # inside an applied deletion: the allele is missing on this haplotype
            continue
        parts.append(reference[cur:v.pos])
        parts.append(v.alts[a - 1])
        cur = v.pos + len(v.ref)
    parts.append(reference[cur:])
    return b"".join(parts)


# ---- numpy canonical k-mer enumeration (third, independent implementation) ------------------
def canonical_kmers(seq: bytes) -> np.ndarray:
    """(n, 2) uint64 canonical 55-mers (boundary layout) of every all-ACGT window, in order."""
    c = _CODE[np.frombuffer(seq, np.uint8)].astype(np.uint64)
    L = len(c)
    if L < K:
        return np.zeros((0, 2), np.uint64)
    n = L - K + 1
    bad = np.concatenate([[0], np.cumsum(c > 3)])
    valid = (bad[K:] - bad[:-K]) == 0
    c = np.where(c > 3, 0, c)
    fw0 = np.zeros(n, np.uint64); fw1 = np.zeros(n, np.uint64)     # boundary words
    rc0 = np.zeros(n, np.uint64); rc1 = np.zeros(n, np.uint64)
    fhi = np.zeros(n, np.uint64); flo = np.zeros(n, np.uint64)     # MSB-first, for the compare
    rhi = np.zeros(n, np.uint64); rlo = np.zeros(n, np.uint64)
    three = np.uint64(3)
    for i in range(K):
        ci = c[i:i + n]
        cc = three - ci
        j = K - 1 - i                      # index of comp(nt_i) in the reverse-complement string
        if i < 32:
            fw0 |= ci << np.uint64(2 * i)
        else:
            fw1 |= ci << np.uint64(2 * (i - 32))
        if j < 32:
            rc0 |= cc << np.uint64(2 * j)
        else:
            rc1 |= cc << np.uint64(2 * (j - 32))
        sh = 2 * (K - 1 - i)               # MSB-first position of nt_i
        if sh >= 64:
            fhi |= ci << np.uint64(sh - 64)
        else:
            flo |= ci << np.uint64(sh)
        shr = 2 * (K - 1 - j)
        if shr >= 64:
            rhi |= cc << np.uint64(shr - 64)
        else:
            rlo |= cc << np.uint64(shr)
    fwd = (fhi < rhi) | ((fhi == rhi) & (flo <= rlo))
    out = np.empty((n, 2), np.uint64)
    out[:, 0] = np.where(fwd, fw0, rc0)
    out[:, 1] = np.where(fwd, fw1, rc1)
    return out[valid]


def unique_kmers(kmers: np.ndarray):
    """(unique (m,2) array, multiplicity) — sorted by (w1, w0)."""
    if len(kmers) == 0:
        return kmers, np.zeros(0, np.int64)
    order = np.lexsort((kmers[:, 0], kmers[:, 1]))
    s = kmers[order]
    new = np.ones(len(s), bool)
    new[1:] = (s[1:] != s[:-1]).any(axis=1)
    idx = np.nonzero(new)[0]
    mult = np.diff(np.append(idx, len(s)))
    return s[idx], mult


def _rev2_64(x: np.ndarray) -> np.ndarray:
    """Reverse the order of the 32 2-bit groups of each uint64."""
    x = ((x >> np.uint64(2)) & np.uint64(0x3333333333333333)) | ((x & np.uint64(0x3333333333333333)) << np.uint64(2))
    x = ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F0F0F0F0F)) | ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4))
    return x.byteswap()


def lexicographic_words(kmers: np.ndarray):
    """(hi, lo) of V = sum_i code(nt_i) << 2*(K-1-i) for packed k-mers in the boundary layout: integer order of V is
    the lexicographic order of the k-mer strings."""
    pad = np.uint64(128 - 2 * K)
    r0, r1 = _rev2_64(np.ascontiguousarray(kmers[:, 0])), _rev2_64(np.ascontiguousarray(kmers[:, 1]))
    return r0 >> pad, (r0 << (np.uint64(64) - pad)) | (r1 >> pad)


def kmc_order(kmers: np.ndarray) -> np.ndarray:
    """Permutation that puts packed k-mers in the record order of a KMC database: lexicographic, A<C<G<T from
    nucleotide 0 (external/kmc_api/kmc_file.cpp:428-515: prefix LUT over the leading nucleotides, sorted suffixes)."""
    hi, lo = lexicographic_words(kmers)
    return np.lexsort((lo, hi))


def nb_counts(rng, copies: np.ndarray, mean: float, var: float) -> np.ndarray:
    """NB(mean*copies, var*copies) draws, saturated at 255 (KmerCounts.cpp:178-189)."""
    p = mean / var
    size = mean * mean / (var - mean) * copies
    return np.minimum(rng.negative_binomial(size, p), 255).astype(np.uint8)


def sample_kmer_counts(haplotypes: list, seed: int, mean: float, var: float, n_errors: int = 0):
    """The k-mer spectrum a KMC run on reads of this sample would hold."""
    rng = np.random.default_rng(seed)
    km = np.concatenate([canonical_kmers(h) for h in haplotypes])
    uniq, copies = unique_kmers(km)
    counts = nb_counts(rng, copies, mean, var)
    keep = counts > 0                      # KMC -ci1: zero-count k-mers are absent
    uniq, counts = uniq[keep], counts[keep]
    if n_errors:
        err = rng.integers(0, 2**64, size=(n_errors, 2), dtype=np.uint64)
        err[:, 1] &= np.uint64((1 << (2 * K - 64)) - 1)
        uniq = np.concatenate([uniq, err])
        counts = np.concatenate([counts, np.ones(n_errors, np.uint8)])
    # NB: record order here is the generator's (fixtures pin it by hash); kmc_order() gives the order of a real KMC database
    return np.ascontiguousarray(uniq), np.ascontiguousarray(counts)


# ---- configs ---------------------------------------------------------------------------------
def config_a(n_variants: int = 10_000, length: int = 1_000_000, n_samples: int = 1, seed: int = 1) -> Workload:
    """BASELINE.json configs[0]: 1 sample, 10k SNVs on a 1 Mb reference."""
    ref = random_reference(length, seed)
    var = make_variants(ref, n_variants, seed + 1)
    g = make_genotypes(len(var), n_samples, seed + 2)
    return Workload("A", "chr1", ref, var, g, ["F"] * n_samples)


def config_b(n_variants: int = 300_000, length: int = 50_800_000, n_prefix: int = 10_000_000, seed: int = 11) -> Workload:
    """configs[1]: chr22-like, 85% SNV / 7.5% deletions / 7.5% insertions, 1 sample."""
    ref = random_reference(length, seed, n_prefix)
    var = make_variants(ref, n_variants, seed + 1, 0.075, 0.075, lo=n_prefix + K)
    g = make_genotypes(len(var), 1, seed + 2)
    return Workload("B", "chr22", ref, var, g, ["F"])


def small_mixed(n_variants: int, length: int, n_samples: int, seed: int, frac_indel: float = 0.15, chrom: str = "chr1") -> Workload:
    ref = random_reference(length, seed)
    var = make_variants(ref, n_variants, seed + 1, frac_indel / 2, frac_indel / 2, max_indel=20)
    rng = np.random.default_rng(seed + 5)
    af = rng.beta(0.5, 0.8, len(var))
    g = make_genotypes(len(var), n_samples, seed + 2, allele_freq=af)
    genders = ["F" if i % 2 == 0 else "M" for i in range(n_samples)]
    return Workload("mixed", chrom, ref, var, g, genders)


def nested_sv(n_sv: int, length: int, n_samples: int, seed: int, sv_len=(150, 500), inner_rate: float = 0.012,
              n_background: int = 0, chrom: str = "chr1", repeat_frac: float = 0.0) -> Workload:
    """Large deletions with SNVs inside their span: the reference turns the inner variants into clusters
    nested under the deletion's cluster (VariantFileParser.cpp:735-1000, has_dependency) and k-mers shared
    between the clusters of such a group into multicluster k-mers (KmerCounts.cpp:137-160)."""
    rng = np.random.default_rng(seed)
    r = np.frombuffer(random_reference(length, seed), np.uint8).copy()
    starts = np.sort(rng.choice(np.arange(2 * K, length - 2 * K - sv_len[1], 4 * sv_len[1]), size=n_sv, replace=False))
    spans = []
    for p in starts.tolist():
        p += int(rng.integers(0, sv_len[1]))
        ln = int(rng.integers(sv_len[0], sv_len[1]))
        rep = ln >= 300 and rng.random() < repeat_frac
        if rep:     # a 110-nt segment duplicated inside the deleted span: k-mers shared by clusters of ONE group
            r[p + 150:p + 260] = r[p + 20:p + 130]
        spans.append((p, ln, rep))
    ref = r.tobytes()
    var = []
    for p, ln, rep in spans:
        var.append(Variant(p, r[p:p + 1 + ln].tobytes(), [r[p:p + 1].tobytes()]))
        inner = np.flatnonzero(rng.random(ln - 2) < inner_rate) + p + 2
        if rep:
            inner = np.array([p + 50, p + 230], np.int64)
        for q in inner.tolist():
            alt = _ACGT[(int(_CODE[r[q]]) + int(rng.integers(1, 4))) % 4]
            var.append(Variant(q, r[q:q + 1].tobytes(), [bytes([alt])]))
    if n_background:
        taken = [(v.pos - K, v.pos + len(v.ref) + K) for v in var if len(v.ref) > 1]
        for v in make_variants(ref, n_background, seed + 3, 0.05, 0.05, max_indel=20):
            if not any(a <= v.pos <= b for a, b in taken):
                var.append(v)
    var.sort(key=lambda v: (v.pos, -len(v.ref)))
    af = rng.beta(0.8, 0.8, len(var))
    g = make_genotypes(len(var), n_samples, seed + 2, allele_freq=af)
    genders = ["F" if i % 2 == 0 else "M" for i in range(n_samples)]
    return Workload("nested", chrom, ref, var, g, genders)


def deep_nested(n_top: int, length: int, n_samples: int, seed: int, n_background: int = 0, chrom: str = "chr1") -> Workload:
    """Deletions inside deletions inside deletions (up to four levels) with SNVs at the bottom: groups of ten and more clusters whose
    dependency forest is several levels deep (VariantFileParser.cpp:735-1000, VariantClusterGroup.cpp:88-137), and variants that overlap
    three or more clusters at once.  The generator of tools/fuzz_graph_builder.py --deep, with genotypes and background variants."""
    rng = np.random.default_rng(seed)
    ref = random_reference(length, seed + 100)
    var = {}

    def put(p, r, alts):
        if p not in var and 60 < p and p + len(r) < length - 60:
            var[p] = Variant(p, r, alts)

    def snv(p):
        put(p, ref[p:p + 1], [bytes([_ACGT[(int(_CODE[ref[p]]) + 1) % 4]])])

    def fill(lo, hi, depth):
        if hi - lo < 250 or depth > 3:
            for p in rng.integers(lo + 60, max(lo + 61, hi - 60), size=int(rng.integers(0, 3))).tolist():
                snv(int(p))
            return
        cuts = np.sort(rng.integers(lo + 70, hi - 70, size=2 * int(rng.integers(1, 4))))
        for a, b in zip(cuts[0::2].tolist(), cuts[1::2].tolist()):
            if b - a > 120:
                put(a, ref[a:a + 1 + b - a], [ref[a:a + 1]])
                fill(a, b, depth + 1)
            else:
                snv(a)

    slot = (length - 400) // n_top
    for t in range(n_top):
        a = 200 + t * slot + int(rng.integers(0, max(1, slot // 8)))
        ln = int(rng.integers(1500, max(1501, min(3500, slot - slot // 8 - 300))))
        put(a, ref[a:a + 1 + ln], [ref[a:a + 1]])
        fill(a, a + ln, 1)
    out = [var[p] for p in sorted(var)]
    if n_background:
        taken = [(v.pos - K, v.pos + len(v.ref) + K) for v in out if len(v.ref) > 1]
        for v in make_variants(ref, n_background, seed + 3, 0.05, 0.05, max_indel=20):
            if v.pos not in var and not any(a <= v.pos <= b for a, b in taken):
                out.append(v)
    out.sort(key=lambda v: (v.pos, -len(v.ref)))
    af = rng.beta(0.8, 0.8, len(out))
    g = make_genotypes(len(out), n_samples, seed + 2, allele_freq=af)
    return Workload("deep", chrom, ref, out, g, ["F" if i % 2 == 0 else "M" for i in range(n_samples)])


def sample_spectra(w: Workload, seed: int = 4, n_errors: int = 0):
    out = []
    for s in range(w.genotypes.shape[0]):
        # males carry one copy of chrX (ChromosomePloidy.cpp:60-75 genotypes them haploid there)
        n_hap = 1 if (w.genders[s] == "M" and w.chrom.lower() in ("x", "chrx")) else 2
        haps = [apply_variants(w.reference, w.variants, w.genotypes[s, :, h]) for h in range(n_hap)]
        out.append(sample_kmer_counts(haps, seed + 17 * s, w.depth_mean, w.depth_var, n_errors))
    return out


def write_kmer_file(path, kmers: np.ndarray, counts: np.ndarray | None = None) -> None:
    with open(path, "wb") as f:
        np.array([len(kmers)], np.uint64).tofile(f)
        np.ascontiguousarray(kmers, np.uint64).tofile(f)
        if counts is not None:
            np.ascontiguousarray(counts, np.uint8).tofile(f)


def write_workdir(w: Workload, wd, spectra=None, seed: int = 4, n_errors: int = 0) -> Path:
    """genome.fa, variants.vcf, samples.tsv and <sample>.kmers.bin under wd."""
    wd = Path(wd)
    wd.mkdir(parents=True, exist_ok=True)
    with open(wd / "genome.fa", "wb") as f:
        f.write(b">" + w.chrom.encode() + b"\n")
        for i in range(0, len(w.reference), 60):
            f.write(w.reference[i:i + 60] + b"\n")
    with open(wd / "variants.vcf", "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for i, v in enumerate(w.variants):
            info = "ACO=" + ",".join(v.aco) if v.aco else "."
            f.write(f"{w.chrom}\t{v.pos + 1}\t{v.id or f'v{i}'}\t{v.ref.decode()}\t{','.join(a.decode() for a in v.alts)}\t.\t.\t{info}\n")
    spectra = spectra if spectra is not None else sample_spectra(w, seed, n_errors)
    with open(wd / "samples.tsv", "w") as f:
        for s, (km, ct) in enumerate(spectra):
            prefix = wd / f"S{s + 1}"
            f.write(f"S{s + 1}\t{w.genders[s]}\t{prefix}\n")
            write_kmer_file(str(prefix) + ".kmers.bin", km, ct)
    return wd
