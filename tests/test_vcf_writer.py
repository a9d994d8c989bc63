"""GenotypeWriter parity: host/btvcf (include/btgpu_vcf.hpp) must reproduce, byte for byte, VCFs written by the
REFERENCE's own GenotypeWriter (tests/golden/vcf_*.vcf.gz, produced by oracle-R: make_vcf_fixtures.py) when it is given the
reference's numbers as flat btg_genotype_result arrays plus the variant / contig description.  CPU only: the writer is host code."""
import gzip
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bayestyper_b200 import btd
from tests.golden.make_vcf_fixtures import VCF_WORKLOADS

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def _build_btvcf():
    exe = ROOT / "host" / "btvcf"
    src = [ROOT / "host" / "btvcf.cpp", ROOT / "host" / "btd.hpp", ROOT / "include" / "btgpu_vcf.hpp", ROOT / "include" / "btgpu.hpp"]
    if not exe.exists() or any(s.stat().st_mtime > exe.stat().st_mtime for s in src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), "-I", str(ROOT / "host"), str(src[0]), "-o", str(exe)])
    return exe


def _strs(items):
    b = [x.encode() if isinstance(x, str) else bytes(x) for x in items]
    off = np.concatenate([[0], np.cumsum([len(x) for x in b])]).astype(np.uint64)
    return np.frombuffer(b"".join(b), np.uint8).copy() if b else np.zeros(0, np.uint8), off


def _arrays_from_vcf(text: str, reference: bytes, trim: bool):
    """The reference's numbers and variant description, as the C ABI's flat arrays.  trim=True describes every alternative allele
    by its right-trimmed sequence and reference length (as VariantInfo holds them), so the writer has to rebuild the suffix."""
    lines = text.splitlines()
    head = [l for l in lines if l.startswith("#")]
    body = [l.split("\t") for l in lines if not l.startswith("#")]
    samples = head[-1].split("\t")[9:]
    S = len(samples)
    opt = [l + "\n" for l in head if l.startswith("##BayesTyperOptions")]
    genome = [l for l in head if l.startswith("##reference=file:")][0][len("##reference=file:"):]
    cnames = [l.split("ID=")[1].split(",")[0] for l in head if l.startswith("##contig")]
    if isinstance(reference, dict):             # several contigs: name -> sequence, the ones missing from the header are decoys
        cnames += [n for n in reference if n not in cnames]
        seqs = [reference[n] for n in cnames]
    else:
        seqs = [reference]
    n_header = sum(l.startswith("##contig") for l in head)
    a = {"meta.n_samples": np.array([S], np.uint32)}
    a["vcf.sample_names"], a["vcf.sample_names_off"] = _strs(samples)
    a["vcf.contig_names"], a["vcf.contig_names_off"] = _strs(cnames)
    a["vcf.contig_seq"], a["vcf.contig_seq_off"] = _strs(seqs)
    a["vcf.contig_decoy"] = np.array([0] * n_header + [1] * (len(cnames) - n_header), np.uint8)
    a["vcf.genome_filename"] = np.frombuffer(genome.encode(), np.uint8).copy()
    a["vcf.graph_options_header"] = np.frombuffer(opt[0].encode(), np.uint8).copy()
    a["vcf.genotype_options_header"] = np.frombuffer(opt[1].encode(), np.uint8).copy()
    ids, vcr, vcgr, alt_seq, alt_aco, alt_len, alt_off = [], [], [], [], [], [], [0]
    pos, dep, vcs, vcgs, hc, an = [], [], [], [], [], []
    gt, gq, gpp, app, nak, fak, mac, saf, ploidy, ac, af, acp, anc = ([] for _ in range(13))
    for t in body:
        info = dict(kv.split("=", 1) for kv in t[7].split(";"))
        alts = t[4].split(",")
        has_dep = alts[-1] == "*"
        real = alts[:-1] if has_dep else alts
        acos = info["ACO"].split(",")
        for i, alt in enumerate(real):
            k = 0
            if trim:
                while k < min(len(alt), len(t[3])) - 1 and alt[-1 - k] == t[3][-1 - k]:
                    k += 1
            alt_seq.append(alt[:len(alt) - k]); alt_len.append(len(t[3]) - k); alt_aco.append("" if acos[i] == "." else acos[i])
        alt_off.append(len(alt_seq))
        nA = 1 + len(alts)
        nG = nA * (nA + 1) // 2
        ids.append(t[2]); pos.append(int(t[1])); dep.append(int(has_dep)); vcs.append(int(info["VCS"])); vcgs.append(int(info["VCGS"]))
        vcr.append(info["VCR"]); vcgr.append(info["VCGR"]); hc.append(int(info["HC"])); an.append(int(info["AN"]))
        ac += [0] + [int(x) for x in info["AC"].split(",")]
        af += [0.0] + [float(x) for x in info["AF"].split(",")]
        acp += [float(x) for x in info["ACP"].split(",")]
        nc = set(int(x) for x in info["ANC"].split(",")) if "ANC" in info else set()
        anc += [int(i in nc) for i in range(nA)]
        for s in t[9:]:
            f = s.split(":")
            if f[0] == "":                      # ":.:.:.:.:.:." — a sample without a genotype on this chromosome
                ploidy.append(0); gt += [0xFFFF, 0xFFFE]; gq.append(0)
                gpp += [0.0] * nG; app += [0.0] * nA; nak += [0.0] * nA; fak += [0.0] * nA; mac += [0.0] * nA; saf += [0] * nA
                continue
            g = f[0].split("/")
            ploidy.append(len(g))
            gt += [0xFFFF if g[0] == "." else int(g[0]), (0xFFFF if g[1] == "." else int(g[1])) if len(g) > 1 else 0xFFFE]
            gq.append(int(f[1]))
            p = [float(x) for x in f[2].split(",")]
            gpp += p + [0.0] * (nG - len(p))
            app += [float(x) for x in f[3].split(",")]; nak += [float(x) for x in f[4].split(",")]; fak += [float(x) for x in f[5].split(",")]
            mac += [float(x) for x in f[6].split(",")]; saf += [int(x) for x in f[7].split(",")]
    a["vcf.ids"], a["vcf.ids_off"] = _strs(ids)
    a["vcf.vcr"], a["vcf.vcr_off"] = _strs(vcr)
    a["vcf.vcgr"], a["vcf.vcgr_off"] = _strs(vcgr)
    a["vcf.alt_seq"], a["vcf.alt_seq_off"] = _strs(alt_seq)
    a["vcf.alt_aco"], a["vcf.alt_aco_off"] = _strs(alt_aco)
    a["vcf.alt_ref_length"] = np.array(alt_len, np.uint32); a["vcf.alt_off"] = np.array(alt_off, np.uint64)
    a["vcf.contig"] = np.array([cnames.index(t[0]) for t in body], np.uint32); a["vcf.position"] = np.array(pos, np.uint32); a["vcf.has_dependency"] = np.array(dep, np.uint8)
    a["vcf.vcs"] = np.array(vcs, np.uint32); a["vcf.vcgs"] = np.array(vcgs, np.uint32)
    for k, v, dt in (("gt", gt, np.uint16), ("gq", gq, np.uint32), ("gpp", gpp, np.float32), ("app", app, np.float32), ("nak", nak, np.float32), ("fak", fak, np.float32),
                     ("mac", mac, np.float32), ("saf", saf, np.uint16), ("ploidy", ploidy, np.uint8), ("an", an, np.uint32), ("ac", ac, np.uint32), ("af", af, np.float32),
                     ("acp", acp, np.float32), ("anc", anc, np.uint8), ("hc", hc, np.uint16)):
        a[k] = np.array(v, dt)
    return a


def _qual_tolerant_equal(got: str, want: str) -> bool:
    """QUAL is computed from the exact float ACP in the reference and from its 6-digit print here: allow the last printed digit."""
    g, w = got.split("\t"), want.split("\t")
    if g[:5] != w[:5] or g[6:] != w[6:]:
        return False
    return g[5] == w[5] or abs(float(g[5]) - float(w[5])) <= 2e-5 * max(1.0, abs(float(w[5])))


@pytest.mark.parametrize("name", list(VCF_WORKLOADS))
@pytest.mark.parametrize("trim", [False, True])
def test_writer_reproduces_the_reference_vcf(tmp_path, name, trim):
    exe = _build_btvcf()
    want = gzip.open(GOLD / f"{name}.vcf.gz", "rt").read()
    w = VCF_WORKLOADS[name]()
    btd.write(tmp_path / "in.btd", _arrays_from_vcf(want, w.reference, trim))
    r = subprocess.run([str(exe), str(tmp_path / "in.btd"), str(tmp_path / "out.vcf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = (tmp_path / "out.vcf").read_text()
    gl, wl = got.splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    n_qual = 0
    for a, b in zip(gl, wl):
        if a != b:
            assert not b.startswith("#") and _qual_tolerant_equal(a, b), f"\n got: {a[:300]}\nwant: {b[:300]}"
            n_qual += 1
    assert n_qual <= len(wl) // 20, "too many QUAL last-digit differences"


def _genome_fixture():
    from tests.golden.make_vcf_genome_fixture import genome_workload
    parts, empty, decoys = genome_workload()
    genome = {n: w.reference for n, w in parts.items()}
    genome[empty[0]] = empty[1]
    return parts, genome, decoys


@pytest.mark.parametrize("shuffle", [False, True])
def test_writer_reproduces_the_reference_vcf_of_a_genome(tmp_path, shuffle):
    """Several contigs (FASTA order chr2, chr1, chrX, a contig without variants) and a decoy: `##contig` lines, record order and
    per-contig ploidy as the reference wrote them (tests/golden/make_vcf_genome_fixture.py).  shuffle=True hands the variants over in
    a scrambled order: the writer, like GenotypeWriter::finalise, sorts them itself."""
    exe = _build_btvcf()
    want = gzip.open(GOLD / "vcf_genome_2s.vcf.gz", "rt").read()
    _, genome, decoys = _genome_fixture()
    if shuffle:
        lines = want.splitlines()
        body = [l for l in lines if not l.startswith("#")]
        np.random.default_rng(3).shuffle(body)
        text = "\n".join([l for l in lines if l.startswith("#")] + body) + "\n"
    else:
        text = want
    btd.write(tmp_path / "in.btd", _arrays_from_vcf(text, {**genome, **decoys}, True))
    r = subprocess.run([str(exe), str(tmp_path / "in.btd"), str(tmp_path / "out.vcf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    gl, wl = (tmp_path / "out.vcf").read_text().splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    for a, b in zip(gl, wl):
        assert a == b or (not b.startswith("#") and _qual_tolerant_equal(a, b)), f"\n got: {a[:300]}\nwant: {b[:300]}"


def test_description_of_a_genome_from_graph_builder_reproduces_the_reference_vcf(tmp_path):
    """build_genome_graphs -> vcf_desc.describe_genome (contig per variant, ids, VCR / VCGR with the contig's name, decoy flag) + the
    reference's numbers moved into unit order -> btvcf == the reference's file."""
    from bayestyper_b200 import graph_builder, vcf_desc
    _build_btvcf()
    want = gzip.open(GOLD / "vcf_genome_2s.vcf.gz", "rt").read()
    parts, genome, decoys = _genome_fixture()
    full = {**genome, **decoys}
    cand = {n: w.variants for n, w in parts.items()}
    graphs = graph_builder.build_genome_graphs(full, cand, decoys=list(decoys))
    head = [l for l in want.splitlines() if l.startswith("#")]
    opt = [l + "\n" for l in head if l.startswith("##BayesTyperOptions")]
    genome_file = [l for l in head if l.startswith("##reference=file:")][0][len("##reference=file:"):]
    S = len(head[-1].split("\t")) - 9
    desc = vcf_desc.describe_genome(full, cand, graphs, head[-1].split("\t")[9:], list(decoys), genome_file, opt[0], opt[1])
    by_file = _arrays_from_vcf(want, full, False)
    file_ids = [bytes(by_file["vcf.ids"][int(a):int(b)]).decode() for a, b in zip(by_file["vcf.ids_off"][:-1], by_file["vcf.ids_off"][1:])]
    unit_ids = [bytes(desc["vcf.ids"][int(a):int(b)]).decode() for a, b in zip(desc["vcf.ids_off"][:-1], desc["vcf.ids_off"][1:])]
    assert sorted(file_ids) == sorted(unit_ids)
    where = {v: i for i, v in enumerate(file_ids)}
    perm = np.array([where[v] for v in unit_ids])
    nA = (1 + np.diff(by_file["vcf.alt_off"].astype(np.int64)) + by_file["vcf.has_dependency"]).astype(np.int64)
    def take(arr, width):
        off = np.concatenate([[0], np.cumsum(width)])
        return np.concatenate([arr[off[v]:off[v + 1]] for v in perm])
    out = {}
    for k, wdt in (("gt", np.full(len(nA), 2 * S)), ("gq", np.full(len(nA), S)), ("ploidy", np.full(len(nA), S)), ("an", np.ones(len(nA), int)), ("hc", np.ones(len(nA), int)),
                   ("gpp", S * nA * (nA + 1) // 2), ("app", S * nA), ("nak", S * nA), ("fak", S * nA), ("mac", S * nA), ("saf", S * nA), ("ac", nA), ("af", nA), ("acp", nA), ("anc", nA)):
        out[k] = take(by_file[k], np.asarray(wdt, np.int64))
    vcf_desc.write_vcf(tmp_path / "out.vcf", out, desc, S)
    gl, wl = (tmp_path / "out.vcf").read_text().splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    for a, b in zip(gl, wl):
        assert a == b or (not b.startswith("#") and _qual_tolerant_equal(a, b)), f"\n got: {a[:300]}\nwant: {b[:300]}"


def test_soft_masked_genome_is_written_in_upper_case(tmp_path):
    """The reference upper-cases the genome on loading (Chromosomes::convertToUpper): a soft-masked contig gives the same file."""
    exe = _build_btvcf()
    want = gzip.open(GOLD / "vcf_mixed_3s.vcf.gz", "rt").read()
    w = VCF_WORKLOADS["vcf_mixed_3s"]()
    masked = bytearray(w.reference)
    masked[1000:9000] = bytes(masked[1000:9000]).lower()
    btd.write(tmp_path / "in.btd", _arrays_from_vcf(want, bytes(masked), True))
    r = subprocess.run([str(exe), str(tmp_path / "in.btd"), str(tmp_path / "out.vcf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for a, b in zip((tmp_path / "out.vcf").read_text().splitlines(), want.splitlines()):
        assert a == b or (not b.startswith("#") and _qual_tolerant_equal(a, b)), f"\n got: {a[:200]}\nwant: {b[:200]}"


def test_sample_without_genotype_and_filters(tmp_path):
    """Ploidy 0 (e.g. a female on chrY) prints the reference's ':.:.:.:.:.:.' (GenotypeWriter.cpp:58,319); AN = 0 gives FILTER AN0;
    a dependent variant gets the '*' allele and ACO '.' (GenotypeWriter.cpp:170-173,250-253)."""
    exe = _build_btvcf()
    ref = b"ACGTACGTACGTACGTACGT"
    hdr = ["##fileformat=VCFv4.2", "##reference=file:/g.fa", "##contig=<ID=chrY,length=20>", "##BayesTyperOptions=a", "##BayesTyperOptions=b",
           "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2"]
    rec = "chrY\t5\tv0\tA\tT,*\t0\tAN0\tAC=0,0;AF=0,0;AN=0;ACP=0,0,0;VCS=1;VCR=chrY:5-5;VCGS=1;VCGR=chrY:5-5;HC=2;ANC=1,2;ACO=.,.\tGT:GQ:GPP:APP:NAK:FAK:MAC:SAF\t:.:.:.:.:.:.\t.:0:0.5,0.5,0:0.5,0.5,0:-1,-1,-1:-1,-1,-1:-1,-1,-1:1,1,1"
    text = "\n".join(hdr + [rec]) + "\n"
    btd.write(tmp_path / "in.btd", _arrays_from_vcf(text, ref, False))
    r = subprocess.run([str(exe), str(tmp_path / "in.btd"), str(tmp_path / "out.vcf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = (tmp_path / "out.vcf").read_text().splitlines()
    assert got[-1] == rec


@pytest.mark.parametrize("name", list(VCF_WORKLOADS))
def test_description_from_graph_builder_reproduces_the_reference_vcf(tmp_path, name):
    """End to end on the host side: clusters built by graph_builder -> vcf_desc.describe (ids, positions, trimmed alleles, VCS / VCR /
    VCGS / VCGR in unit order) + the reference's numbers reordered into unit order -> btvcf == the reference's file."""
    from bayestyper_b200 import graph_builder, vcf_desc
    exe = _build_btvcf()
    want = gzip.open(GOLD / f"{name}.vcf.gz", "rt").read()
    w = VCF_WORKLOADS[name]()
    graphs = graph_builder.build_unit_graphs(w.chrom, w.reference, w.variants)
    head = [l for l in want.splitlines() if l.startswith("#")]
    opt = [l + "\n" for l in head if l.startswith("##BayesTyperOptions")]
    genome = [l for l in head if l.startswith("##reference=file:")][0][len("##reference=file:"):]
    desc = vcf_desc.describe(w.chrom, w.reference, w.variants, graphs, head[-1].split("\t")[9:], genome, opt[0], opt[1])
    # the reference's numbers, parsed in file order, moved to unit order (match by id)
    by_file = _arrays_from_vcf(want, w.reference, False)
    file_ids = [bytes(by_file["vcf.ids"][int(a):int(b)]).decode() for a, b in zip(by_file["vcf.ids_off"][:-1], by_file["vcf.ids_off"][1:])]
    unit_ids = [bytes(desc["vcf.ids"][int(a):int(b)]).decode() for a, b in zip(desc["vcf.ids_off"][:-1], desc["vcf.ids_off"][1:])]
    assert sorted(file_ids) == sorted(unit_ids)
    where = {v: i for i, v in enumerate(file_ids)}
    perm = np.array([where[v] for v in unit_ids])
    S = len(head[-1].split("\t")) - 9
    nA = (1 + np.diff(by_file["vcf.alt_off"].astype(np.int64)) + by_file["vcf.has_dependency"]).astype(np.int64)
    def take(arr, width):          # per-variant blocks of `width[v]` entries
        off = np.concatenate([[0], np.cumsum(width)])
        return np.concatenate([arr[off[v]:off[v + 1]] for v in perm])
    out = {}
    for k, wdt in (("gt", np.full(len(nA), 2 * S)), ("gq", np.full(len(nA), S)), ("ploidy", np.full(len(nA), S)), ("an", np.ones(len(nA), int)), ("hc", np.ones(len(nA), int)),
                   ("gpp", S * nA * (nA + 1) // 2), ("app", S * nA), ("nak", S * nA), ("fak", S * nA), ("mac", S * nA), ("saf", S * nA), ("ac", nA), ("af", nA), ("acp", nA), ("anc", nA)):
        out[k] = take(by_file[k], np.asarray(wdt, np.int64))
    vcf_desc.write_vcf(tmp_path / "out.vcf", out, desc, S)          # the call a user of the Python mirror makes
    gl, wl = (tmp_path / "out.vcf").read_text().splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    for a, b in zip(gl, wl):
        assert a == b or (not b.startswith("#") and _qual_tolerant_equal(a, b)), f"\n got: {a[:300]}\nwant: {b[:300]}"


def test_parameter_files_reproduce_the_reference(tmp_path):
    """<out>_noise_parameters.txt and <out>_genomic_parameters.txt (include/btgpu_params.hpp) byte for byte against the files the
    reference wrote for the chrX fixture (oracle-R run of make_vcf_fixtures.py; InferenceEngine.cpp:165-172,205,229,266,
    CountDistribution.cpp:70-78,128-137), given the reference's numbers."""
    exe = _build_btvcf()
    noise = gzip.open(GOLD / "params_chrx_2s_noise.txt.gz", "rt").read()
    genomic = (GOLD / "params_chrx_2s_genomic.txt").read_text()
    rows = [l.split("\t") for l in noise.splitlines()[1:]]
    trace = np.array([[float(x) for x in r] for r in rows], np.float64)
    gm = [l.split("\t") for l in genomic.splitlines()[1:]]
    mean, var = np.array([float(x[1]) for x in gm]), np.array([float(x[2]) for x in gm])
    nb_p = mean / var
    nb_size = mean * nb_p / (1 - nb_p)                       # mean = size (1 - p) / p
    want_vcf = gzip.open(GOLD / "vcf_chrx_2s.vcf.gz", "rt").read()
    w = VCF_WORKLOADS["vcf_chrx_2s"]()
    a = _arrays_from_vcf(want_vcf, w.reference, False)
    a["noise_trace"] = trace
    a["tab.nb_p_size"] = np.stack([nb_p, nb_size], 1)
    btd.write(tmp_path / "in.btd", a)
    r = subprocess.run([str(exe), str(tmp_path / "in.btd"), str(tmp_path / "out.vcf")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "out_noise_parameters.txt").read_text() == noise
    assert (tmp_path / "out_genomic_parameters.txt").read_text() == genomic
