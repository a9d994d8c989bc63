"""Host-side construction of variant clusters, variant-cluster groups and their graphs.

This is the part of `bayesTyper cluster` that SURVEY.md §8 leaves on the host: VariantFileParser's parsing and clustering
(src/bayesTyper/VariantFileParser.cpp:241-545 parseVariants, :581-733 allele bookkeeping and copy-number length, :735-978
clusterVariants, :1000-1040 mergeVariantClusters, :1107-1156 getVariantClusterGroupDependencies), the VariantClusterGraph
constructor (src/bayesTyper/VariantClusterGraph.cpp:62-377, contained clusters included) and the VariantClusterGroup constructor
(src/bayesTyper/VariantClusterGroup.cpp:47-105).  It is restated here so that the device stages can be driven end to end without
the reference; tests/test_graph_builder.py checks it array by array against what the reference built (oracle-R fixtures, nested
deletions included).  Orders that the reference takes from `std::unordered_map` / `unordered_set` iteration are reproduced with
`stdhash_order.UnorderedUInt`.

One behaviour of the reference is deliberately not reproduced: when an inference unit ends exactly in front of the first variant of a
new contig, the reference's next unit compares that variant's position with the last position of the PREVIOUS contig and exits with
"Variants need to be sorted by position" (VariantFileParser.cpp:283-293 runs before prev_position is updated, :268 sees equal contig
names on re-entry); build_genome_units simply starts the unit there.
"""
from __future__ import annotations

import numpy as np

from .stdhash_order import UnorderedUInt

K = 55
NONE16 = 0xFFFF
NONE32 = 0xFFFFFFFF
_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i
_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _right_trim(ref: bytes, alt: bytes):
    """VariantFileParser::rightTrimAllele (VariantFileParser.cpp:563-580)."""
    while len(ref) > 1 and len(alt) > 1 and ref[-1] == alt[-1]:
        ref, alt = ref[:-1], alt[:-1]
    return ref, alt


def _left_identical(ref: bytes, alt: bytes) -> int:
    n = 0
    for a, b in zip(ref, alt):
        if a != b:
            break
        n += 1
    return n


def _canonical(kmer: bytes) -> bytes:
    rc = kmer.translate(_COMP)[::-1]
    return kmer if kmer < rc else rc


def _approx_equal(a: float, b: float) -> bool:
    """Utils::doubleCompare (include/bayesTyper/Utils.hpp:81-87)."""
    return a == b or abs(a - b) < abs(min(a, b)) * 2.220446049250313e-16 * 100


def copy_number_variant_length(allele: bytes, chrom: bytes, start: int, k: int = K, threshold: float = 0.5) -> int:
    """How far downstream of an allele the genome keeps repeating the allele's k-mers
    (VariantFileParser::copyNumberVariantLength, VariantFileParser.cpp:641-733): extends the END of the group, never a cluster."""
    if len(allele) < k:
        return 0
    kmers = set()
    run = 0
    for i, c in enumerate(allele):
        run = run + 1 if c in b"ACGT" else 0
        if run >= k:
            kmers.add(_canonical(allele[i - k + 1:i + 1]))
    if not kmers:
        return 0
    n = len(chrom)
    length = 0
    window_end = min(start + length + len(allele), n)
    while True:
        run = bases = identical = 0
        best = (0.0, 0)
        for p in range(start + length, window_end):
            run = run + 1 if chrom[p] in b"ACGT" else 0
            if run >= k and _canonical(chrom[p - k + 1:p + 1]) in kmers:
                identical += 1
            bases += 1
            if identical:
                frac = identical / (bases - k + 1)
                if _approx_equal(frac, best[0]) or frac > best[0]:
                    best = (frac, bases)
        if best[0] < threshold:
            break
        length += best[1]
        if window_end == n:
            break
        window_end = min(start + length + len(allele), n)
    return length


class _Variant:
    """VariantCluster::Variant (include/bayesTyper/VariantCluster.hpp:56-71) plus the index into the caller's list."""
    __slots__ = ("input_idx", "has_dependency", "num_redundant", "alts", "aco")

    def __init__(self, input_idx, has_dependency):
        self.input_idx, self.has_dependency, self.num_redundant, self.alts, self.aco = input_idx, has_dependency, None, [], []


class _Cluster:
    """VariantCluster (include/bayesTyper/VariantCluster.hpp:98-108)."""
    __slots__ = ("idx", "left", "right", "variants", "contained")

    def __init__(self, idx, left, right):
        self.idx, self.left, self.right, self.variants, self.contained = idx, left, right, {}, []


def _cluster_variant(var, pos, ends, flanks, group, merge_sets, k):
    """VariantFileParser::clusterVariants (VariantFileParser.cpp:735-978).  flanks: position -> cluster, walked in position order."""
    for key in sorted(flanks):                              # flanks that no later variant can reach
        if pos - key < k:
            break
        del flanks[key]
    first, second = None, []

    def overlap(cl):
        nonlocal first
        if first is None:
            first = cl
            return True
        if first is not cl and cl not in second:
            second.append(cl)
        return False

    for key in sorted(flanks):
        cl = flanks[key]
        if abs(pos - key) + 1 <= k:
            if overlap(cl) and pos >= key:
                del flanks[key]
                continue
        for e in ends:
            if abs(e - key) + 1 <= k:
                if overlap(cl) and pos >= key:
                    raise AssertionError("unreachable in the reference: an end position is never closer to an earlier flank than the start")
            elif pos < key < e:
                overlap(cl)
    # the reference keeps `second` in a std::set of POINTERS and feeds the merge set in that order: with two or more extra clusters the
    # surviving index depends on malloc's placement of the clusters; creation order is the deterministic choice made here
    second.sort(key=lambda c: c.idx)
    if first is None:
        cl = _Cluster(len(group), pos, ends[-1])
        cl.variants[pos] = var
        for e in ends:
            flanks[e] = cl
        if ends[-1] - pos >= k:
            flanks.setdefault(pos, cl)
        group.insert(cl.idx, cl)
    else:
        if pos in first.variants:
            raise ValueError(f"several variants at position {pos + 1}: they need to be one multi-allelic variant")
        first.variants[pos] = var
        first.right = max(ends[-1], first.right)
        for e in ends:
            flanks.setdefault(e, first)
        if ends[-1] - pos >= k:
            flanks.setdefault(pos, first)
    if second:
        found = None
        i = 0
        while i < len(merge_sets):
            s = merge_sets[i]
            if first.idx in s:
                if found is None:
                    found = s
                elif s is not found:
                    for x in s:
                        found.insert(x)
                    merge_sets.pop(i)
                    continue
            merged = False
            for cl in second:
                if cl.idx in s:
                    if found is None:
                        found = s
                    elif s is not found:
                        for x in s:
                            found.insert(x)
                        merge_sets.pop(i)
                        merged = True
                        break
            if not merged:
                i += 1
        if found is None:
            found = UnorderedUInt()
            merge_sets.append(found)
        found.insert(first.idx)
        for cl in second:
            found.insert(cl.idx)


def _merge_clusters(group, merge_sets):
    """VariantFileParser::mergeVariantClusters (VariantFileParser.cpp:1000-1040): the first cluster of a set, in the set's
    own iteration order, absorbs the others and keeps its index."""
    for s in merge_sets:
        it = iter(s)
        keep = group[next(it)]
        for other_idx in it:
            other = group[other_idx]
            keep.left, keep.right = min(keep.left, other.left), max(keep.right, other.right)
            for p, v in other.variants.items():
                if p in keep.variants:
                    raise ValueError("two clusters to merge hold a variant at the same position")
                keep.variants[p] = v
            group.erase(other_idx)


def parse_variants(chrom: str, reference: bytes, variants, k: int = K, max_allele_length: int = 500000,
                   copy_number_variant_threshold: float = 0.5):
    """VariantFileParser::parseVariants for one contig (VariantFileParser.cpp:241-545).

    variants: objects with .pos (0-based), .ref (bytes), .alts (list of bytes), optionally .id and .aco (per alternative allele
    call-set origin), sorted by position.
    Returns (groups, regions, breaks): every group an UnorderedUInt cluster index -> _Cluster, in file order; regions the inclusive
    (start, end) intercluster stretches of at least k nucleotides; breaks[i] says whether an inference unit may end in front of
    variant i (it lies at least k past the end of the running group, :286)."""
    chrom_up = reference.upper()
    n = len(reference)
    copy_number_variant_threshold = float(np.float32(copy_number_variant_threshold))      # the option is a float in the reference
    groups, regions, breaks = [], [], []
    group, merge_sets, flanks = UnorderedUInt(), [], {}
    dependencies = set()
    prev_pos, prev_var_end, group_end = None, -1, -1

    def flush():
        nonlocal group, merge_sets, flanks
        if len(group):
            _merge_clusters(group, merge_sets)
            groups.append(group)
            group = UnorderedUInt()
        merge_sets, flanks = [], {}

    def add_region(a, b):
        if b - a + 1 >= k:
            regions.append((a, b))

    for vi, v in enumerate(variants):
        pos = v.pos
        if prev_pos is not None and pos < prev_pos:
            raise ValueError(f"variants need to be sorted by position: {prev_pos + 1} is before {pos + 1}")
        dependencies = {d for d in dependencies if d >= pos}
        breaks.append(pos - group_end >= k)
        prev_pos = pos
        ref = bytes(v.ref).upper()
        alts = [bytes(a).upper() for a in v.alts]
        var = _Variant(vi, bool(dependencies))
        if alts and alts[-1] == b"*":
            if not var.has_dependency:
                raise ValueError(f"'*' allele at position {pos + 1} without an overlapping upstream variant")
            alts.pop()
        if not alts:
            raise ValueError(f"variant at position {pos + 1} has no alternative allele")
        if len(set(alts)) != len(alts):
            raise ValueError(f"duplicate alternative alleles at position {pos + 1}")
        if pos + len(ref) > n:
            raise ValueError(f"variant at position {pos + 1} runs past the end of the contig")
        origins = getattr(v, "aco", None) or [""] * len(alts)               # INFO ACO, one per alternative allele ('*' dropped above)
        pairs = [_right_trim(ref, a) for a in alts]
        excluded = chrom_up[pos:pos + len(ref)] != ref or pos < k - 1
        included = []
        if not excluded:
            for i, (r, a) in enumerate(pairs):
                if pos + len(r) - 1 + k > n or len(r) > max_allele_length or len(a) > max_allele_length:
                    continue
                dependencies.add(pos + len(r) - 1)
                included.append(i)
        if excluded or not included:
            continue
        if pos - group_end >= k:
            flush()
        if pos > prev_var_end + 1:
            add_region(prev_var_end + 1, pos - 1)
        ends = set()
        for i in included:
            r, a = pairs[i]
            li = _left_identical(r, a)
            var.num_redundant = li if var.num_redundant is None else min(var.num_redundant, li)
            var.alts.append((len(r), a))
            var.aco.append(origins[i])
            after = pos + len(r)
            cnv = max(copy_number_variant_length(r, chrom_up, after, k, copy_number_variant_threshold),
                      copy_number_variant_length(a, chrom_up, after, k, copy_number_variant_threshold))
            ends.add(after - 1)
            group_end = max(group_end, after - 1 + cnv)
        prev_var_end = max(prev_var_end, max(ends))
        _cluster_variant(var, pos, sorted(ends), flanks, group, merge_sets, k)
    flush()
    if prev_var_end + 1 <= n - 1:
        add_region(prev_var_end + 1, n - 1)
    return groups, regions, breaks


def _group_dependencies(group):
    """VariantFileParser::getVariantClusterGroupDependencies (VariantFileParser.cpp:1107-1156): cluster -> the tightest cluster
    that contains it; registers the contained clusters on their containers (position order)."""
    deps = UnorderedUInt()
    items = group.items()
    for idx, cl in items:
        container = None
        for idx2, other in items:
            if idx2 == idx:
                continue
            if cl.left > other.left and cl.right < other.right:
                if container is None or (other.left > container.left and other.right < container.right):
                    container = other
            elif not (cl.left < other.left and cl.right > other.right):
                if not (cl.right < other.left or other.right < cl.left):
                    raise ValueError("clusters of a group overlap without one containing the other")
        if container is not None:
            deps.insert(idx, container.idx)
    for idx, container_idx in deps.items():
        cl = group[idx]
        group[container_idx].contained.append((cl.idx, cl.left, cl.right))
    for _, cl in items:
        cl.contained.sort(key=lambda t: t[1])
    return deps


class _Graph:
    """One cluster's graph, vertex for vertex as VariantClusterGraph's constructor lays it out."""

    def __init__(self):
        self.seq, self.flags, self.var, self.allele, self.nested, self.refvar, self.in_src = [], [], [], [], [], [], []

    def add_vertex(self):
        self.seq.append(bytearray()); self.flags.append(0); self.var.append(NONE16); self.allele.append(NONE16)
        self.nested.append(NONE32); self.refvar.append([]); self.in_src.append([])
        return len(self.seq) - 1

    def add_edge(self, u, v):
        self.in_src[v].append(u)

    def init_vertex(self, cur, codes, allele_idx, refvars, nested, redundant):
        """initVertex (VariantClusterGraph.cpp:316-377): a vertex that follows a contained cluster is disconnected from its
        predecessor; non-ACGT runs split the vertex.  Returns the last vertex."""
        self.var[cur], self.allele[cur] = allele_idx
        self.refvar[cur] = list(refvars)
        self.nested[cur] = nested
        self.flags[cur] = (1 if redundant else 0) | (2 if nested != NONE32 else 0)
        prev_disc = False
        for c in codes:
            if c > 3:
                if not prev_disc:
                    nxt = self.add_vertex()
                    self.add_edge(cur, nxt)
                    self.var[nxt], self.allele[nxt] = allele_idx
                    self.refvar[nxt] = list(refvars)
                    self.flags[nxt] = 2          # is_disconnected
                    cur = nxt
                prev_disc = True
            else:
                self.seq[cur].append(c)
                prev_disc = False
        return cur

    def add_vertices(self, cur, pieces, allele_idx, open_refvars, nested, redundant):
        """addVertices (VariantClusterGraph.cpp:284-314): one vertex per piece of sequence between contained clusters."""
        refvars = sorted(r for r in open_refvars if r != allele_idx[0])
        cur = self.init_vertex(cur, pieces[0], allele_idx, refvars, NONE32, redundant)
        for piece, inner in zip(pieces[1:], nested):
            nxt = self.add_vertex()
            self.add_edge(cur, nxt)
            cur = self.init_vertex(nxt, piece, allele_idx, refvars, inner, False)
        return cur


def build_cluster_graph(chrom_codes: np.ndarray, variants, contained=(), k: int = K):
    """variants: list of (pos, num_redundant, [(ref_len, alt_codes)...]) sorted by pos; contained: (cluster index, left flank,
    right flank) of the clusters nested in this one, by position.  VariantClusterGraph.cpp:62-282."""
    g = _Graph()
    added = {}                   # position -> ([vertices], [variant indices whose reference allele ends here])
    open_ref = set()
    contained = list(contained)
    next_contained = 0
    first = variants[0][0]
    cur = g.add_vertex()
    cur = g.add_vertices(cur, [chrom_codes[first - (k - 1):first]], (NONE16, NONE16), open_ref, [], False)
    prev_vertex = cur
    added[first] = ([cur], [])
    for vi, (pos, n_red, alts) in enumerate(variants):
        redundant = n_red > 0
        max_ref = 0
        for ai, (ref_len, alt_codes) in enumerate(alts):
            max_ref = max(max_ref, ref_len)
            nxt = g.add_vertex()
            g.add_edge(cur, nxt)
            nxt = g.add_vertices(nxt, [alt_codes], (vi, ai + 1), open_ref, [], redundant)
            added.setdefault(pos + ref_len, ([], []))[0].append(nxt)
        added[pos + max_ref][1].append(vi)
        open_ref.add(vi)
        last_variant = vi + 1 == len(variants)
        next_pos = None if last_variant else variants[vi + 1][0]
        more = True
        while more:
            cur_pos = min(added)
            next_vertices, to_erase = added.pop(cur_pos)
            for r in to_erase:
                open_ref.discard(r)
            if not added:
                more = False
                cur_last = cur_pos + k - 1 if last_variant else next_pos
            else:
                cur_last = min(added)
                if not last_variant and cur_last > next_pos:
                    more = False
                    cur_last = next_pos
            pieces, nested = [], []
            while next_contained < len(contained) and contained[next_contained][1] < cur_last:
                inner, left, right = contained[next_contained]
                if not (cur_pos <= left and right + k <= cur_last):
                    raise ValueError("contained cluster does not fit inside one stretch of its container's graph")
                pieces.append(chrom_codes[cur_pos:left])
                nested.append(inner)
                cur_pos = right + 1
                next_contained += 1
            pieces.append(chrom_codes[cur_pos:cur_last])
            cur = g.add_vertex()
            is_ref = False
            for v in next_vertices:
                if v == prev_vertex:
                    is_ref = True
                g.add_edge(v, cur)
            if is_ref:
                cur = g.add_vertices(cur, pieces, (vi, 0), open_ref, nested, redundant)
            else:
                cur = g.add_vertices(cur, pieces, (NONE16, NONE16), open_ref, nested, False)
            added.setdefault(cur_last, ([], []))[0].append(cur)
        prev_vertex = cur
    if next_contained != len(contained):
        raise ValueError("contained cluster lies outside its container's graph")
    return g


def _build_contig(chrom, reference, variants, k, max_allele_length, copy_number_variant_threshold):
    """The groups of one contig, unsorted: (contig, nucleotide codes, clusters, sources, out_edges, start, end, variants)."""
    codes = _CODE[np.frombuffer(reference, np.uint8)]
    groups, regions, breaks = parse_variants(chrom, reference, variants, k, max_allele_length, copy_number_variant_threshold)
    built = []
    for group in groups:
        deps = _group_dependencies(group)
        clusters = [cl for _, cl in group.items()]                       # the order VariantClusterGroup's vertices take
        slot = {cl.idx: i for i, cl in enumerate(clusters)}
        sources = [i for i, cl in enumerate(clusters) if cl.idx not in deps]
        out_edges = [[] for _ in clusters]
        for inner, container in deps.items():
            out_edges[slot[container]].append(slot[inner])
        start = min(cl.left for cl in clusters) + 1
        end = max(cl.right for cl in clusters) + 1
        built.append((chrom, codes, clusters, sources, out_edges, start, end, sum(len(cl.variants) for cl in clusters)))
    return built, regions, breaks


def build_unit_graphs(chrom: str, reference: bytes, variants, k: int = K, max_allele_length: int = 500000,
                      copy_number_variant_threshold: float = 0.5) -> dict:
    """Variant-cluster groups of one contig sorted like main.cpp:247, as the CSR arrays of graphs.btd / btg_graphs_desc.

    variants: objects with .pos (0-based), .ref (bytes), .alts (list of bytes), sorted by position.  On top of the arrays the
    reference's graphs hold, `var_input_idx` maps every variant of the unit back to the caller's list, and `group_start` /
    `group_end` carry the 1-based region of every group (VariantClusterGroup::region)."""
    built, regions, _ = _build_contig(chrom, reference, variants, k, max_allele_length, copy_number_variant_threshold)
    out = _emit(built, k)
    del out["_order"]
    out["regions"] = np.array(regions, np.int64).reshape(-1, 2)
    return out


def _build_genome(genome, candidates, decoys, k, max_allele_length, copy_number_variant_threshold):
    """Groups of all contigs (unsorted, file order), regions as (contig index, is_decoy, start, end) and, per candidate contig,
    the unit break flags of its variants."""
    names = list(genome)
    index = {n: i for i, n in enumerate(names)}
    decoys = set(decoys)
    built, regions, visited, breaks = [], [], set(), {}
    for chrom, variants in candidates.items():
        if chrom not in genome:
            raise ValueError(f'variants on contig "{chrom}", which the genome does not hold')
        if chrom in decoys:
            breaks[chrom] = [v.pos + 1 >= k for v in variants]          # the running group end is -1 on a contig without clusters
            continue
        visited.add(chrom)
        b, r, breaks[chrom] = _build_contig(chrom, genome[chrom], variants, k, max_allele_length, copy_number_variant_threshold)
        built.extend(b)
        regions.extend((index[chrom], 0, x, y) for x, y in r)
    for chrom in names:
        if chrom not in visited and len(genome[chrom]) >= k:
            regions.append((index[chrom], int(chrom in decoys), 0, len(genome[chrom]) - 1))
    return names, index, built, regions, breaks


def _finish_genome(out, names, index, built):
    group_contig = np.array([index[built[i][0]] for i in out.pop("_order")], np.uint32)
    out["contig_names"] = names
    out["group_contig"] = group_contig
    out["var_contig"] = np.repeat(group_contig, np.diff(out["cl_var_off"][out["group_cluster_off"].astype(np.int64)].astype(np.int64)))
    return out


def build_genome_graphs(genome: dict, candidates: dict, decoys=(), k: int = K, max_allele_length: int = 500000,
                        copy_number_variant_threshold: float = 0.5) -> dict:
    """The unit of a whole genome: genome = contig -> sequence (FASTA order, decoy contigs included), candidates = contig ->
    position-sorted variants (VCF order), decoys = names of the decoy contigs.

    Like the reference, variants on decoy contigs are dropped (VariantFileParser.cpp:332-341) and a variant on a contig that
    the genome does not hold is an error (the reference asserts in Chromosomes::isDecoy, Chromosomes.cpp:145); every contig
    contributes its intercluster regions — a contig without usable variants as one
    region, decoy contigs flagged (:273-280,512-536) — and the groups of all contigs are sorted together (main.cpp:247).
    On top of build_unit_graphs' arrays: `contig_names`, `group_contig` / `var_contig` (index into contig_names; `var_input_idx` is the
    index into that contig's candidate list) and `regions` as (contig index, is_decoy, start, end) rows."""
    names, index, built, regions, _ = _build_genome(genome, candidates, decoys, k, max_allele_length, copy_number_variant_threshold)
    out = _finish_genome(_emit(built, k), names, index, built)
    out["regions"] = np.array(regions, np.int64).reshape(-1, 4)
    return out


def build_genome_units(genome: dict, candidates: dict, decoys=(), min_unit_variants: int = 5_000_000, k: int = K, max_allele_length: int = 500000,
                       copy_number_variant_threshold: float = 0.5):
    """The inference units `bayesTyper cluster` splits a candidate set into (main.cpp:219,233-247; VariantFileParser.cpp:286-290):
    floor(variants / min_unit_variants) units (at least one) of ceil(variants / units) variant lines each — excluded lines count —
    every unit running on to the next variant that lies at least k past the end of its group.  Returns (list of per-unit arrays as
    build_genome_graphs gives them, regions of the whole genome)."""
    names, index, built, regions, breaks = _build_genome(genome, candidates, decoys, k, max_allele_length, copy_number_variant_threshold)
    total = sum(len(v) for v in candidates.values())
    n_units = max(1, int(np.floor(np.float32(total) / np.float32(min_unit_variants))))
    per_unit = int(np.ceil(np.float32(total) / np.float32(n_units)))
    unit_of, unit, count = {}, 0, 0
    for chrom, variants in candidates.items():
        for i in range(len(variants)):
            if count >= per_unit and breaks[chrom][i]:
                unit, count = unit + 1, 0
            count += 1
            unit_of[(chrom, i)] = unit
    per_unit_built = [[] for _ in range(unit + 1)]
    for b in built:
        first = min(v.input_idx for cl in b[2] for v in cl.variants.values())
        per_unit_built[unit_of[(b[0], first)]].append(b)
    if any(not bu for bu in per_unit_built):            # the reference asserts that every unit gains clusters (VariantFileParser.cpp:216-218)
        raise ValueError("an inference unit holds no usable variant (only excluded or decoy lines): raise min_unit_variants")
    units = [_finish_genome(_emit(bu, k), names, index, bu) for bu in per_unit_built]
    return units, np.array(regions, np.int64).reshape(-1, 4)


def _emit(built, k):
    """Groups in the unit's order (number of variants desc, then region string desc: VariantClusterGroupCompare,
    VariantClusterGroup.cpp:278-291) as CSR arrays."""
    order = sorted(range(len(built)), key=lambda i: (-built[i][7], _neg_str(f"{built[i][0]}:{built[i][5]}-{built[i][6]}")))
    out = {k_: [] for k_ in ("seq", "v_flags", "v_var", "v_allele", "v_nested", "v_refvar", "v_in_src", "var_pos", "var_dep", "var_nalt",
                             "alt_reflen", "alt_aco", "var_input_idx", "cluster_idx", "group_src", "group_edge_src", "group_edge_dst", "group_nvar",
                             "group_start", "group_end")}
    cl_vertex_off, v_seq_off, v_in_off, v_refvar_off, cl_var_off, var_alt_off, alt_seq_off = [0], [0], [0], [0], [0], [0], [0]
    group_cluster_off, group_src_off, group_edge_off = [0], [0], [0]
    alt_seq = bytearray()
    for gi in order:
        _chrom, codes, clusters, sources, out_edges, start, end, nvar = built[gi]
        out["group_nvar"].append(nvar); out["group_start"].append(start); out["group_end"].append(end)
        out["group_src"].extend(sources); group_src_off.append(len(out["group_src"]))
        for u, targets in enumerate(out_edges):
            for t in targets:
                out["group_edge_src"].append(u); out["group_edge_dst"].append(t)
        group_edge_off.append(len(out["group_edge_src"]))
        for cl in clusters:
            cvars = [cl.variants[p] for p in sorted(cl.variants)]
            g = build_cluster_graph(codes, [(p, cl.variants[p].num_redundant, [(rl, _CODE[np.frombuffer(a, np.uint8)]) for rl, a in cl.variants[p].alts])
                                            for p in sorted(cl.variants)], cl.contained, k)
            out["cluster_idx"].append(cl.idx)
            for v in range(len(g.seq)):
                out["seq"].append(np.frombuffer(bytes(g.seq[v]), np.uint8)); v_seq_off.append(v_seq_off[-1] + len(g.seq[v]))
                out["v_flags"].append(g.flags[v]); out["v_var"].append(g.var[v]); out["v_allele"].append(g.allele[v])
                out["v_nested"].append(g.nested[v])
                out["v_refvar"].extend(g.refvar[v]); v_refvar_off.append(len(out["v_refvar"]))
                out["v_in_src"].extend(g.in_src[v]); v_in_off.append(len(out["v_in_src"]))
            cl_vertex_off.append(len(out["v_flags"]))
            for p, var in zip(sorted(cl.variants), cvars):
                out["var_pos"].append(p + 1); out["var_dep"].append(int(var.has_dependency)); out["var_nalt"].append(len(var.alts))
                out["var_input_idx"].append(var.input_idx)
                out["alt_aco"].extend(var.aco)
                for rl, a in var.alts:
                    out["alt_reflen"].append(rl); alt_seq += a; alt_seq_off.append(len(alt_seq))
                var_alt_off.append(len(out["alt_reflen"]))
            cl_var_off.append(len(out["var_pos"]))
        group_cluster_off.append(len(out["cluster_idx"]))
    return {
        "group_cluster_off": np.array(group_cluster_off, np.uint64), "group_nvar": np.array(out["group_nvar"], np.uint32),
        "group_src_off": np.array(group_src_off, np.uint64), "group_src": np.array(out["group_src"], np.uint32),
        "group_edge_off": np.array(group_edge_off, np.uint64), "group_edge_src": np.array(out["group_edge_src"], np.uint32),
        "group_edge_dst": np.array(out["group_edge_dst"], np.uint32),
        "cluster_idx": np.array(out["cluster_idx"], np.uint32),
        "cl_vertex_off": np.array(cl_vertex_off, np.uint64), "cl_var_off": np.array(cl_var_off, np.uint64),
        "v_seq_off": np.array(v_seq_off, np.uint64), "seq": np.concatenate(out["seq"]) if out["seq"] else np.zeros(0, np.uint8),
        "v_flags": np.array(out["v_flags"], np.uint8), "v_var": np.array(out["v_var"], np.uint16), "v_allele": np.array(out["v_allele"], np.uint16),
        "v_nested": np.array(out["v_nested"], np.uint32),
        "v_refvar_off": np.array(v_refvar_off, np.uint64), "v_refvar": np.array(out["v_refvar"], np.uint16),
        "v_in_off": np.array(v_in_off, np.uint64), "v_in_src": np.array(out["v_in_src"], np.uint32),
        "var_pos": np.array(out["var_pos"], np.uint32), "var_dep": np.array(out["var_dep"], np.uint8), "var_nalt": np.array(out["var_nalt"], np.uint16),
        "var_alt_off": np.array(var_alt_off, np.uint64), "alt_reflen": np.array(out["alt_reflen"], np.uint32),
        "alt_seq_off": np.array(alt_seq_off, np.uint64), "alt_seq": np.frombuffer(bytes(alt_seq), np.uint8),
        "var_input_idx": np.array(out["var_input_idx"], np.int64), "alt_aco": list(out["alt_aco"]),
        "group_start": np.array(out["group_start"], np.uint32), "group_end": np.array(out["group_end"], np.uint32),
        "_order": order,
    }


class _neg_str:
    """Sort key that orders strings descending (std::string operator>)."""

    def __init__(self, s):
        self.s = s

    def __lt__(self, o):
        return self.s > o.s

    def __eq__(self, o):
        return self.s == o.s


def intercluster_regions(chrom: str, reference: bytes, variants, k: int = K):
    """Stretches of >= k nucleotides between consecutive variants' reference spans, head and tail included
    (VariantFileParser::addSequenceToInterclusterRegions, VariantFileParser.cpp:171-183,470-545): (start, end) inclusive."""
    return parse_variants(chrom, reference, variants, k)[1]


def build_genome_graphs_native(genome: dict, candidates: dict, decoys=(), k: int = K, max_allele_length: int = 500000,
                               copy_number_variant_threshold: float = 0.5) -> dict:
    """build_genome_graphs through the native builder (host/btcluster, include/btgpu_cluster.hpp): same arrays, ~15x faster on a
    chr22-sized candidate set.  The candidate set and the genome go through temporary files, as the tool reads the reference's inputs."""
    import subprocess
    import tempfile
    from pathlib import Path

    from . import btd, build
    if k != K:
        raise ValueError("the native builder is compiled for k = 55")
    build.build_host()
    exe = Path(build.ROOT) / "host" / "btcluster"
    if not exe.exists():
        raise RuntimeError("host/btcluster is not built (needs g++ and zlib); use build_genome_graphs")
    decoys = set(decoys)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for path, names in ((td / "genome.fa", [n for n in genome if n not in decoys]), (td / "decoy.fa", [n for n in genome if n in decoys])):
            with open(path, "wb") as f:
                for n in names:
                    f.write(b">" + n.encode() + b"\n" + bytes(genome[n]) + b"\n")
        with open(td / "candidates.vcf", "wb") as f:
            f.write(b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
            for n, variants in candidates.items():
                for i, v in enumerate(variants):
                    aco = getattr(v, "aco", None)
                    f.write(b"\t".join([n.encode(), str(v.pos + 1).encode(), (getattr(v, "id", None) or f"{n}_{i}").encode(), bytes(v.ref), b",".join(bytes(a) for a in v.alts),
                                        b".", b".", ("ACO=" + ",".join(aco)).encode() if aco else b"."]) + b"\n")
        cmd = [str(exe), str(td / "genome.fa"), str(td / "candidates.vcf"), str(td / "out.btd"), "--max-allele-length", str(max_allele_length),
               "--copy-number-variant-threshold", repr(float(copy_number_variant_threshold))]
        if decoys:
            cmd += ["--decoy", str(td / "decoy.fa")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise ValueError(r.stderr.strip() or "btcluster failed")
        out = btd.read(td / "out.btd")
    out["contig_names"] = bytes(out["contig_names"]).decode().split("\n")
    aco = bytes(out.pop("alt_aco")).decode()
    off = out.pop("alt_aco_off")
    out["alt_aco"] = [aco[int(a):int(b)] for a, b in zip(off[:-1], off[1:])]
    return out
