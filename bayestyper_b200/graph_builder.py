"""Host-side construction of variant clusters and their graphs for NON-NESTED candidate sets
(SNVs, insertions, deletions that do not contain other variants).

This is the part of `bayesTyper cluster` that SURVEY.md §8 leaves on the host (VariantFileParser's
clustering, src/bayesTyper/VariantFileParser.cpp:185-545,735-978, and the VariantClusterGraph
constructor, src/bayesTyper/VariantClusterGraph.cpp:62-377).  It is restated here only so that the
device stages can be driven end to end without the reference; tests/test_graph_builder.py checks it
vertex by vertex against graphs the reference built (oracle-R fixtures).  Variants whose reference
span contains another variant (has_dependency / nested clusters) are rejected explicitly.
"""
from __future__ import annotations

import numpy as np

K = 55
NONE16 = 0xFFFF
NONE32 = 0xFFFFFFFF
_CODE = np.full(256, 4, np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i


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


class _Graph:
    """One cluster's graph, built like VariantClusterGraph's constructor (no contained clusters)."""

    def __init__(self):
        self.seq, self.flags, self.var, self.allele, self.refvar, self.in_src = [], [], [], [], [], []

    def add_vertex(self):
        self.seq.append(bytearray()); self.flags.append(0); self.var.append(NONE16); self.allele.append(NONE16)
        self.refvar.append([]); self.in_src.append([])
        return len(self.seq) - 1

    def add_edge(self, u, v):
        self.in_src[v].append(u)

    def init_vertex(self, cur, codes, allele_idx, refvars, redundant):
        """initVertex (VariantClusterGraph.cpp:316-377): non-ACGT runs split the vertex; returns the last vertex."""
        self.var[cur], self.allele[cur] = allele_idx
        self.refvar[cur] = list(refvars)
        self.flags[cur] = 1 if redundant else 0
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

    def add_vertices(self, cur, codes, allele_idx, open_refvars, redundant):
        """addVertices (VariantClusterGraph.cpp:284-314) without nested clusters."""
        refvars = sorted(r for r in open_refvars if r != allele_idx[0])
        return self.init_vertex(cur, codes, allele_idx, refvars, redundant)


def build_cluster_graph(chrom_codes: np.ndarray, variants):
    """variants: list of (pos, num_redundant, [(ref_len, alt_codes)...]) sorted by pos.  VariantClusterGraph.cpp:62-282."""
    g = _Graph()
    added = {}                   # position -> ([vertices], [variant indices whose reference allele ends here])
    open_ref = set()
    first = variants[0][0]
    cur = g.add_vertex()
    cur = g.add_vertices(cur, chrom_codes[first - (K - 1):first], (NONE16, NONE16), open_ref, False)
    prev_vertex = cur
    added[first] = ([cur], [])
    for vi, (pos, n_red, alts) in enumerate(variants):
        redundant = n_red > 0
        max_ref = 0
        for ai, (ref_len, alt_codes) in enumerate(alts):
            max_ref = max(max_ref, ref_len)
            nxt = g.add_vertex()
            g.add_edge(cur, nxt)
            nxt = g.add_vertices(nxt, alt_codes, (vi, ai + 1), open_ref, redundant)
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
                cur_last = cur_pos + K - 1 if last_variant else next_pos
            else:
                cur_last = min(added)
                if not last_variant and cur_last > next_pos:
                    more = False
                    cur_last = next_pos
            cur = g.add_vertex()
            is_ref = False
            for v in next_vertices:
                if v == prev_vertex:
                    is_ref = True
                g.add_edge(v, cur)
            if is_ref:
                cur = g.add_vertices(cur, chrom_codes[cur_pos:cur_last], (vi, 0), open_ref, redundant)
            else:
                cur = g.add_vertices(cur, chrom_codes[cur_pos:cur_last], (NONE16, NONE16), open_ref, False)
            added.setdefault(cur_last, ([], []))[0].append(cur)
        prev_vertex = cur
    return g


def build_unit_graphs(chrom: str, reference: bytes, variants, k: int = K) -> dict:
    """Clusters (= groups: no nesting), sorted like main.cpp:247, as the CSR arrays of graphs.btd / btg_graphs_desc.

    variants: objects with .pos (0-based), .ref (bytes), .alts (list of bytes); sorted, non-overlapping."""
    codes = _CODE[np.frombuffer(reference, np.uint8)]
    prepared = []
    prev_end = -1
    for v in variants:
        if v.pos <= prev_end:
            raise ValueError("overlapping variants (has_dependency / nested clusters) are not supported by this builder")
        alts, n_red, end = [], None, v.pos
        for alt in v.alts:
            r, a = _right_trim(v.ref.upper(), alt.upper())
            li = _left_identical(r, a)
            n_red = li if n_red is None else min(n_red, li)
            alts.append((len(r), _CODE[np.frombuffer(a, np.uint8)], a))
            end = max(end, v.pos + len(r) - 1)
        prepared.append((v.pos, n_red, alts, end))
        prev_end = max(prev_end, end)
    # clusters: a variant joins while it lies within k-1 of the running end (clusterVariants, VariantFileParser.cpp:735-770)
    clusters, cur, run_end = [], [], None
    for p in prepared:
        if cur and p[0] - run_end >= k:
            clusters.append(cur); cur = []
        cur.append(p)
        run_end = p[3] if len(cur) == 1 else max(run_end, p[3])
    if cur:
        clusters.append(cur)
    # group order: number of variants desc, then region string desc (VariantClusterGroupCompare, VariantClusterGroup.cpp:278-291)
    def region(cl):
        return f"{chrom}:{cl[0][0] + 1}-{max(x[3] for x in cl) + 1}"
    order = sorted(range(len(clusters)), key=lambda i: (-len(clusters[i]), _neg_str(region(clusters[i]))))
    out = {k_: [] for k_ in ("seq", "v_flags", "v_var", "v_allele", "v_refvar", "v_in_src", "var_pos", "var_dep", "var_nalt", "alt_reflen")}
    cl_vertex_off, v_seq_off, v_in_off, v_refvar_off, cl_var_off, var_alt_off, alt_seq_off = [0], [0], [0], [0], [0], [0], [0]
    alt_seq = bytearray()
    for ci in order:
        cl = clusters[ci]
        g = build_cluster_graph(codes, [(p, nr, [(rl, ac) for rl, ac, _ in alts]) for p, nr, alts, _ in cl])
        for v in range(len(g.seq)):
            out["seq"].append(np.frombuffer(bytes(g.seq[v]), np.uint8)); v_seq_off.append(v_seq_off[-1] + len(g.seq[v]))
            out["v_flags"].append(g.flags[v]); out["v_var"].append(g.var[v]); out["v_allele"].append(g.allele[v])
            out["v_refvar"].extend(g.refvar[v]); v_refvar_off.append(len(out["v_refvar"]))
            out["v_in_src"].extend(g.in_src[v]); v_in_off.append(len(out["v_in_src"]))
        cl_vertex_off.append(len(out["v_flags"]))
        for p, nr, alts, _ in cl:
            out["var_pos"].append(p + 1); out["var_dep"].append(0); out["var_nalt"].append(len(alts))
            for rl, _, a in alts:
                out["alt_reflen"].append(rl); alt_seq += a; alt_seq_off.append(len(alt_seq))
            var_alt_off.append(len(out["alt_reflen"]))
        cl_var_off.append(len(out["var_pos"]))
    C = len(order)
    return {
        "group_cluster_off": np.arange(C + 1, dtype=np.uint64), "group_nvar": np.diff(cl_var_off).astype(np.uint32),
        "group_src_off": np.arange(C + 1, dtype=np.uint64), "group_src": np.zeros(C, np.uint32),
        "group_edge_off": np.zeros(C + 1, np.uint64), "group_edge_src": np.zeros(0, np.uint32), "group_edge_dst": np.zeros(0, np.uint32),
        "cluster_idx": np.zeros(C, np.uint32),
        "cl_vertex_off": np.array(cl_vertex_off, np.uint64), "cl_var_off": np.array(cl_var_off, np.uint64),
        "v_seq_off": np.array(v_seq_off, np.uint64), "seq": np.concatenate(out["seq"]) if out["seq"] else np.zeros(0, np.uint8),
        "v_flags": np.array(out["v_flags"], np.uint8), "v_var": np.array(out["v_var"], np.uint16), "v_allele": np.array(out["v_allele"], np.uint16),
        "v_nested": np.full(len(out["v_flags"]), NONE32, np.uint32),
        "v_refvar_off": np.array(v_refvar_off, np.uint64), "v_refvar": np.array(out["v_refvar"], np.uint16),
        "v_in_off": np.array(v_in_off, np.uint64), "v_in_src": np.array(out["v_in_src"], np.uint32),
        "var_pos": np.array(out["var_pos"], np.uint32), "var_dep": np.array(out["var_dep"], np.uint8), "var_nalt": np.array(out["var_nalt"], np.uint16),
        "var_alt_off": np.array(var_alt_off, np.uint64), "alt_reflen": np.array(out["alt_reflen"], np.uint32),
        "alt_seq_off": np.array(alt_seq_off, np.uint64), "alt_seq": np.frombuffer(bytes(alt_seq), np.uint8),
        "cluster_order": np.array(order, np.int64),
    }


class _neg_str:
    """Sort key that orders strings descending (std::string operator>)."""

    def __init__(self, s):
        self.s = s

    def __lt__(self, o):
        return self.s > o.s

    def __eq__(self, o):
        return self.s == o.s


def intercluster_regions(reference_len: int, variants, k: int = K):
    """Gaps of >= k nucleotides between consecutive variants' reference spans, head and tail included
    (VariantFileParser::addSequenceToInterclusterRegions, VariantFileParser.cpp:171-183,470-545): (start, end) inclusive."""
    out = []
    prev_end = -1
    for v in variants:
        end = v.pos
        for alt in v.alts:
            r, _ = _right_trim(v.ref.upper(), alt.upper())
            end = max(end, v.pos + len(r) - 1)
        if v.pos > prev_end + 1 and (v.pos - 1) - (prev_end + 1) + 1 >= k:
            out.append((prev_end + 1, v.pos - 1))
        prev_end = max(prev_end, end)
    if reference_len - 1 >= prev_end + 1 and (reference_len - 1) - (prev_end + 1) + 1 >= k:
        out.append((prev_end + 1, reference_len - 1))
    return out
