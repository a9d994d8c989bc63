/*
 * kmer_oracle.c — CPU restatement of the reference's k-mer primitives.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker.
 *
 * Parity status: PINNED. Checked in tests/test_oracle_kmer.py against the
 * known-answer vectors of SURVEY.md §4 (generated from the unmodified
 * reference sources) and against oracle/_ref (the reference TUs compiled in
 * place) when it is built.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 *
 * Packed k-mer layout used at the boundary (= the in-memory layout of the
 * reference's std::bitset<2k>, include/bayesTyper/Kmer.hpp): nucleotide i
 * occupies bits [2i, 2i+1] of a little-endian multi-word integer; word 0 holds
 * nucleotides 0..31, word 1 holds 32..k-1.  Code: A=0 C=1 G=2 T=3 with bit 2i
 * the low bit (include/bayesTyper/Nucleotide.hpp:39-70).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* external/ntHash/nthash.hpp:18-28 */
static const int      multiShift = 27;
static const uint64_t multiSeed  = 0x90b45d39fb6da1faULL;
static const uint64_t seedTab4[4] = {
    0x3c8bfbb395c60474ULL, /* A */
    0x3193c18562a02b4cULL, /* C */
    0x20323ed082572324ULL, /* G */
    0x295549f54be24456ULL  /* T */
};

static inline uint64_t rol64(uint64_t v, unsigned s) {
    s &= 63u;
    return s ? (v << s) | (v >> (64 - s)) : v;
}

/* Nucleotide::ntToBit<1>, include/bayesTyper/Nucleotide.hpp:39-70.
 * Returns 0..3, or -1 for a non-ACGT character (which resets the k-mer). */
int bto_nt_code(char nt) {
    switch (nt) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

static inline unsigned get_nt(const uint64_t *km, unsigned i) {
    return (unsigned)((km[i >> 5] >> ((i & 31u) * 2u)) & 3u);
}
static inline void set_nt(uint64_t *km, unsigned i, unsigned c) {
    km[i >> 5] &= ~(3ULL << ((i & 31u) * 2u));
    km[i >> 5] |= ((uint64_t)c) << ((i & 31u) * 2u);
}

/* Nucleotide::ntToBit<k>(string), Nucleotide.hpp:72-94. Returns 0 on success. */
int bto_pack_kmer(const char *seq, unsigned k, uint64_t *out2) {
    out2[0] = out2[1] = 0;
    for (unsigned i = 0; i < k; i++) {
        int c = bto_nt_code(seq[i]);
        if (c < 0) return -1;
        set_nt(out2, i, (unsigned)c);
    }
    return 0;
}

/* Nucleotide::bitToNt<k>, Nucleotide.hpp:96-129 (and KmerBloom.cpp:97-129). */
void bto_unpack_kmer(const uint64_t *km, unsigned k, char *out) {
    static const char tab[4] = {'A', 'C', 'G', 'T'};
    for (unsigned i = 0; i < k; i++) out[i] = tab[get_nt(km, i)];
    out[k] = 0;
}

/* Reverse complement as KmerReverseComplement builds it (Kmer.tpp:116-146):
 * rc[k-1-i] = ~fw[i]. */
void bto_revcomp(const uint64_t *km, unsigned k, uint64_t *out2) {
    out2[0] = out2[1] = 0;
    for (unsigned i = 0; i < k; i++) set_nt(out2, k - 1 - i, 3u - get_nt(km, i));
}

/* KmerPair::getLexicographicalLowestKmer, Kmer.tpp:226-255: compare from
 * nucleotide 0, high bit first then low bit (== compare 2-bit codes,
 * A<C<G<T); first difference decides; full tie -> forward.
 * Returns 1 when the forward k-mer is the canonical one. */
int bto_forward_is_canonical(const uint64_t *fw, const uint64_t *rc, unsigned k) {
    for (unsigned i = 0; i < k; i++) {
        unsigned a = get_nt(fw, i), b = get_nt(rc, i);
        if (a < b) return 1;
        if (a > b) return 0;
    }
    return 1;
}

void bto_canonical(const uint64_t *km, unsigned k, uint64_t *out2) {
    uint64_t rc[2];
    bto_revcomp(km, k, rc);
    if (bto_forward_is_canonical(km, rc, k)) { out2[0] = km[0]; out2[1] = km[1]; }
    else { out2[0] = rc[0]; out2[1] = rc[1]; }
}

/* NTP64(seq,k), external/ntHash/nthash.hpp:262-267 (msTab[c][j] == rol(seed[c], j),
 * nthash.hpp:30-118). Operates on the packed k-mer, i.e. on bitToNt(kmer). */
uint64_t bto_ntp64(const uint64_t *km, unsigned k) {
    uint64_t h = 0;
    for (unsigned i = 0; i < k; i++) h ^= rol64(seedTab4[get_nt(km, i)], (k - 1 - i) % 64);
    return h;
}

/* NTP64(seq,k,seed), nthash.hpp:275-282. */
uint64_t bto_ntp64_seeded(const uint64_t *km, unsigned k, unsigned seed) {
    uint64_t h = bto_ntp64(km, k);
    h *= (uint64_t)seed ^ ((uint64_t)k * multiSeed);
    h ^= h >> multiShift;
    return h;
}

/* getRhval, nthash.hpp:200-205: forward hash of the reverse-complement string. */
uint64_t bto_nt_rhval(const uint64_t *km, unsigned k) {
    uint64_t h = 0;
    for (unsigned i = 0; i < k; i++) h ^= rol64(seedTab4[3u - get_nt(km, i)], i);
    return h;
}

/* KmerBloom::calcOptNumBloomBits, src/kmerBloom/KmerBloom.cpp:132-138.
 * NOTE the types: fpr is float, std::log(float) is float, uint64*float is
 * float; ln2 = std::log(2) is double (integer overload). */
uint64_t bto_bloom_num_bits(uint64_t num_kmers, float fpr) {
    double ln2 = log(2.0);
    float prod = (float)num_kmers * logf(fpr);
    return (uint64_t)ceil(-((double)prod / ln2 / ln2));
}

/* KmerBloom::calcOptNumHashes, KmerBloom.cpp:140-146. */
unsigned bto_bloom_num_hashes(uint64_t num_bits, uint64_t num_kmers) {
    double frac = (double)num_bits / (double)num_kmers;
    return (unsigned)ceil(frac * log(2.0));
}

/* KmerBloom(num_kmers, fpr) ctor, KmerBloom.cpp:53-60: num_kmers = max(n,1). */
void bto_bloom_params(uint64_t num_kmers_in, float fpr, uint64_t *num_kmers, uint64_t *num_bits, unsigned *num_hashes) {
    uint64_t n = num_kmers_in ? num_kmers_in : 1;
    *num_kmers = n;
    *num_bits = bto_bloom_num_bits(n, fpr);
    *num_hashes = bto_bloom_num_hashes(*num_bits, n);
}

/* ThreadedKmerBloom ctor, KmerBloom.cpp:204-216: per-sub-filter n is
 * std::ceil(num_kmers / static_cast<float>(65536)) converted to uint64. */
uint64_t bto_threaded_bloom_sub_kmers(uint64_t num_kmers) {
    return (uint64_t)ceilf((float)num_kmers / 65536.0f);
}

/* ThreadedKmerBloom::rootIndex, KmerBloom.cpp:275-279. */
unsigned bto_threaded_bloom_root(const uint64_t *km, unsigned k) {
    return (unsigned)(bto_ntp64_seeded(km, k, 1029283129u) % 65536ULL);
}

/* Probe bit locations, BloomFilter::containsF / insertF,
 * external/ntHash/BloomFilter.hpp:149-161, 56-66. */
void bto_bloom_locs(const uint64_t *km, unsigned k, uint64_t m, unsigned nh, uint64_t *locs) {
    uint64_t h = bto_ntp64(km, k);
    locs[0] = h % m;
    for (unsigned i = 1; i < nh; i++) {
        uint64_t mh = h * ((uint64_t)i ^ ((uint64_t)k * multiSeed));
        mh ^= mh >> multiShift;
        locs[i] = mh % m;
    }
}

/* BloomFilter::insertF, BloomFilter.hpp:56-66: bit loc -> byte loc/8, mask 1<<(7-loc%8). */
void bto_bloom_insert(uint8_t *filter, uint64_t m, unsigned nh, unsigned k, const uint64_t *kmers, size_t n) {
    uint64_t locs[64];
    for (size_t j = 0; j < n; j++) {
        bto_bloom_locs(kmers + 2 * j, k, m, nh, locs);
        for (unsigned i = 0; i < nh; i++) filter[locs[i] / 8] |= (uint8_t)(1u << (7 - locs[i] % 8));
    }
}

/* BloomFilter::containsF, BloomFilter.hpp:149-161. probes_out (optional) gets the
 * number of probes executed under the reference's early-exit semantics. */
void bto_bloom_lookup(const uint8_t *filter, uint64_t m, unsigned nh, unsigned k, const uint64_t *kmers, size_t n,
                      uint8_t *hit, uint8_t *probes_out) {
    uint64_t locs[64];
    for (size_t j = 0; j < n; j++) {
        bto_bloom_locs(kmers + 2 * j, k, m, nh, locs);
        unsigned i = 0;
        uint8_t ok = 1;
        for (; i < nh; i++) {
            if ((filter[locs[i] / 8] & (1u << (7 - locs[i] % 8))) == 0) { ok = 0; i++; break; }
        }
        hit[j] = ok;
        if (probes_out) probes_out[j] = (uint8_t)i;
    }
}

/* Rolling enumeration of canonical k-mers over a nucleotide string, as
 * KmerPair::move drives it (Kmer.tpp:44-74,116-146,182-190): a non-ACGT
 * character resets the pair; a k-mer is emitted for every position at which the
 * pair is complete.  This is the loop of KmerCounter::countInterclusterKmersCallback
 * (src/bayesTyper/KmerCounter.cpp:291-334) and VariantClusterGraph::countPathKmers
 * (src/bayesTyper/VariantClusterGraph.cpp:800-846).
 * out: canonical k-mers (2 words each), out_pos: index of the LAST nucleotide of
 * each emitted k-mer. Returns the number of k-mers emitted (<= cap). */
size_t bto_scan_sequence(const char *seq, size_t len, unsigned k, uint64_t *out, uint32_t *out_pos, size_t cap) {
    uint64_t fw[2] = {0, 0}, rc[2] = {0, 0};
    unsigned filled = 0;
    size_t n = 0;
    for (size_t p = 0; p < len; p++) {
        int c = bto_nt_code(seq[p]);
        if (c < 0) { filled = 0; continue; }
        if (filled == k) {
            /* forward: >>= 2, write top; reverse complement: <<= 2, write ~bits at bottom */
            fw[0] = (fw[0] >> 2) | (fw[1] << 62);
            fw[1] >>= 2;
            set_nt(fw, k - 1, (unsigned)c);
            rc[1] = (rc[1] << 2) | (rc[0] >> 62);
            rc[0] <<= 2;
            rc[1] &= (1ULL << (k * 2 - 64)) - 1; /* 32 < k < 64: two words */
            set_nt(rc, 0, 3u - (unsigned)c);
        } else {
            set_nt(fw, filled, (unsigned)c);
            set_nt(rc, k - 1 - filled, 3u - (unsigned)c);
            filled++;
        }
        if (filled == k) {
            if (n < cap) {
                if (bto_forward_is_canonical(fw, rc, k)) { out[2 * n] = fw[0]; out[2 * n + 1] = fw[1]; }
                else { out[2 * n] = rc[0]; out[2 * n + 1] = rc[1]; }
                if (out_pos) out_pos[n] = (uint32_t)p;
            }
            n++;
        }
    }
    return n;
}
