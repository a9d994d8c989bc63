// kmer.cuh — device primitives of the k-mer match path (sm_100a).
//
// What the reference does with std::bitset<110> bit loops, a heap std::string
// per lookup and 55 table loads per hash (include/bayesTyper/Kmer.tpp:182-255,
// src/kmerBloom/KmerBloom.cpp:97-129, external/ntHash/nthash.hpp:262-267) is
// done here on a 128-bit register pair:
//
//   internal form  V = sum_i code(nt_i) << 2*(K-1-i)   ("MSB-first": nt 0 on top)
//   - lexicographic order of k-mer strings == integer order of V, so
//     KmerPair::getLexicographicalLowestKmer is one 128-bit compare;
//   - the boundary layout (bitset<2K>, nt i at bits [2i,2i+1]) is V with its
//     2-bit groups reversed: two BREV + a pair swap;
//   - NTP64 is linear over XOR, so it is evaluated 4 nucleotides at a time from
//     a 256-entry shared-memory table T[b] = XOR_t rol(seed[(b>>2t)&3], t):
//         h = XOR_g rol(T[byte_g(V)], 4g) ^ pad_const
//     (14 LDS.64 instead of 55 dependent table loads), and rolled in O(1) per
//     nucleotide when scanning a sequence.
#pragma once
#include <cstdint>

// The arithmetic below is plain integer code; it is marked host+device so that
// tests/test_kmer_header_host.py can compile this header with g++ and check it
// against the oracle without a GPU.  The product only ever runs it on the device.
#ifdef __CUDACC__
#define BTG_HD __host__ __device__ __forceinline__
#else
#define BTG_HD inline
#endif

namespace btg {

BTG_HD uint64_t brev64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((x & 0x0f0f0f0f0f0f0f0fULL) << 4);
    return __builtin_bswap64(x);
#endif
}
BTG_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
#ifdef __CUDACC__
// experiment knob (BTG_PROBE_MODE): which load flavour the single-byte Bloom probes use
__device__ int g_probe_mode = 0;
#endif
BTG_HD uint8_t load_byte(const uint8_t *p) {
#ifdef __CUDA_ARCH__
    unsigned v;
    switch (g_probe_mode) {
        case 1: asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (uint8_t)v;
        case 2: asm volatile("ld.global.ca.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (uint8_t)v;
        case 3: asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (uint8_t)v;
        case 4: asm volatile("ld.global.cv.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (uint8_t)v;
        default: return __ldg(p);
    }
#else
    return *p;
#endif
}

constexpr int K = 55;  // BT_KMER_SIZE (CMakeLists.txt:13)
static_assert(K > 32 && K < 64, "two-word k-mers only");

constexpr uint64_t kSeedA = 0x3c8bfbb395c60474ULL;  // external/ntHash/nthash.hpp:24-27
constexpr uint64_t kSeedC = 0x3193c18562a02b4cULL;
constexpr uint64_t kSeedG = 0x20323ed082572324ULL;
constexpr uint64_t kSeedT = 0x295549f54be24456ULL;
constexpr uint64_t kMultiSeed = 0x90b45d39fb6da1faULL;  // nthash.hpp:21
constexpr int kMultiShift = 27;                         // nthash.hpp:18
constexpr uint64_t kKmul = (uint64_t)K * kMultiSeed;    // "k * multiSeed" in nthash.hpp:280, BloomFilter.hpp:154
constexpr unsigned kThreadedSeed = 1029283129u;         // KmerBloom.cpp:278
constexpr unsigned kThreadedRoots = 65536u;             // KmerBloom.cpp:206

constexpr uint64_t kHiMask = (1ULL << (2 * K - 64)) - 1;  // valid bits of the high word
constexpr int kPadBits = 128 - 2 * K;                      // 18

BTG_HD uint64_t rol64(uint64_t v, int s) {
    s &= 63;
    return s ? (v << s) | (v >> (64 - s)) : v;
}
BTG_HD uint64_t ror64(uint64_t v, int s) { return rol64(v, 64 - (s & 63)); }

BTG_HD uint64_t seed_of(unsigned c) {
    return c == 0 ? kSeedA : (c == 1 ? kSeedC : (c == 2 ? kSeedG : kSeedT));
}

// XOR_{r=K}^{4*nbytes-1} rol(seedA, r): the contribution of the zero padding
// above the top nucleotide in the last (partial) byte of V.
constexpr uint64_t pad_const() {
    uint64_t c = 0;
    for (int r = K; r < 4 * ((2 * K + 7) / 8); r++) {
        int s = r & 63;
        c ^= s ? ((kSeedA << s) | (kSeedA >> (64 - s))) : kSeedA;
    }
    return c;
}
constexpr uint64_t kPadConst = pad_const();
constexpr int kNumBytes = (2 * K + 7) / 8;  // 14

struct Kmer128 {
    uint64_t hi, lo;  // V = hi:lo, MSB-first
};

// reverse the order of the 2-bit groups of a 64-bit word
BTG_HD uint64_t rev2_64(uint64_t x) {
    x = brev64(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

// boundary layout (w0 = nts 0..31, w1 = nts 32..K-1, nt i at bits 2i) -> internal
BTG_HD Kmer128 from_boundary(uint64_t w0, uint64_t w1) {
    uint64_t r0 = rev2_64(w0), r1 = rev2_64(w1);  // 128-bit reverse = (r0 : r1)
    Kmer128 v;
    v.hi = r0 >> kPadBits;
    v.lo = (r0 << (64 - kPadBits)) | (r1 >> kPadBits);
    return v;
}
BTG_HD void to_boundary(const Kmer128 &v, uint64_t &w0, uint64_t &w1) {
    // inverse of the above: reverse (V << pad)
    uint64_t hi = (v.hi << kPadBits) | (v.lo >> (64 - kPadBits));
    uint64_t lo = v.lo << kPadBits;
    w0 = rev2_64(hi);
    w1 = rev2_64(lo);
}

// reverse complement (Kmer.tpp:116-146 builds it incrementally: rc[K-1-i] = ~fw[i])
BTG_HD Kmer128 revcomp(const Kmer128 &v) {
    uint64_t hi = ~((v.hi << kPadBits) | (v.lo >> (64 - kPadBits)));
    uint64_t lo = ~(v.lo << kPadBits);  // low 18 bits become ones; they reverse to the top and are shifted out
    uint64_t r_hi = rev2_64(lo), r_lo = rev2_64(hi);  // reversed 128-bit value r_hi:r_lo
    // the 110 payload bits now sit in the low 110 bits of r_hi:r_lo except the
    // 18 pad ones on top of r_hi
    Kmer128 o;
    o.hi = r_hi & kHiMask;
    o.lo = r_lo;
    return o;
}

// KmerPair::getLexicographicalLowestKmer (Kmer.tpp:226-255); tie -> forward
BTG_HD bool forward_is_canonical(const Kmer128 &f, const Kmer128 &r) {
    return (f.hi < r.hi) || (f.hi == r.hi && f.lo <= r.lo);
}

// ---- ntHash ----------------------------------------------------------------
// T[b] = XOR_{t=0..3} rol(seed[(b >> 2t) & 3], t), built once per CTA
BTG_HD uint64_t hash_table_entry(unsigned b) {
    uint64_t v = 0;
    for (int t = 0; t < 4; t++) v ^= rol64(seed_of((b >> (2 * t)) & 3u), t);
    return v;
}
#ifdef __CUDACC__
__device__ __forceinline__ void build_hash_table(uint64_t *T /* [256] shared */) {
    for (unsigned b = threadIdx.x; b < 256; b += blockDim.x) {
        T[b] = hash_table_entry(b);
    }
}
#endif

// NTP64(bitToNt(kmer), K)  (nthash.hpp:262-267)
BTG_HD uint64_t ntp64(const Kmer128 &v, const uint64_t *T) {
    uint64_t h = kPadConst;
#pragma unroll
    for (int g = 0; g < 8; g++) h ^= rol64(T[(v.lo >> (8 * g)) & 0xff], 4 * g);
#pragma unroll
    for (int g = 8; g < kNumBytes; g++) h ^= rol64(T[(v.hi >> (8 * (g - 8))) & 0xff], 4 * g);
    return h;
}

// NTP64(seq,k,seed)  (nthash.hpp:275-282)
BTG_HD uint64_t ntp64_seeded(uint64_t h, unsigned seed) {
    h *= (uint64_t)seed ^ kKmul;
    h ^= h >> kMultiShift;
    return h;
}

// rolling state over a sequence: forward/revcomp registers + the two hashes
//   F = NTP64(forward window), R = NTP64(reverse-complement string) (= getRhval, nthash.hpp:200-205)
struct Roller {
    Kmer128 f, r;
    uint64_t F, R;
    int filled;  // nucleotides currently in the window (<= K)

    BTG_HD void reset() {
        f.hi = f.lo = r.hi = r.lo = 0;
        F = R = 0;
        filled = 0;
    }
    // push one nucleotide code (0..3).  Returns true when the window is complete
    // (KmerPair::move, Kmer.tpp:182-190).
    BTG_HD bool push(unsigned c) {
        const unsigned cc = 3u - c;
        if (filled == K) {
            const unsigned out = (unsigned)(f.hi >> (2 * K - 64 - 2)) & 3u;  // nt leaving
            // F' = rol1(F) ^ rol(seed[out],K) ^ seed[in]              (nthash.hpp:270-272)
            F = rol64(F, 1) ^ rol64(seed_of(out), K) ^ seed_of(c);
            // R' = ror1(R ^ seed[comp(out)]) ^ rol(seed[comp(in)],K-1) (algebra of nthash.hpp:243-247)
            R = ror64(R ^ seed_of(3u - out), 1) ^ rol64(seed_of(cc), K - 1);
            f.hi = ((f.hi << 2) | (f.lo >> 62)) & kHiMask;
            f.lo = (f.lo << 2) | c;
            r.lo = (r.lo >> 2) | (r.hi << 62);
            r.hi = (r.hi >> 2) | ((uint64_t)cc << (2 * K - 64 - 2));
        } else {
            // partial window: same recurrences without an outgoing nucleotide
            F = rol64(F, 1) ^ seed_of(c);
            R = R ^ rol64(seed_of(cc), filled);
            f.hi = ((f.hi << 2) | (f.lo >> 62)) & kHiMask;
            f.lo = (f.lo << 2) | c;
            // rc gets comp(c) at string index K-1-filled.. built so that when
            // complete the string is comp(nt_{K-1})..comp(nt_0): shift right and
            // insert on top, exactly as in the complete case
            r.lo = (r.lo >> 2) | (r.hi << 62);
            r.hi = (r.hi >> 2) | ((uint64_t)cc << (2 * K - 64 - 2));
            filled++;
        }
        return filled == K;
    }
    BTG_HD bool fwd_canonical() const { return forward_is_canonical(f, r); }
    BTG_HD Kmer128 canonical() const { return fwd_canonical() ? f : r; }
    BTG_HD uint64_t canonical_hash() const { return fwd_canonical() ? F : R; }
};

// ---- Bloom probes -----------------------------------------------------------
struct BloomView {
    const uint8_t *bits;  // (m+7)/8 bytes, bit loc at byte loc/8 mask 1<<(7-loc%8)  (BloomFilter.hpp:149-161)
    uint64_t m;           // filter size in bits
    uint64_t magic;       // floor((2^64-1)/m) for the remainder
    uint32_t nh;          // number of hash functions
};

// h % m without a 64-bit divide: q = mulhi(h, floor((2^64-1)/m)) is at most 2 short
BTG_HD uint64_t mod_m(uint64_t h, uint64_t m, uint64_t magic) {
    uint64_t q = mulhi64(h, magic);
    uint64_t r = h - q * m;
    while (r >= m) r -= m;
    return r;
}

BTG_HD uint64_t probe_loc(const BloomView &b, uint64_t h, unsigned i) {
    if (i == 0) return mod_m(h, b.m, b.magic);
    uint64_t mh = h * ((uint64_t)i ^ kKmul);  // BloomFilter.hpp:154
    mh ^= mh >> kMultiShift;
    return mod_m(mh, b.m, b.magic);
}

BTG_HD bool probe_bit(const BloomView &b, uint64_t loc) {
    return (load_byte(b.bits + (loc >> 3)) >> (7u - (unsigned)(loc & 7u))) & 1u;
}

// BloomFilter::containsF.  Probes are issued in independent groups (2, then 4s)
// so several DRAM sectors are in flight per thread; the boolean result equals
// the reference's sequential early-exit loop.  *probes (optional) = number of
// probes that loop would have executed.
BTG_HD bool bloom_contains(const BloomView &b, uint64_t h, unsigned *probes = nullptr) {
    unsigned i = 0;
    unsigned first_zero = b.nh;  // index of first zero bit
    // group 0: two probes
    {
        unsigned n = b.nh < 2 ? b.nh : 2;
        bool b0 = probe_bit(b, probe_loc(b, h, 0));
        bool b1 = n > 1 ? probe_bit(b, probe_loc(b, h, 1)) : true;
        if (!b0) first_zero = 0;
        else if (!b1) first_zero = 1;
        i = n;
    }
    while (first_zero == b.nh && i < b.nh) {
        unsigned n = b.nh - i < 4 ? b.nh - i : 4;
        bool bb[4];
#pragma unroll
        for (unsigned j = 0; j < 4; j++) bb[j] = j < n ? probe_bit(b, probe_loc(b, h, i + j)) : true;
#pragma unroll
        for (int j = 3; j >= 0; j--)
            if (!bb[j]) first_zero = i + j;
        i += n;
    }
    if (probes) *probes = first_zero == b.nh ? b.nh : first_zero + 1;
    return first_zero == b.nh;
}

#ifdef __CUDACC__
__device__ __forceinline__ void bloom_insert(uint8_t *bits, const BloomView &b, uint64_t h) {
    for (unsigned i = 0; i < b.nh; i++) {
        uint64_t loc = probe_loc(b, h, i);
        // byte-wise OR done on the containing aligned 32-bit word (little endian)
        uint64_t byte = loc >> 3;
        unsigned *word = (unsigned *)(bits + (byte & ~3ULL));
        unsigned mask = (1u << (7u - (unsigned)(loc & 7u))) << (8u * (unsigned)(byte & 3u));
        atomicOr(word, mask);
    }
}
#endif

// nucleotide char -> code, 4 = invalid (Nucleotide::ntToBit<1>, Nucleotide.hpp:39-70)
BTG_HD unsigned nt_code(char ch) {
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

}  // namespace btg
