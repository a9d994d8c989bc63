/*
 * btgpu.h — C ABI of libbtgpu.so: B200 (sm_100a) implementation of BayesTyper's
 * two hot paths (k-mer match, per-cluster Gibbs sampler).
 *
 * The reference (bioinformatics-centre/BayesTyper, /root/reference) has no FFI;
 * its seams are C++ class interfaces.  Each entry point below names the
 * reference interface it replaces (file:line relative to /root/reference).
 * INTEGRATION.md shows the reference-side binding for each.
 *
 * Conventions
 *   - return 0 on success, a negative BTG_E* code on failure; the message is
 *     available from btg_last_error() (thread-local).  Nothing here calls
 *     exit() (the reference does: src/kmerBloom/KmerBloom.cpp:67-71).
 *   - plain pointers and sizes only.  Pointers named *_dev are device pointers
 *     on the library's current device; all others are host pointers owned by
 *     the caller.  Opaque handles own their device memory.
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = the library's
 *     own stream); *_dev entry points are asynchronous on that stream, host
 *     entry points return after their results are in the caller's buffers.
 *   - packed k-mers are 2 x uint64_t each, in the in-memory layout of the
 *     reference's std::bitset<2k> (include/bayesTyper/Kmer.hpp:40-50):
 *     nucleotide i in bits [2i,2i+1] (A=0 C=1 G=2 T=3, include/bayesTyper/
 *     Nucleotide.hpp:39-70); word 0 = nucleotides 0..31.  A bitset<110> can be
 *     memcpy'd into this layout.
 *   - k is fixed at compile time in the reference (-DBT_KMER_SIZE, default 55,
 *     CMakeLists.txt:13); this library is built for BTG_KMER_SIZE = 55 and
 *     rejects any other k.
 */
#ifndef BTGPU_H
#define BTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define BTG_KMER_SIZE 55
#define BTG_MAX_SAMPLES 30 /* src/bayesTyper/main.cpp:72 */

#define BTG_OK 0
#define BTG_EINVAL (-1)  /* bad argument */
#define BTG_ECUDA (-2)   /* CUDA runtime error */
#define BTG_EIO (-3)     /* file error */
#define BTG_ENOMEM (-4)  /* allocation failed */
#define BTG_ESTATE (-5)  /* call sequence error */

/* ---- library ---------------------------------------------------------- */
int btg_init(int device);            /* select device, create the library stream */
void btg_shutdown(void);
const char *btg_last_error(void);
int btg_version(void);
int btg_device_sm_count(void);
/* pinned host memory for callers that want overlapped staging */
void *btg_host_alloc(size_t bytes);
void btg_host_free(void *p);
/* number of kernel launches issued by this library since btg_init / last reset */
uint64_t btg_launch_count(void);
void btg_launch_count_reset(void);

/* ---- KmerBloom  (include/kmerBloom/KmerBloom.hpp:48-77) ----------------- */
typedef struct btg_bloom btg_bloom;

/* KmerBloom(num_kmers, fpr)            src/kmerBloom/KmerBloom.cpp:53-60 */
btg_bloom *btg_bloom_create(uint64_t num_kmers, float fpr, int k);
/* KmerBloom(prefix): <prefix>.bloomMeta / .bloomData    KmerBloom.cpp:62-89 */
btg_bloom *btg_bloom_load(const char *prefix, int k);
/* same, from memory (meta fields + raw filter bytes, (num_bits+7)/8 of them) */
btg_bloom *btg_bloom_from_bytes(const uint8_t *data, uint64_t num_kmers, uint64_t num_bits, int k);
/* KmerBloom::save(prefix)              KmerBloom.cpp:148-164 */
int btg_bloom_save(const btg_bloom *b, const char *prefix);
int btg_bloom_info(const btg_bloom *b, uint64_t *num_kmers, uint64_t *num_bits, uint32_t *num_hashes);
/* raw filter bytes, BloomFilter::storeFilter layout (external/ntHash/BloomFilter.hpp:260-264) */
int btg_bloom_download(const btg_bloom *b, uint8_t *out, uint64_t nbytes);
/* KmerBloom::addKmer(bitset)           KmerBloom.cpp:178-182 -> BloomFilter::insertF */
int btg_bloom_insert(btg_bloom *b, const uint64_t *kmers, size_t n);
/* KmerBloom::lookup(bitset) const      KmerBloom.cpp:196-200 -> BloomFilter::containsF */
int btg_bloom_lookup(const btg_bloom *b, const uint64_t *kmers, size_t n, uint8_t *hit);
/* device-resident variants (inputs/outputs already in HBM) */
int btg_bloom_insert_dev(btg_bloom *b, const uint64_t *kmers_dev, size_t n, void *stream);
int btg_bloom_lookup_dev(const btg_bloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, void *stream);
/* same, additionally writing the number of probes the reference's early-exit
 * loop would have executed per k-mer (BloomFilter.hpp:149-161) — the unit of
 * the algorithmic-bytes model.  probes_dev may be NULL. */
int btg_bloom_lookup_probes_dev(const btg_bloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev,
                                uint8_t *probes_dev, void *stream);
void btg_bloom_free(btg_bloom *b);

/* ---- ThreadedKmerBloom (KmerBloom.hpp:79-108; KmerBloom.cpp:203-286) ----- *
 * 65,536 independent sub-filters selected by NTP64(kmer,k,1029283129)%65536.
 * The reference's per-sub-filter mutex (getKmerLock) exists only to make
 * test-then-set atomic; btg_tbloom_test_and_insert gives that operation
 * directly (first occurrence in call order wins, as under the lock).        */
typedef struct btg_tbloom btg_tbloom;
btg_tbloom *btg_tbloom_create(uint64_t num_kmers, float fpr, int k);
int btg_tbloom_info(const btg_tbloom *b, uint64_t *sub_kmers, uint64_t *sub_bits, uint32_t *num_hashes);
int btg_tbloom_insert(btg_tbloom *b, const uint64_t *kmers, size_t n);
int btg_tbloom_lookup(const btg_tbloom *b, const uint64_t *kmers, size_t n, uint8_t *hit);
int btg_tbloom_insert_dev(btg_tbloom *b, const uint64_t *kmers_dev, size_t n, void *stream);
int btg_tbloom_lookup_dev(const btg_tbloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, void *stream);
int btg_tbloom_download(const btg_tbloom *b, uint8_t *out, uint64_t nbytes); /* 65536 * ((sub_bits+7)/8) bytes */
void btg_tbloom_free(btg_tbloom *b);

/* ---- k-mer primitives (Kmer.tpp:182-255; nthash.hpp:262-282) ------------- */
/* NTP64 of each packed k-mer as given (no canonicalisation) */
int btg_kmer_hash(const uint64_t *kmers, size_t n, uint64_t *hash_out);
/* KmerPair::getLexicographicalLowestKmer of each packed k-mer */
int btg_kmer_canonical(const uint64_t *kmers, size_t n, uint64_t *canon_out);
/* Rolling enumeration over an ASCII nucleotide string exactly as KmerPair::move
 * drives it (non-ACGT resets): writes the canonical k-mer ending at every
 * position p >= k-1 whose window is all-ACGT; valid_out[p]=1 there, else 0 and
 * the k-mer slot is zeroed.  kmers_out holds 2*len words.  Follows the scan
 * loops of KmerCounter.cpp:291-334 and VariantClusterGraph.cpp:800-846. */
int btg_scan_sequence(const char *seq, size_t len, uint64_t *kmers_out, uint8_t *valid_out);
int btg_scan_sequence_dev(const char *seq_dev, size_t len, uint64_t *kmers_out_dev, uint8_t *valid_out_dev, void *stream);
/* fused scan + Bloom probe (the a10 genome scan against a filter): hit_out[p]=1
 * iff the window ending at p is valid and its canonical k-mer is in b */
int btg_scan_sequence_lookup_dev(const btg_bloom *b, const char *seq_dev, size_t len, uint8_t *hit_out_dev, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BTGPU_H */
