// bloom.cu — KmerBloom / ThreadedKmerBloom on the device, plus the k-mer
// primitive entry points (hash, canonical, rolling scan).
//
// Replaces include/kmerBloom/KmerBloom.hpp:48-108 (src/kmerBloom/KmerBloom.cpp) and
// external/ntHash/BloomFilter.hpp:40-66,149-161,260-264.
//
// Kernels are HBM-bound byte/integer work: 16 B of packed k-mer per unit read
// with one coalesced 128-bit load, then up to nh single-bit probes, each a
// random 32 B sector.  Grids are persistent-style: SM count x resident CTAs,
// grid-stride over the k-mers.
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"
#include "kmer.cuh"

using namespace btg;

struct btg_bloom {
    uint8_t *bits = nullptr;  // device, padded to 16 B
    uint64_t num_kmers = 0;
    uint64_t num_bits = 0;
    uint32_t num_hashes = 0;
    uint64_t nbytes() const { return (num_bits + 7) / 8; }
    BloomView view() const { return BloomView{bits, num_bits, ~0ULL / num_bits, num_hashes}; }
};

struct btg_tbloom {
    uint8_t *bits = nullptr;  // device: kThreadedRoots sub-filters, stride sub_stride bytes
    uint64_t sub_kmers = 0, sub_bits = 0, sub_stride = 0;
    uint32_t num_hashes = 0;
};

namespace {

// KmerBloom::calcOptNumBloomBits (KmerBloom.cpp:132-138): uint64*float product in
// float, divided by double ln2 twice.
uint64_t opt_num_bits(float fpr, uint64_t num_kmers) {
    double ln2 = std::log(2.0);
    float prod = static_cast<float>(num_kmers) * std::log(fpr);
    return static_cast<uint64_t>(std::ceil(-(static_cast<double>(prod) / ln2 / ln2)));
}
// KmerBloom::calcOptNumHashes (KmerBloom.cpp:140-146)
uint32_t opt_num_hashes(uint64_t num_bits, uint64_t num_kmers) {
    double frac = static_cast<double>(num_bits) / static_cast<double>(num_kmers);
    return static_cast<uint32_t>(std::ceil(frac * std::log(2.0)));
}

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) k_bloom_lookup(BloomView b, const ulonglong2 *__restrict__ kmers, size_t n,
                                                         uint8_t *__restrict__ hit, uint8_t *__restrict__ probes) {
    __shared__ uint64_t T[256];
    build_hash_table(T);
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 w = __ldg(kmers + i);
        Kmer128 v = from_boundary(w.x, w.y);
        uint64_t h = ntp64(v, T);
        unsigned np;
        bool ok = bloom_contains(b, h, &np);
        hit[i] = ok;
        if (probes) probes[i] = (uint8_t)np;
    }
}

__global__ void __launch_bounds__(kBlock) k_bloom_insert(BloomView b, uint8_t *bits, const ulonglong2 *__restrict__ kmers, size_t n) {
    __shared__ uint64_t T[256];
    build_hash_table(T);
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 w = __ldg(kmers + i);
        Kmer128 v = from_boundary(w.x, w.y);
        bloom_insert(bits, b, ntp64(v, T));
    }
}

struct TBloomView {
    uint8_t *bits;
    uint64_t sub_bits, sub_stride, magic;
    uint32_t nh;
};

__device__ __forceinline__ BloomView sub_view(const TBloomView &t, uint64_t h) {
    unsigned root = (unsigned)(ntp64_seeded(h, kThreadedSeed) % kThreadedRoots);  // KmerBloom.cpp:275-279
    return BloomView{t.bits + (uint64_t)root * t.sub_stride, t.sub_bits, t.magic, t.nh};
}

__global__ void __launch_bounds__(kBlock) k_tbloom_lookup(TBloomView t, const ulonglong2 *__restrict__ kmers, size_t n, uint8_t *__restrict__ hit) {
    __shared__ uint64_t T[256];
    build_hash_table(T);
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 w = __ldg(kmers + i);
        uint64_t h = ntp64(from_boundary(w.x, w.y), T);
        BloomView b = sub_view(t, h);
        hit[i] = bloom_contains(b, h);
    }
}

__global__ void __launch_bounds__(kBlock) k_tbloom_insert(TBloomView t, const ulonglong2 *__restrict__ kmers, size_t n) {
    __shared__ uint64_t T[256];
    build_hash_table(T);
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 w = __ldg(kmers + i);
        uint64_t h = ntp64(from_boundary(w.x, w.y), T);
        BloomView b = sub_view(t, h);
        bloom_insert(const_cast<uint8_t *>(b.bits), b, h);
    }
}

__global__ void __launch_bounds__(kBlock) k_kmer_hash(const ulonglong2 *__restrict__ kmers, size_t n, uint64_t *__restrict__ out) {
    __shared__ uint64_t T[256];
    build_hash_table(T);
    __syncthreads();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 w = __ldg(kmers + i);
        out[i] = ntp64(from_boundary(w.x, w.y), T);
    }
}

__global__ void __launch_bounds__(kBlock) k_kmer_canonical(const ulonglong2 *__restrict__ kmers, size_t n, ulonglong2 *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 w = __ldg(kmers + i);
        Kmer128 f = from_boundary(w.x, w.y);
        Kmer128 r = revcomp(f);
        Kmer128 c = forward_is_canonical(f, r) ? f : r;
        uint64_t w0, w1;
        to_boundary(c, w0, w1);
        out[i] = make_ulonglong2(w0, w1);
    }
}

// Rolling scan.  Each thread owns kScanChunk consecutive end positions; it
// primes its window with the K-1 preceding characters, then rolls (O(1) per
// nucleotide).  MODE 0: write canonical k-mers + valid flags.  MODE 1: probe b.
constexpr int kScanChunk = 64;

template <int MODE>
__global__ void __launch_bounds__(kBlock) k_scan(const char *__restrict__ seq, size_t len, ulonglong2 *__restrict__ kmers_out,
                                                 uint8_t *__restrict__ flag_out, BloomView b) {
    const size_t nchunks = (len + kScanChunk - 1) / kScanChunk;
    for (size_t ch = blockIdx.x * (size_t)blockDim.x + threadIdx.x; ch < nchunks; ch += (size_t)gridDim.x * blockDim.x) {
        const size_t p0 = ch * kScanChunk;
        const size_t p1 = p0 + kScanChunk < len ? p0 + kScanChunk : len;
        Roller roll;
        roll.reset();
        size_t start = p0 >= (size_t)(K - 1) ? p0 - (K - 1) : 0;
        for (size_t p = start; p < p0; p++) {
            unsigned c = nt_code(__ldg(seq + p));
            if (c > 3) roll.reset();
            else roll.push(c);
        }
        for (size_t p = p0; p < p1; p++) {
            unsigned c = nt_code(__ldg(seq + p));
            bool complete = false;
            if (c > 3) roll.reset();
            else complete = roll.push(c);
            if (MODE == 0) {
                uint64_t w0 = 0, w1 = 0;
                if (complete) {
                    Kmer128 cn = roll.canonical();
                    to_boundary(cn, w0, w1);
                }
                kmers_out[p] = make_ulonglong2(w0, w1);
                flag_out[p] = complete;
            } else {
                flag_out[p] = complete ? bloom_contains(b, roll.canonical_hash()) : 0;
            }
        }
    }
}

// ---- host staging ------------------------------------------------------------
// Streams n packed k-mers from caller memory through two device buffers so the
// H2D copy of chunk i+1 overlaps the kernel on chunk i; `launch` enqueues the
// kernel for one chunk; `drain` (optional) copies one chunk of results back.
constexpr size_t kStageKmers = 8u << 20;  // 8 Mi k-mers = 128 MiB per buffer

template <class Launch>
int staged(const uint64_t *kmers, size_t n, size_t out_bytes_per_kmer, void *host_out, Launch launch) {
    if (n == 0) return BTG_OK;
    auto &c = ctx();
    const size_t chunk = n < kStageKmers ? n : kStageKmers;
    const int nslots = n > chunk ? 2 : 1;
    cudaStream_t streams[2] = {c.stream, c.copy_stream};  // slot i runs H2D -> kernel -> D2H in-order on streams[i]
    ulonglong2 *d_in[2] = {nullptr, nullptr};
    uint8_t *d_out[2] = {nullptr, nullptr};
    int rc = BTG_OK;
    cudaStreamSynchronize(c.stream);
    for (int i = 0; i < nslots; i++) {
        if (cudaMalloc(&d_in[i], chunk * sizeof(ulonglong2)) != cudaSuccess ||
            (out_bytes_per_kmer && cudaMalloc(&d_out[i], chunk * out_bytes_per_kmer) != cudaSuccess)) {
            set_error("staging allocation failed");
            rc = BTG_ENOMEM;
        }
    }
    size_t off = 0;
    int slot = 0;
    while (rc == BTG_OK && off < n) {
        size_t m = n - off < chunk ? n - off : chunk;
        cudaStream_t s = streams[slot];
        cudaError_t e = cudaMemcpyAsync(d_in[slot], kmers + 2 * off, m * sizeof(ulonglong2), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            launch(d_in[slot], m, d_out[slot], s);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && out_bytes_per_kmer)
            e = cudaMemcpyAsync((uint8_t *)host_out + off * out_bytes_per_kmer, d_out[slot], m * out_bytes_per_kmer, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) {
            set_error("staged k-mer pass failed: %s", cudaGetErrorString(e));
            rc = BTG_ECUDA;
            break;
        }
        off += m;
        slot = (slot + 1) % nslots;
    }
    for (int i = 0; i < nslots; i++) {
        cudaError_t e = cudaStreamSynchronize(streams[i]);
        if (rc == BTG_OK && e != cudaSuccess) {
            set_error("staged k-mer pass failed: %s", cudaGetErrorString(e));
            rc = BTG_ECUDA;
        }
    }
    for (int i = 0; i < nslots; i++) {
        cudaFree(d_in[i]);
        cudaFree(d_out[i]);
    }
    return rc;
}

int check_k(int k) {
    if (k != K) {
        set_error("k=%d unsupported: library built for BT_KMER_SIZE=%d", k, K);
        return BTG_EINVAL;
    }
    return BTG_OK;
}

btg_bloom *alloc_bloom(uint64_t num_kmers, uint64_t num_bits) {
    if (num_bits == 0 || num_kmers == 0) {
        set_error("empty Bloom filter (num_kmers=%llu num_bits=%llu)", (unsigned long long)num_kmers, (unsigned long long)num_bits);
        return nullptr;
    }
    auto *b = new btg_bloom();
    b->num_kmers = num_kmers;
    b->num_bits = num_bits;
    b->num_hashes = opt_num_hashes(num_bits, num_kmers);
    size_t padded = (b->nbytes() + 15) & ~size_t(15);
    if (cudaMalloc(&b->bits, padded) != cudaSuccess) {
        set_error("cudaMalloc(%zu) for Bloom filter failed", padded);
        delete b;
        return nullptr;
    }
    cudaMemsetAsync(b->bits, 0, padded, ctx().stream);
    cudaStreamSynchronize(ctx().stream);
    return b;
}

}  // namespace

namespace btg_internal {
BloomView bloom_view(const btg_bloom *b) { return b->view(); }
}  // namespace btg_internal

// experiment hook: select the probe load flavour (see kmer.cuh)
extern "C" __attribute__((visibility("default"))) int btg_debug_set_probe_mode(int mode) {
    return cudaMemcpyToSymbol(g_probe_mode, &mode, sizeof(int)) == cudaSuccess ? 0 : -2;
}

extern "C" {

btg_bloom *btg_bloom_create(uint64_t num_kmers_in, float fpr, int k) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (check_k(k)) return nullptr;
    if (!(fpr > 0.f && fpr < 1.f)) { set_error("fpr must be in (0,1)"); return nullptr; }
    uint64_t n = num_kmers_in ? num_kmers_in : 1;  // max(n, 1), KmerBloom.cpp:54
    return alloc_bloom(n, opt_num_bits(fpr, n));
}

btg_bloom *btg_bloom_from_bytes(const uint8_t *data, uint64_t num_kmers, uint64_t num_bits, int k) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (check_k(k)) return nullptr;
    btg_bloom *b = alloc_bloom(num_kmers, num_bits);
    if (!b) return nullptr;
    if (cudaMemcpy(b->bits, data, b->nbytes(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("upload of Bloom filter failed");
        btg_bloom_free(b);
        return nullptr;
    }
    return b;
}

btg_bloom *btg_bloom_load(const char *prefix, int k) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (check_k(k)) return nullptr;
    std::string p(prefix);
    std::ifstream meta(p + ".bloomMeta");
    if (!meta.is_open()) { set_error("Unable to open file %s.bloomMeta", prefix); return nullptr; }
    std::string line;
    std::getline(meta, line);
    std::vector<std::string> f;
    {
        std::stringstream ss(line);
        for (std::string item; std::getline(ss, item, '\t');) f.push_back(item);
    }
    if (f.size() != 3) { set_error("%s.bloomMeta: expected 3 tab-separated fields", prefix); return nullptr; }
    uint64_t num_kmers, num_bits;
    int file_k;
    try {
        num_kmers = std::stoull(f[0]);
        num_bits = std::stoull(f[1]);
        file_k = std::stoi(f[2]);
    } catch (...) { set_error("%s.bloomMeta: malformed", prefix); return nullptr; }
    if (file_k != k) { set_error("%s.bloomMeta: k=%d, expected %d", prefix, file_k, k); return nullptr; }
    std::ifstream data(p + ".bloomData", std::ios::in | std::ios::binary);
    if (!data.is_open()) { set_error("Unable to open file %s.bloomData", prefix); return nullptr; }
    btg_bloom *b = alloc_bloom(num_kmers, num_bits);
    if (!b) return nullptr;
    // stream the file through a pinned buffer in 64 MiB pieces
    const size_t piece = 64u << 20;
    uint8_t *buf = (uint8_t *)btg_host_alloc(piece);
    if (!buf) { btg_bloom_free(b); return nullptr; }
    uint64_t off = 0, total = b->nbytes();
    bool ok = true;
    while (off < total) {
        size_t m = total - off < piece ? total - off : piece;
        data.read((char *)buf, m);
        size_t got = data.gcount();  // the reference reads what is there and leaves the rest uninitialised; we zero-fill
        if (got && cudaMemcpy(b->bits + off, buf, got, cudaMemcpyHostToDevice) != cudaSuccess) { ok = false; break; }
        if (got < m) break;
        off += m;
    }
    btg_host_free(buf);
    if (!ok) { set_error("upload of %s.bloomData failed", prefix); btg_bloom_free(b); return nullptr; }
    return b;
}

int btg_bloom_save(const btg_bloom *b, const char *prefix) {
    BTG_REQUIRE_INIT();
    if (!b || !prefix) { set_error("null argument"); return BTG_EINVAL; }
    std::string p(prefix);
    std::ofstream meta(p + ".bloomMeta");
    if (!meta.is_open()) { set_error("Unable to write file %s.bloomMeta", prefix); return BTG_EIO; }
    meta << std::to_string(b->num_kmers) << "\t" << std::to_string(b->num_bits) << "\t" << std::to_string(K) << std::endl;
    meta.close();
    std::vector<uint8_t> host(b->nbytes());
    BTG_CUDA(cudaMemcpy(host.data(), b->bits, host.size(), cudaMemcpyDeviceToHost));
    std::ofstream data(p + ".bloomData", std::ios::out | std::ios::binary);
    if (!data.is_open()) { set_error("Unable to write file %s.bloomData", prefix); return BTG_EIO; }
    data.write((const char *)host.data(), host.size());
    return data.good() ? BTG_OK : BTG_EIO;
}

int btg_bloom_info(const btg_bloom *b, uint64_t *num_kmers, uint64_t *num_bits, uint32_t *num_hashes) {
    if (!b) { set_error("null bloom"); return BTG_EINVAL; }
    if (num_kmers) *num_kmers = b->num_kmers;
    if (num_bits) *num_bits = b->num_bits;
    if (num_hashes) *num_hashes = b->num_hashes;
    return BTG_OK;
}

int btg_bloom_download(const btg_bloom *b, uint8_t *out, uint64_t nbytes) {
    BTG_REQUIRE_INIT();
    if (!b || !out || nbytes < b->nbytes()) { set_error("bad download buffer"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy(out, b->bits, b->nbytes(), cudaMemcpyDeviceToHost));
    return BTG_OK;
}

int btg_bloom_lookup_probes_dev(const btg_bloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, uint8_t *probes_dev, void *stream) {
    BTG_REQUIRE_INIT();
    if (!b || (n && (!kmers_dev || !hit_dev))) { set_error("null argument"); return BTG_EINVAL; }
    if (n == 0) return BTG_OK;
    k_bloom_lookup<<<btg_grid_for(n, kBlock, 8), kBlock, 0, pick_stream(stream)>>>(b->view(), (const ulonglong2 *)kmers_dev, n, hit_dev, probes_dev);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_bloom_lookup_dev(const btg_bloom *b, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, void *stream) {
    return btg_bloom_lookup_probes_dev(b, kmers_dev, n, hit_dev, nullptr, stream);
}

int btg_bloom_insert_dev(btg_bloom *b, const uint64_t *kmers_dev, size_t n, void *stream) {
    BTG_REQUIRE_INIT();
    if (!b || (n && !kmers_dev)) { set_error("null argument"); return BTG_EINVAL; }
    if (n == 0) return BTG_OK;
    k_bloom_insert<<<btg_grid_for(n, kBlock, 8), kBlock, 0, pick_stream(stream)>>>(b->view(), b->bits, (const ulonglong2 *)kmers_dev, n);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_bloom_lookup(const btg_bloom *b, const uint64_t *kmers, size_t n, uint8_t *hit) {
    BTG_REQUIRE_INIT();
    if (!b || (n && (!kmers || !hit))) { set_error("null argument"); return BTG_EINVAL; }
    BloomView v = b->view();
    return staged(kmers, n, 1, hit, [&](ulonglong2 *d_in, size_t m, uint8_t *d_out, cudaStream_t s) {
        k_bloom_lookup<<<btg_grid_for(m, kBlock, 8), kBlock, 0, s>>>(v, d_in, m, d_out, nullptr);
        BTG_LAUNCHED();
    });
}

int btg_bloom_insert(btg_bloom *b, const uint64_t *kmers, size_t n) {
    BTG_REQUIRE_INIT();
    if (!b || (n && !kmers)) { set_error("null argument"); return BTG_EINVAL; }
    BloomView v = b->view();
    uint8_t *bits = b->bits;
    return staged(kmers, n, 0, nullptr, [&](ulonglong2 *d_in, size_t m, uint8_t *, cudaStream_t s) {
        k_bloom_insert<<<btg_grid_for(m, kBlock, 8), kBlock, 0, s>>>(v, bits, d_in, m);
        BTG_LAUNCHED();
    });
}

void btg_bloom_free(btg_bloom *b) {
    if (!b) return;
    cudaFree(b->bits);
    delete b;
}

// ---- ThreadedKmerBloom ---------------------------------------------------------
static TBloomView tview(const btg_tbloom *t) { return TBloomView{t->bits, t->sub_bits, t->sub_stride, ~0ULL / t->sub_bits, t->num_hashes}; }

btg_tbloom *btg_tbloom_create(uint64_t num_kmers, float fpr, int k) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (check_k(k)) return nullptr;
    if (!(fpr > 0.f && fpr < 1.f)) { set_error("fpr must be in (0,1)"); return nullptr; }
    auto *t = new btg_tbloom();
    // KmerBloom.cpp:213: new KmerBloom(std::ceil(num_kmers / static_cast<float>(root_size)), fpr)
    uint64_t sub = static_cast<uint64_t>(std::ceil(num_kmers / static_cast<float>(kThreadedRoots)));
    t->sub_kmers = sub ? sub : 1;
    t->sub_bits = opt_num_bits(fpr, t->sub_kmers);
    t->num_hashes = opt_num_hashes(t->sub_bits, t->sub_kmers);
    t->sub_stride = ((t->sub_bits + 7) / 8 + 3) & ~uint64_t(3);  // 4-byte aligned for the word atomics
    size_t total = (size_t)t->sub_stride * kThreadedRoots;
    if (cudaMalloc(&t->bits, total) != cudaSuccess) {
        set_error("cudaMalloc(%zu) for threaded Bloom filter failed", total);
        delete t;
        return nullptr;
    }
    cudaMemsetAsync(t->bits, 0, total, ctx().stream);
    cudaStreamSynchronize(ctx().stream);
    return t;
}

int btg_tbloom_info(const btg_tbloom *t, uint64_t *sub_kmers, uint64_t *sub_bits, uint32_t *num_hashes) {
    if (!t) { set_error("null bloom"); return BTG_EINVAL; }
    if (sub_kmers) *sub_kmers = t->sub_kmers;
    if (sub_bits) *sub_bits = t->sub_bits;
    if (num_hashes) *num_hashes = t->num_hashes;
    return BTG_OK;
}

int btg_tbloom_insert_dev(btg_tbloom *t, const uint64_t *kmers_dev, size_t n, void *stream) {
    BTG_REQUIRE_INIT();
    if (!t || (n && !kmers_dev)) { set_error("null argument"); return BTG_EINVAL; }
    if (n == 0) return BTG_OK;
    k_tbloom_insert<<<btg_grid_for(n, kBlock, 8), kBlock, 0, pick_stream(stream)>>>(tview(t), (const ulonglong2 *)kmers_dev, n);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_tbloom_lookup_dev(const btg_tbloom *t, const uint64_t *kmers_dev, size_t n, uint8_t *hit_dev, void *stream) {
    BTG_REQUIRE_INIT();
    if (!t || (n && (!kmers_dev || !hit_dev))) { set_error("null argument"); return BTG_EINVAL; }
    if (n == 0) return BTG_OK;
    k_tbloom_lookup<<<btg_grid_for(n, kBlock, 8), kBlock, 0, pick_stream(stream)>>>(tview(t), (const ulonglong2 *)kmers_dev, n, hit_dev);
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_tbloom_insert(btg_tbloom *t, const uint64_t *kmers, size_t n) {
    BTG_REQUIRE_INIT();
    if (!t || (n && !kmers)) { set_error("null argument"); return BTG_EINVAL; }
    TBloomView v = tview(t);
    return staged(kmers, n, 0, nullptr, [&](ulonglong2 *d_in, size_t m, uint8_t *, cudaStream_t s) {
        k_tbloom_insert<<<btg_grid_for(m, kBlock, 8), kBlock, 0, s>>>(v, d_in, m);
        BTG_LAUNCHED();
    });
}

int btg_tbloom_lookup(const btg_tbloom *t, const uint64_t *kmers, size_t n, uint8_t *hit) {
    BTG_REQUIRE_INIT();
    if (!t || (n && (!kmers || !hit))) { set_error("null argument"); return BTG_EINVAL; }
    TBloomView v = tview(t);
    return staged(kmers, n, 1, hit, [&](ulonglong2 *d_in, size_t m, uint8_t *d_out, cudaStream_t s) {
        k_tbloom_lookup<<<btg_grid_for(m, kBlock, 8), kBlock, 0, s>>>(v, d_in, m, d_out);
        BTG_LAUNCHED();
    });
}

int btg_tbloom_download(const btg_tbloom *t, uint8_t *out, uint64_t nbytes) {
    BTG_REQUIRE_INIT();
    const uint64_t sub_bytes = t ? (t->sub_bits + 7) / 8 : 0;
    if (!t || !out || nbytes < sub_bytes * kThreadedRoots) { set_error("bad download buffer"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy2D(out, sub_bytes, t->bits, t->sub_stride, sub_bytes, kThreadedRoots, cudaMemcpyDeviceToHost));
    return BTG_OK;
}

void btg_tbloom_free(btg_tbloom *t) {
    if (!t) return;
    cudaFree(t->bits);
    delete t;
}

// ---- k-mer primitives ------------------------------------------------------------
int btg_kmer_hash(const uint64_t *kmers, size_t n, uint64_t *hash_out) {
    BTG_REQUIRE_INIT();
    if (n && (!kmers || !hash_out)) { set_error("null argument"); return BTG_EINVAL; }
    return staged(kmers, n, 8, hash_out, [&](ulonglong2 *d_in, size_t m, uint8_t *d_out, cudaStream_t s) {
        k_kmer_hash<<<btg_grid_for(m, kBlock, 8), kBlock, 0, s>>>(d_in, m, (uint64_t *)d_out);
        BTG_LAUNCHED();
    });
}

int btg_kmer_canonical(const uint64_t *kmers, size_t n, uint64_t *canon_out) {
    BTG_REQUIRE_INIT();
    if (n && (!kmers || !canon_out)) { set_error("null argument"); return BTG_EINVAL; }
    return staged(kmers, n, 16, canon_out, [&](ulonglong2 *d_in, size_t m, uint8_t *d_out, cudaStream_t s) {
        k_kmer_canonical<<<btg_grid_for(m, kBlock, 8), kBlock, 0, s>>>(d_in, m, (ulonglong2 *)d_out);
        BTG_LAUNCHED();
    });
}

int btg_scan_sequence_dev(const char *seq_dev, size_t len, uint64_t *kmers_out_dev, uint8_t *valid_out_dev, void *stream) {
    BTG_REQUIRE_INIT();
    if (len && (!seq_dev || !kmers_out_dev || !valid_out_dev)) { set_error("null argument"); return BTG_EINVAL; }
    if (len == 0) return BTG_OK;
    size_t nchunks = (len + kScanChunk - 1) / kScanChunk;
    k_scan<0><<<btg_grid_for(nchunks, kBlock, 4), kBlock, 0, pick_stream(stream)>>>(seq_dev, len, (ulonglong2 *)kmers_out_dev, valid_out_dev, BloomView{});
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_scan_sequence_lookup_dev(const btg_bloom *b, const char *seq_dev, size_t len, uint8_t *hit_out_dev, void *stream) {
    BTG_REQUIRE_INIT();
    if (!b || (len && (!seq_dev || !hit_out_dev))) { set_error("null argument"); return BTG_EINVAL; }
    if (len == 0) return BTG_OK;
    size_t nchunks = (len + kScanChunk - 1) / kScanChunk;
    k_scan<1><<<btg_grid_for(nchunks, kBlock, 4), kBlock, 0, pick_stream(stream)>>>(seq_dev, len, nullptr, hit_out_dev, b->view());
    BTG_LAUNCHED();
    BTG_CUDA(cudaGetLastError());
    return BTG_OK;
}

int btg_scan_sequence(const char *seq, size_t len, uint64_t *kmers_out, uint8_t *valid_out) {
    BTG_REQUIRE_INIT();
    if (len && (!seq || !kmers_out || !valid_out)) { set_error("null argument"); return BTG_EINVAL; }
    if (len == 0) return BTG_OK;
    char *d_seq = nullptr;
    uint64_t *d_k = nullptr;
    uint8_t *d_v = nullptr;
    int rc = BTG_OK;
    auto &c = ctx();
    if (cudaMalloc(&d_seq, len) != cudaSuccess || cudaMalloc(&d_k, len * 16) != cudaSuccess || cudaMalloc(&d_v, len) != cudaSuccess) {
        set_error("scan allocation failed");
        rc = BTG_ENOMEM;
    } else {
        cudaMemcpyAsync(d_seq, seq, len, cudaMemcpyHostToDevice, c.stream);
        rc = btg_scan_sequence_dev(d_seq, len, d_k, d_v, c.stream);
        if (rc == BTG_OK) {
            cudaMemcpyAsync(kmers_out, d_k, len * 16, cudaMemcpyDeviceToHost, c.stream);
            cudaMemcpyAsync(valid_out, d_v, len, cudaMemcpyDeviceToHost, c.stream);
            cudaError_t e = cudaStreamSynchronize(c.stream);
            if (e != cudaSuccess) { set_error("scan failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; }
        }
    }
    cudaFree(d_seq);
    cudaFree(d_k);
    cudaFree(d_v);
    return rc;
}

}  // extern "C"
