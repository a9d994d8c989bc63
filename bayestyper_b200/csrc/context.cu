// context.cu — library lifecycle (btg_init / btg_shutdown / errors).
#include <cstdlib>
#include <deque>
#include <map>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace btg {

static Context g_ctx;
Context &ctx() { return g_ctx; }
std::atomic<uint64_t> g_launches{0};

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- block cache behind dmalloc / dfree ------------------------------------------------------------------------------------
// Freed blocks are kept (per device) and handed out again to a request of the same size class; units of one run repeat their
// sizes, so after the first unit every allocation is a hit.  The cache holds at most a quarter of the device memory (oldest blocks
// go first) and is emptied when a cudaMalloc fails.
namespace {
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;          // size -> block
    std::unordered_map<void *, size_t> size_of;          // every live or cached block -> its size
    std::deque<void *> age;                              // cached blocks, oldest first
    size_t cached = 0, limit = 0;
    int device = -1;

    static size_t round_up(size_t b) {
        if (b < 4096) return (b + 255) & ~size_t(255);
        const size_t g = b < (1u << 20) ? 4096 : (size_t)1 << 20;   // 4 KB granules below 1 MB, 1 MB above
        return (b + g - 1) / g * g;
    }
    void drop_locked(void *p) {
        auto it = size_of.find(p);
        if (it == size_of.end()) return;
        auto range = free_blocks.equal_range(it->second);
        for (auto f = range.first; f != range.second; ++f)
            if (f->second == p) { free_blocks.erase(f); break; }
        cached -= it->second;
        size_of.erase(it);
        cudaFree(p);
    }
    void trim_locked(size_t keep) {
        while (cached > keep && !age.empty()) {
            void *p = age.front();
            age.pop_front();
            if (size_of.count(p)) {
                bool is_free = false;
                auto range = free_blocks.equal_range(size_of[p]);
                for (auto f = range.first; f != range.second; ++f) if (f->second == p) { is_free = true; break; }
                if (is_free) drop_locked(p);
            }
        }
    }
};
BlockCache g_cache;
}  // namespace

cudaError_t dmalloc_bytes(void **p, size_t bytes) {
    *p = nullptr;
    static const bool off = getenv("BTG_NO_POOL") != nullptr;
    if (off) return cudaMalloc(p, bytes ? bytes : 1);
    const size_t want = BlockCache::round_up(bytes ? bytes : 1);
    std::lock_guard<std::mutex> lock(g_cache.mu);
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_cache.device != dev) {            // one device per process (btg_init); a change of device empties the cache
        g_cache.trim_locked(0);
        g_cache.device = dev;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        g_cache.limit = total_b / 4;
    }
    auto it = g_cache.free_blocks.find(want);
    if (it != g_cache.free_blocks.end()) {
        *p = it->second;
        g_cache.free_blocks.erase(it);
        g_cache.cached -= want;
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {                 // give the cached blocks back to the device and try once more
        cudaGetLastError();
        cudaDeviceSynchronize();
        g_cache.trim_locked(0);
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) { *p = nullptr; return e; }
    g_cache.size_of[*p] = want;
    return cudaSuccess;
}

void dfree(void *p) {
    if (!p) return;
    static const bool off = getenv("BTG_NO_POOL") != nullptr;
    if (off) { cudaFree(p); return; }
    cudaDeviceSynchronize();                // cudaFree's contract: no stream still uses the block when it is handed out again
    std::lock_guard<std::mutex> lock(g_cache.mu);
    auto it = g_cache.size_of.find(p);
    if (it == g_cache.size_of.end()) { cudaFreeAsync(p, ctx().stream); return; }     // not ours: a stream-ordered allocation handed over by btg_counter goes back to its pool
    g_cache.free_blocks.emplace(it->second, p);
    g_cache.age.push_back(p);
    g_cache.cached += it->second;
    if (g_cache.cached > g_cache.limit) g_cache.trim_locked(g_cache.limit);
}

void release_cached_blocks() {
    std::lock_guard<std::mutex> lock(g_cache.mu);
    g_cache.trim_locked(0);
}

size_t free_device_memory() {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    std::lock_guard<std::mutex> lock(g_cache.mu);
    return free_b + g_cache.cached;
}

}  // namespace btg

extern "C" {

int btg_init(int device) {
    auto &c = btg::ctx();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        btg::set_error("no CUDA device available (%s); libbtgpu has no CPU fallback", cudaGetErrorString(e));
        return BTG_ECUDA;
    }
    if (device < 0 || device >= n) {
        btg::set_error("device %d out of range (have %d)", device, n);
        return BTG_EINVAL;
    }
    BTG_CUDA(cudaSetDevice(device));
    if (c.ready && c.device == device) return BTG_OK;
    if (c.ready) btg_shutdown();
    cudaDeviceProp prop;
    BTG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        btg::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return BTG_ECUDA;
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    if (const char *g = getenv("BTG_L2_FETCH")) {  // experiment knob: DRAM fetch granularity of L2 misses (32/64/128 B)
        size_t before = 0, after = 0;
        cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "[btgpu] cudaLimitMaxL2FetchGranularity %zu -> %zu (%s)\n", before, after, cudaGetErrorString(e));
    }
    BTG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    BTG_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.ready = true;
    btg::g_launches = 0;
    return BTG_OK;
}

void btg_shutdown(void) {
    auto &c = btg::ctx();
    if (!c.ready) return;
    cudaStreamSynchronize(c.stream);
    btg::release_cached_blocks();
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copy_stream);
    c.stream = c.copy_stream = nullptr;
    c.ready = false;
}

const char *btg_last_error(void) { return btg::g_err; }
int btg_version(void) { return 100; }
int btg_device_sm_count(void) { return btg::ctx().ready ? btg::ctx().sm_count : 0; }

void *btg_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        btg::set_error("cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void btg_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// device buffers for hosts that do not link the CUDA runtime themselves (host/btpipeline.cpp): sample k-mer records and region buffers that the
// *_dev entry points read
void *btg_device_alloc(size_t bytes) {
    if (!btg::ctx().ready) { btg::set_error("btg_init() has not been called"); return nullptr; }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { btg::set_error("cudaMalloc(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
    return p;
}
void btg_device_free(void *p) {
    if (p) { cudaStreamSynchronize(btg::ctx().stream); cudaFree(p); }
}
int btg_copy_to_device(void *dst_dev, const void *src_host, size_t bytes) {
    BTG_REQUIRE_INIT();
    if (bytes == 0) return BTG_OK;
    BTG_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, btg::ctx().stream));
    BTG_CUDA(cudaStreamSynchronize(btg::ctx().stream));
    return BTG_OK;
}
int btg_copy_to_host(void *dst_host, const void *src_dev, size_t bytes) {
    BTG_REQUIRE_INIT();
    if (bytes == 0) return BTG_OK;
    BTG_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, btg::ctx().stream));
    BTG_CUDA(cudaStreamSynchronize(btg::ctx().stream));
    return BTG_OK;
}

void *btg_get_stream(void) { return btg::ctx().ready ? (void *)btg::ctx().stream : nullptr; }
uint64_t btg_launch_count(void) { return btg::g_launches.load(); }
void btg_launch_count_reset(void) { btg::g_launches = 0; }

}  // extern "C"
