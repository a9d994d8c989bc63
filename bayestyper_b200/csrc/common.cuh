// common.cuh — library context, error plumbing and launch accounting for libbtgpu.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: a no-op unless a tool (nsys, ncu --nvtx) injects itself

#include "../../include/btgpu.h"

namespace btg {

struct Context {
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;   // library stream (host entry points)
    cudaStream_t copy_stream = nullptr;
    bool ready = false;
};

Context &ctx();
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

// NVTX range over one entry point of the C ABI (the reference has no tracing; a timeline of the stage calls —
// btg_find_sample_paths, btg_counter_*, btg_estimate_* — is what a profiler needs to attribute kernels to stages)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

inline cudaStream_t pick_stream(void *s) { return s ? (cudaStream_t)s : ctx().stream; }

// Device allocations of the per-unit objects (unit arrays, sampler arena, results, path-search scratch) go through a block cache
// (context.cu): a multi-GB cudaMalloc / cudaFree pair maps and unmaps physical memory in the driver (measured on configs[1]: 20-500 ms
// per freed unit, different every time), the cache hands the same blocks to the next unit.  dfree waits for the device like cudaFree
// does.  (The stream-ordered pool of the CUDA runtime was tried for these blocks too: sharing it with the ~100 small stream-ordered
// allocations of counter.cu made multi-GB requests remap physical memory, 0.4 s stalls at random stages.)
cudaError_t dmalloc_bytes(void **p, size_t bytes);
void dfree(void *p);
void release_cached_blocks();   // btg_shutdown
size_t free_device_memory();   // cudaMemGetInfo's free bytes + what the pool holds without using it
template <class T> inline cudaError_t dmalloc(T **p, size_t bytes) { return dmalloc_bytes(reinterpret_cast<void **>(p), bytes); }

}  // namespace btg

#define BTG_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            btg::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return BTG_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

#define BTG_CUDA_NULL(call)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            btg::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return nullptr;                                                                         \
        }                                                                                           \
    } while (0)

#define BTG_REQUIRE_INIT()                                                   \
    btg::NvtxRange _btg_nvtx_range(__func__);                                \
    do {                                                                     \
        if (!btg::ctx().ready) {                                             \
            btg::set_error("btg_init() has not been called");                \
            return BTG_ESTATE;                                               \
        }                                                                    \
    } while (0)

#define BTG_LAUNCHED() (btg::g_launches.fetch_add(1, std::memory_order_relaxed))

// grid sizing: a multiple of the SM count (148 on B200), capped by the work
inline unsigned btg_grid_for(size_t work_items, unsigned block, unsigned ctas_per_sm) {
    size_t need = (work_items + block - 1) / block;
    size_t cap = (size_t)btg::ctx().sm_count * ctas_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}
