// common.cuh — library context, error plumbing and launch accounting for libbtgpu.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

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

inline cudaStream_t pick_stream(void *s) { return s ? (cudaStream_t)s : ctx().stream; }

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
