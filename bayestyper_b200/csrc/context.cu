// context.cu — library lifecycle (btg_init / btg_shutdown / errors).
#include <cstdlib>

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
