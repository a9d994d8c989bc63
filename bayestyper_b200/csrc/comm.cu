// comm.cu — peer mailboxes for the sharded lock-step modes: one rank per process and GPU; every rank allocates a
// mailbox in its own HBM, exports it as a CUDA IPC handle (64 bytes the host exchanges by whatever means it has —
// torch.distributed, MPI, a file), and maps the mailboxes of its peers.  The exchange itself happens inside the
// chain kernel (comm.cuh).
#include "common.cuh"
#include "comm.cuh"

using namespace btg;

static_assert(sizeof(cudaIpcMemHandle_t) == BTG_COMM_HANDLE_BYTES, "handle size");

extern "C" {

btg_comm *btg_comm_create(uint32_t world, uint32_t rank, uint8_t *handle_out) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (world == 0 || world > kMaxRanks || rank >= world || !handle_out) { set_error("bad communicator arguments (world 1..%u)", kMaxRanks); return nullptr; }
    auto *c = new btg_comm();
    c->world = world; c->rank = rank;
    if (cudaMalloc(&c->local, sizeof(Mailbox)) != cudaSuccess || cudaMalloc(&c->error, 4) != cudaSuccess) {
        set_error("mailbox allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        btg_comm_free(c);
        return nullptr;
    }
    cudaMemset(c->local, 0, sizeof(Mailbox));
    cudaMemset(c->error, 0, 4);
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof h);
    if (world > 1 && cudaIpcGetMemHandle(&h, c->local) != cudaSuccess) {
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
        btg_comm_free(c);
        return nullptr;
    }
    memcpy(handle_out, &h, sizeof h);
    c->peers[rank] = c->local;
    c->connected = world == 1;
    return c;
}

int btg_comm_connect(btg_comm *c, const uint8_t *all_handles) {
    BTG_REQUIRE_INIT();
    if (!c || !all_handles) { set_error("null argument"); return BTG_EINVAL; }
    for (uint32_t r = 0; r < c->world; r++) {
        if (r == c->rank || c->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)r * sizeof h, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %u) failed: %s", r, cudaGetErrorString(e)); cudaGetLastError(); return BTG_ECUDA; }
        c->peers[r] = (Mailbox *)p;
        c->opened[r] = true;
    }
    c->connected = true;
    return BTG_OK;
}

void btg_comm_free(btg_comm *c) {
    if (!c) return;
    for (uint32_t r = 0; r < kMaxRanks; r++)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peers[r]);
    cudaFree(c->local);
    cudaFree(c->error);
    delete c;
}

}  // extern "C"
