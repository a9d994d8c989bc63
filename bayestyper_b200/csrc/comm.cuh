// comm.cuh — the one exchange step of the sharded lock-step modes, done INSIDE the persistent chain
// kernel over peer memory: every rank's per-sample (n_obs, sum) of the CountAllocation histogram is written straight
// into every peer's mailbox (NVLink stores through CUDA-IPC peer mappings), followed by a release flag; the rank then
// waits for the flags of all peers and adds the rows up in rank order.  Integer sums, so every rank obtains the same
// totals and — sharing CountDistribution's random stream — draws the same noise rates: the equivalent of the
// reference's per-iteration merge of all threads' CountAllocations (InferenceEngine.cpp:226-229,445-448) without a
// separate collective launch.
#pragma once
#include <cstdint>

namespace btg {

constexpr uint32_t kMaxRanks = 8;
constexpr uint32_t kMailRow = 64;  // u64 per (slot, source rank): 2*S <= 60 values + padding

// mailbox of one rank (lives in that rank's HBM, written by all ranks)
struct Mailbox {
    unsigned long long data[2][kMaxRanks][kMailRow];
    unsigned long long flag[2][kMaxRanks];  // sequence number of the iteration the row belongs to
};

struct PeerExchange {
    uint32_t world, rank;
    Mailbox *mail[kMaxRanks];       // mail[r] = rank r's mailbox as mapped in this process (mail[rank] is local)
    unsigned long long seq0;        // sequence number of this launch's iteration 0 (monotonic over the communicator's life)
    unsigned long long timeout_ns;  // give up waiting for a peer after this long ...
    uint32_t *error;                // ... and raise this flag (the host turns it into an error return)
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Sum-all-reduce of vals[0..n) (n <= kMailRow) across the ranks, in place, by ONE block (all of its threads call this;
// blockDim.x >= world).  `seq` must advance by one per call on every rank.  Thread t < world serves peer t: it
// pushes this rank's row into peer t's mailbox, then waits for peer t's row in the local mailbox.
__device__ inline void peer_allreduce_block(const PeerExchange &px, unsigned long long seq, unsigned long long *vals, uint32_t n) {
    if (px.world <= 1) return;
    const uint32_t t = threadIdx.x, slot = (uint32_t)(seq & 1ull);
    __syncthreads();  // vals complete
    if (t < px.world) {
        Mailbox *peer = px.mail[t];
        for (uint32_t i = 0; i < n; i++) st_relaxed_sys(&peer->data[slot][px.rank][i], vals[i]);
        st_release_sys(&peer->flag[slot][px.rank], seq);  // release: the row is visible before the flag
        const Mailbox *mine = px.mail[px.rank];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(&mine->flag[slot][t]) != seq) {
            if (global_timer_ns() - t0 > px.timeout_ns) { atomicExch(px.error, 1u + t); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (t < n) {
        const Mailbox *mine = px.mail[px.rank];
        unsigned long long acc = 0;
        for (uint32_t r = 0; r < px.world; r++) acc += ld_relaxed_sys(&mine->data[slot][r][t]);
        vals[t] = acc;
    }
    __syncthreads();
}
#endif

}  // namespace btg

// host handle (comm.cu)
struct btg_comm {
    uint32_t world = 1, rank = 0;
    btg::Mailbox *local = nullptr;
    btg::Mailbox *peers[btg::kMaxRanks] = {};
    bool opened[btg::kMaxRanks] = {};
    unsigned long long seq = 1;      // next unused sequence number (0 = "never written")
    uint32_t *error = nullptr;       // device flag
    bool connected = false;
};
