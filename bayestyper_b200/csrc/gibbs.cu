// gibbs.cu — the per-cluster Gibbs sampler on the device (sm_100a).
//
// Replaces, for one inference unit resident in HBM:
//   InferenceEngine::estimateGenotypes / estimateNoise        src/bayesTyper/InferenceEngine.cpp:135-382
//   VariantClusterGroup::estimateGenotypes / runGibbsSample   src/bayesTyper/VariantClusterGroup.cpp:220-250
//   VariantClusterGenotyper (ctor, reset, sampleDiplotypes, sampleDiplotype, calcDiplotypeLogProb,
//     sampleHaplotypeFrequencies, getNoiseCounts, getGenotypes…)  src/bayesTyper/VariantClusterGenotyper.cpp
//   VariantClusterHaplotypes (sampleKmerSubset, updateAlleleKmerStats…) src/bayesTyper/VariantClusterHaplotypes.cpp
//   CountDistribution / NegativeBinomialDistribution tables   src/bayesTyper/CountDistribution.cpp
//   (Sparse)FrequencyDistribution, SparsityEstimator, DiscreteSampler, KmerStats, CountAllocation
//
// Mapping.  Variant-cluster groups are independent and, inside a group, the sampler is
// a strictly sequential chain (sample s+1 sees the haplotype counts left by sample s; iteration i+1 sees
// the frequencies drawn in iteration i).  The unit of parallelism is therefore the GROUP: one thread
// walks one group through all chains x iterations with its state in a private arena slice; groups are
// sorted by cost so that the 32 lanes of a warp carry similar work.  All arithmetic is f64 (the
// reference's), table lookups go to the L2-resident log-pmf cache ([S][256][256] doubles).
//
// Groups made of ONE cluster (the bulk of any unit) run one thread per cluster in k_estimate_genotypes.  Groups
// with nested clusters (VariantClusterGroup::runGibbsSample recursion, multicluster k-mers sharing a multiplicity
// record) run one thread per GROUP in k_estimate_genotypes_nested, which walks the group's clusters in the
// reference's depth-first order every iteration.  In the joint noise mode such units take the warp-per-group kernel of gibbs_wide.cu.
#include "gibbs_core.cuh"

// clusters whose cache fill costs more than this many table lookups PER SAMPLE are worked on by a warp in the lock-step chains (dense
// k-mer tile, grid-wide fill in the first iteration); the others run one per thread.  BTG_NOISE_BIG overrides (sweeps: profiles/).
static uint32_t big_fill_cost() {
    static const uint32_t v = getenv("BTG_NOISE_BIG") ? (uint32_t)strtoul(getenv("BTG_NOISE_BIG"), nullptr, 10) : kBigFillCost;
    return v;
}

namespace {

// ---------------------------------------------------------------------------------------------
// CountDistribution tables
// ---------------------------------------------------------------------------------------------
// CountDistribution::updateGenomicCache / genomicCountLogPmf (CountDistribution.cpp:215-238,267-312): one thread per (s, m, c)
__global__ void k_genomic_table(const double *__restrict__ p, const double *__restrict__ size, uint32_t S, double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * 65536u) return;
    const uint32_t s = i >> 16, m = (i >> 8) & 255u, c = i & 255u;
    double v;
    if (m == 0) {
        v = c == 0 ? 0.0 : -INFINITY;
    } else {
        v = nbLogPmf(p[s], size[s], c, m);
        if (c == 255) {
            uint32_t limit = c;
            double prev;
            do {
                limit++;
                prev = v;
                v = logAddition(v, nbLogPmf(p[s], size[s], limit, m));
                if (v > 0) { v = 0; break; }
            } while (!doubleCompare(prev, v));
        }
    }
    out[i] = v;
}

__global__ void k_noise_table(const double *__restrict__ rates, uint32_t S, double *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * 256u) return;
    out[i] = noiseCountLogPmf(rates[i >> 8], i & 255u);
}

__global__ void k_lgamma_int(double *out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lgamma((double)i);  // out[0] = inf, never read
}

// InferenceEngine::estimateGenotypesCallback (InferenceEngine.cpp:278-333): one thread = one group, all chains
// Chains of one cluster are independent in this mode: every chain re-keys the cluster's two random streams with its chain
// index, shuffles the original k-mer order and starts from reset frequencies; tallies and allele statistics are sums over
// chains.  (The reference keeps ONE mt19937 running through all chains of a genotyper, InferenceEngine.cpp:292-306 — a
// property of its generator, not of the model; oracle-P follows the per-chain contract.)  That lets the few large clusters,
// which otherwise set the kernel's tail (one thread: 1.3 s while the mean thread takes 0.16 s, profiles/r1_gibbs_tail.txt), run
// their chains on kChainSplit threads with private arena positions; k_merge_split adds the pieces up and summarises.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(64, MIN_BLOCKS) k_estimate_genotypes(DevUnit du, Tables T, btg_gibbs_opts o, ResultView R, int reconverge, unsigned long long *dbg, int hot) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long t_in = dbg ? global_timer_ns() : 0;
    // no early return: every lane of the warp reaches the __syncwarp()s below
    const uint32_t n_virtual = du.n_split * kChainSplit;
    bool live;
    uint32_t cluster = 0, pos = NONE32, chain0 = 0, chain_step = 1;
    bool split = false;
    if (t < n_virtual) {  // the large clusters first: their blocks start before everything else
        const uint32_t j = t / kChainSplit, v = t % kChainSplit;
        cluster = du.split_cluster[j];
        pos = du.split_pos[(size_t)j * kChainSplit + v];
        chain0 = v; chain_step = kChainSplit;
        split = true; live = true;
    } else {
        const uint32_t i = t - n_virtual;
        live = i < du.n_regular;
        cluster = du.order[live ? i : 0];
        if (live && du.split_of[cluster] != NONE32) live = false;  // handled above
    }
    extern __shared__ __align__(16) uint8_t hot_smem[];
    Cl cl;
    cl.bind(du, cluster, pos);
    if (hot && live && !split && cl.H <= kHotH) cl.bind_hot(hot_smem, threadIdx.x, blockDim.x);  // small cluster: hot state in shared memory
    const uint64_t gidx = group_index(o, cl.g);
    const uint8_t *ploidy = du.group_ploidy + (size_t)cl.g * du.S;
    Philox prng, fr;
    prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, 0);
    fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, 0);
    if (live) cl_construct(cl, o, gidx, 0);
    const uint32_t iters = (uint32_t)o.gibbs_burn_in + o.gibbs_samples;
    // warp-uniform trip count (lanes of split clusters take fewer chains; they idle through the rest)
    for (uint32_t round = 0; round < o.n_chains; round++) {
        const uint32_t chain = chain0 + round * chain_step;
        const bool run = live && chain < o.n_chains;
        if (run) {
            prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, chain);
            fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, chain);
            cl_reset<false, false, true>(cl, o, prng);
        }
        if (reconverge) __syncwarp();
        if (__all_sync(0xFFFFFFFFu, !run)) continue;  // nothing left for this warp in this round
        for (uint32_t it = 0; it < iters; it++) {
            if (run) cl_sample_diplotypes(cl, T, ploidy, it >= o.gibbs_burn_in, prng);
            if (reconverge) __syncwarp();
            if (run) cl_sample_frequencies(cl, fr);
            if (reconverge) __syncwarp();
        }
    }
    if (live && !split) cl_summarise(cl, o, ploidy, R);
    if (dbg && live) {  // BTG_GIBBS_TIMING=1: slowest thread and the sum of all per-thread times
        const unsigned long long dt = global_timer_ns() - t_in;
        atomicMax(dbg, (dt << 32) | cl.c);
        atomicAdd(dbg + 1, dt);
        if (cl.H > 4) atomicAdd(dbg + 2, dt);
    }
}

// chain-split clusters: add the tallies and allele statistics of the virtual threads 1.. into thread 0's arena position
// (ascending order, so the f64 sums are reproducible) and summarise
__global__ void __launch_bounds__(64) k_merge_split(DevUnit du, btg_gibbs_opts o, ResultView R) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= du.n_split) return;
    Cl cl, part;
    cl.bind(du, du.split_cluster[j]);
    for (uint32_t v = 1; v < kChainSplit; v++) {
        part.bind(du, du.split_cluster[j], du.split_pos[(size_t)j * kChainSplit + v]);
        for (uint32_t i = 0; i < cl.Dall * cl.S; i++) cl.tally[i] += part.tally[i];
        for (uint32_t i = 0; i < cl.n_alleles * cl.S * 3; i++) { cl.as_n[i] += part.as_n[i]; cl.as_f[i] += part.as_f[i]; }
    }
    cl_summarise(cl, o, du.group_ploidy + (size_t)cl.g * du.S, R);
}


// InferenceEngine::estimateGenotypesCallback for a group with several clusters: one thread = one GROUP.  Per chain the branch
// orderings are shuffled (cumulatively, VariantClusterGroup.cpp:208-218) and flattened into the depth-first order that
// runGibbsSample's recursion (…Group.cpp:236-250) visits; every iteration walks that order, each cluster passing the
// NestedVariantClusterInfo of its children on before they run.
__global__ void __launch_bounds__(64) k_estimate_genotypes_nested(DevUnit du, Tables T, btg_gibbs_opts o, ResultView R) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= du.n_nested_groups) return;
    const uint32_t g = du.nested_groups[i], S = du.S;
    const uint64_t c0 = du.group_cluster_off[g];
    const uint32_t n = (uint32_t)(du.group_cluster_off[g + 1] - c0);
    const uint64_t gidx = group_index(o, g);
    const uint8_t *ploidy = du.group_ploidy + (size_t)g * S;
    const uint64_t s0 = du.group_src_off[g], s1 = du.group_src_off[g + 1];
    const uint64_t e0 = du.cl_edge_off[c0], e1 = du.cl_edge_off[c0 + n];
    for (uint64_t e = s0; e < s1; e++) du.src_mut[e] = du.group_src[e];
    for (uint64_t e = e0; e < e1; e++) du.edge_mut[e] = du.edge_dst[e];
    Cl cl;
    for (uint32_t j = 0; j < n; j++) {  // VariantClusterGroup::initGenotyper: genotypers are constructed once
        cl.bind(du, (uint32_t)(c0 + j));
        cl_construct(cl, o, gidx, 0);
        Philox prng, fr;
        prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, 0);
        fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, 0);
        prng.save(cl.rng, kRng0);
        fr.save(cl.rng, kRng1);
    }
    const uint32_t iters = (uint32_t)o.gibbs_burn_in + o.gibbs_samples;
    for (uint32_t chain = 0; chain < o.n_chains; chain++) {
        for (uint32_t j = 0; j < n; j++) {
            cl.bind(du, (uint32_t)(c0 + j));
            Philox prng;
            prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
            cl_reset<true>(cl, o, prng);
            prng.save(cl.rng, kRng0);
        }
        {   // shuffleBranchOrdering: sources, then every vertex's out-edges in vertex order (stream kind 3)
            Philox br;
            br.init(o.random_seed, gidx, 0, kRngBranch, chain);
            for (uint64_t m = s1 - s0; m > 1; m--) {
                const uint32_t j = br.uniform_int((uint32_t)m);
                const uint32_t t = du.src_mut[s0 + m - 1]; du.src_mut[s0 + m - 1] = du.src_mut[s0 + j]; du.src_mut[s0 + j] = t;
            }
            for (uint32_t v = 0; v < n; v++) {
                const uint64_t b0 = du.cl_edge_off[c0 + v];
                for (uint64_t m = du.cl_edge_off[c0 + v + 1] - b0; m > 1; m--) {
                    const uint32_t j = br.uniform_int((uint32_t)m);
                    const uint32_t t = du.edge_mut[b0 + m - 1]; du.edge_mut[b0 + m - 1] = du.edge_mut[b0 + j]; du.edge_mut[b0 + j] = t;
                }
            }
        }
        {   // depth-first pre-order of the forest
            uint32_t top = 0, len = 0;
            for (uint64_t e = s1; e > s0; e--) du.dfs_stack[c0 + top++] = du.src_mut[e - 1];
            while (top > 0) {
                const uint32_t v = du.dfs_stack[c0 + --top];
                du.dfs_order[c0 + len++] = v;
                for (uint64_t e = du.cl_edge_off[c0 + v + 1]; e > du.cl_edge_off[c0 + v]; e--) du.dfs_stack[c0 + top++] = du.edge_mut[e - 1];
            }
        }
        for (uint64_t e = s0; e < s1; e++) {  // the info a source vertex receives: the chromosome ploidy, no enclosing allele
            const uint32_t ns = du.nest_slot[c0 + du.src_mut[e]];
            for (uint32_t s = 0; s < S; s++) { du.nest_pl[(size_t)ns * S + s] = ploidy[s]; du.nest_k[(size_t)ns * S + s] = 0; }
        }
        for (uint32_t it = 0; it < iters; it++) {
            const bool collect = it >= o.gibbs_burn_in;
            for (uint32_t pos = 0; pos < n; pos++) {
                const uint32_t v = du.dfs_order[c0 + pos];
                cl.bind(du, (uint32_t)(c0 + v));
                const uint32_t ns = du.nest_slot[c0 + v];
                Philox prng, fr;
                prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
                fr.load(cl.rng, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
                cl_sample_diplotypes<true>(cl, T, du.nest_pl + (size_t)ns * S, collect, prng);
                if (collect) cl_add_nested_stats(cl, ns);
                cl_sample_frequencies(cl, fr);
                prng.save(cl.rng, kRng0);
                fr.save(cl.rng, kRng1);
                for (uint64_t e = du.cl_edge_off[c0 + v]; e < du.cl_edge_off[c0 + v + 1]; e++) {
                    const uint32_t t = du.edge_mut[e], nt = du.nest_slot[c0 + t];
                    for (uint32_t s = 0; s < S; s++) {
                        du.nest_pl[(size_t)nt * S + s] = du.nest_pl[(size_t)ns * S + s];
                        const uint32_t nk = du.nest_k[(size_t)ns * S + s];
                        du.nest_k[(size_t)nt * S + s] = (uint8_t)nk;
                        for (uint32_t k = 0; k < nk; k++) {
                            const size_t a = ((size_t)ns * S + s) * 2 + k, b = ((size_t)nt * S + s) * 2 + k;
                            du.nest_n[b] = du.nest_n[a]; du.nest_f[2 * b] = du.nest_f[2 * a]; du.nest_f[2 * b + 1] = du.nest_f[2 * a + 1];
                        }
                    }
                    cl_update_nested_info(cl, nt, du.cluster_idx[c0 + t]);
                }
            }
        }
    }
    for (uint32_t j = 0; j < n; j++) {  // collectGenotypes: every cluster is summarised with the chromosome ploidy
        cl.bind(du, (uint32_t)(c0 + j));
        cl_summarise(cl, o, ploidy, R);
    }
}

// VariantClusterGroup::collectGenotypes for every cluster (joint mode collects after all chains)
__global__ void __launch_bounds__(64) k_summarise(DevUnit du, btg_gibbs_opts o, ResultView R) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= du.C) return;
    Cl cl;
    cl.bind(du, du.order[i]);
    cl_summarise(cl, o, du.group_ploidy + (size_t)cl.g * du.S, R);
}

__global__ void k_noise_update(NoiseState ns, uint32_t S, float prior_shape, float prior_scale, uint32_t seed, int mode, int accumulate,
                               double chain_label, double iter_label, double mean_div) {
    __shared__ double sh_rates[BTG_MAX_SAMPLES];
    noise_update_block(ns, S, prior_shape, prior_scale, seed, mode, accumulate, chain_label, iter_label, mean_div, sh_rates);
}

// One whole chain of estimateNoise as ONE persistent cooperative kernel (InferenceEngine.cpp:191-253): every thread keeps its
// clusters' state hot in L1 across the 350 iterations; the per-iteration "join + merge + sampleNoiseParameters" of the
// reference (thread spawn/join per iteration, InferenceEngine.cpp:213-226) becomes two grid-wide barriers around block 0's
// histogram -> Gamma draw -> Poisson-row rebuild.
// joint = 0: estimateNoise (fresh genotypers each chain, streams of chain `chain`, nothing collected)
// joint = 1: estimateNoiseAndGenotypes (InferenceEngine.cpp:384-472): genotypers are constructed in the first chain only and
//            persist (streams of chain 0), samples are collected after the burn-in
#ifndef BTG_NOISE_MINBLOCKS
#define BTG_NOISE_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(256, BTG_NOISE_MINBLOCKS) k_noise_chain(DevUnit du, Tables T, btg_gibbs_opts o, const uint32_t *sel, uint32_t n_sel, uint32_t n_big, uint32_t chain,
                                                       uint32_t iters, NoiseState ns, float prior_shape, float prior_scale, unsigned long long *hist, int joint,
                                                       PeerExchange px, const uint32_t *fill_tasks, uint32_t n_fill_tasks, GridBarrier gb, int hot) {
    extern __shared__ __align__(16) uint8_t hot_smem[];
    __shared__ unsigned long long sh_tot[kMailRow];
    __shared__ double sh_rates[BTG_MAX_SAMPLES];
    // getNoiseCounts of the block's clusters: only (n_obs, sum) per sample are ever read from the merged CountAllocation,
    // so they are summed in shared memory and leave the block as <= 2S global atomics per iteration
    __shared__ unsigned long long sh_stat[BTG_MAX_SAMPLES * 2];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const unsigned long long t_start = ns.phase_ns && blockIdx.x == 0 && threadIdx.x == 0 ? global_timer_ns() : 0;
    for (uint32_t i = tid >> 5; i < n_big; i += nthreads >> 5) {  // large clusters: constructed by a warp, reset by its lane 0
        Cl cl;
        cl.bind(du, sel[i]);
        const uint64_t gidx = group_index(o, cl.g);
        const uint32_t stream_chain = joint ? 0 : chain;
        if (!joint || chain == 1) cl_construct_warp(cl, o, gidx, stream_chain, tid & 31u);
        if ((tid & 31u) == 0) {
            Philox prng, fr;
            if (!joint || chain == 1) {
                prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, stream_chain);
                fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, stream_chain);
            } else {
                prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
                fr.load(cl.rng, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
            }
            cl_reset<false, true>(cl, o, prng);
            prng.save(cl.rng, kRng0);
            fr.save(cl.rng, kRng1);
        }
        __syncwarp();
    }
    for (uint32_t i = n_big + tid; i < n_sel; i += nthreads) {  // initGenotypersCallback: fresh genotypers every chain
        Cl cl;
        cl.bind(du, sel[i]);
        // a thread's FIRST cluster keeps its hot state in the thread's slice of shared memory for the whole chain when it is small
        // (estimateNoise only: the joint mode keeps its genotypers across chains, i.e. across launches, in the arena)
        const bool resident = hot && i == n_big + tid && cl.H <= kHotH;
        if (resident) cl.bind_hot(hot_smem, threadIdx.x, blockDim.x);
        const uint64_t gidx = group_index(o, cl.g);
        Philox prng, fr;
        if (!joint || chain == 1) {
            const uint32_t stream_chain = joint ? 0 : chain;
            cl_construct(cl, o, gidx, stream_chain);
            prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, stream_chain);
            fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, stream_chain);
        } else {
            prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
            fr.load(cl.rng, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
        }
        cl_reset<false, true>(cl, o, prng);
        if (resident && cl.tile_fits_hot()) cl.move_tile_to_hot();
        prng.save(cl.rng, kRng0);
        fr.save(cl.rng, kRng1);
    }
    if (blockIdx.x == 0 && ns.trace) noise_update_block(ns, du.S, prior_shape, prior_scale, o.random_seed, 3, 0, (double)chain, 0, 1, sh_rates);
    grid_barrier(gb);
    if (t_start) ns.phase_ns[7] += global_timer_ns() - t_start;  // construct + reset of every selected cluster
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
#if BTG_NOISE_TIMING
    unsigned long long sub_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // per-thread clock sums of the one-thread sub-steps, flushed once at the end
#endif
    const bool timing = ns.phase_ns && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_prev = timing ? global_timer_ns() : 0;
    auto lap = [&](int phase) {
        if (timing) { const unsigned long long t = global_timer_ns(); ns.phase_ns[phase] += t - t_prev; t_prev = t; }
    };
    for (uint32_t it = 1; it <= iters; it++) {
        if (threadIdx.x < 2 * du.S) sh_stat[threadIdx.x] = 0;
        __syncthreads();
        // phase A: the diplotype caches of the large clusters sel[0 .. n_big) are filled by the whole grid: fill task t =
        // (cluster, part, parts) gives one warp every parts-th round of 32 cache entries of that cluster, so the slowest
        // cluster no longer sets the pace of the iteration with a single warp
        // (tasks are dealt from the LAST warp downwards: the one-thread clusters below occupy the first threads of the grid)
        // That matters in the FIRST iteration of a chain only: reset leaves every haplotype live (H (H + 1) / 2 entries per sample), after one
        // sampleHaplotypeFrequencies the sparse prior keeps a handful.  From the second iteration on a large cluster is ONE warp's job — fill
        // (lanes = entries or terms), sample, count — in the same phase as everything else: no fill tasks whose only content is the bind of a
        // cluster with nothing to fill (80 per warp and iteration on configs[1]), no second phase, one grid barrier less.
        const bool spread = it == 1;
        // the step of a large cluster's owner warp once its caches are filled (lane 0 samples; counts and cache clear by all lanes)
        auto owner_step = [&](Cl &cl, uint32_t lane) {
            const unsigned long long t_in = BTG_NOISE_TIMING && ns.phase_ns ? global_timer_ns() : 0;
            if (lane == 0) noise_iteration_thread(cl, du, T, o, joint && it > o.gibbs_burn_in);
            if (BTG_NOISE_TIMING && ns.phase_ns && lane == 0) { const unsigned long long v = ((global_timer_ns() - t_in) << 32) | cl.c; atomicMax(ns.phase_ns + 5, v); atomicMax(ns.phase_ns + 8 + 3 * (it - 1) + 1, v); }
            __syncwarp();
            const uint32_t n_sub = cl.misc[kNSub];
            for (uint32_t s = 0; s < cl.S; s++) {  // getNoiseCounts
                const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
                uint32_t n0 = 0, c0 = 0;
                const uint32_t g = du.sample_gender[s];
                for (uint32_t j = lane; j < n_sub; j += 32)
                    if ((uint8_t)(cl.tileDiplMult(j, da, db) + cl.tile_ic[j * 2 + g]) == 0) { n0++; c0 += cl.tile_c[j * cl.S + s]; }
                if (n0) { atomicAdd(sh_stat + 2 * s, (unsigned long long)n0); atomicAdd(sh_stat + 2 * s + 1, (unsigned long long)c0); }
            }
            for (uint32_t j = lane; j < cl.S * cl.Dall; j += 32) cl.ucache[j] = nan;  // clearGenotyperCache
            __syncwarp();
        };
        if (spread) {
            for (uint32_t t = (nthreads >> 5) - 1 - (tid >> 5); t < n_fill_tasks; t += nthreads >> 5) {
                const unsigned long long t_in = BTG_NOISE_TIMING && ns.phase_ns ? global_timer_ns() : 0;
                Cl cl;
                cl.bind(du, sel[fill_tasks[3 * t]]);
                cl_fill_cache_warp(cl, T, du.group_ploidy + (size_t)cl.g * du.S, tid & 31u, fill_tasks[3 * t + 1], fill_tasks[3 * t + 2]);
                if (BTG_NOISE_TIMING && ns.phase_ns) { const unsigned long long v = ((global_timer_ns() - t_in) << 32) | cl.c; atomicMax(ns.phase_ns + 4, v); atomicMax(ns.phase_ns + 8 + 3 * (it - 1), v); }
            }
        } else {
            for (uint32_t i = (nthreads >> 5) - 1 - (tid >> 5); i < n_big; i += nthreads >> 5) {
                Cl cl;
                cl.bind(du, sel[i]);
                cl_fill_cache_warp(cl, T, du.group_ploidy + (size_t)cl.g * du.S, tid & 31u, 0, 1);
                __syncwarp();
                owner_step(cl, tid & 31u);
            }
        }
        // ... while the one-thread clusters sel[n_big .. n_sel) take their whole step in the same phase (they do not depend on the
        // fill tasks): the warps that hold no fill task are not idle at the barrier while the large caches are filled
        constexpr uint32_t kAccSamples = 4;
        uint32_t acc_n0[kAccSamples] = {0, 0, 0, 0}, acc_c0[kAccSamples] = {0, 0, 0, 0};
        for (uint32_t i = n_big + tid; i < n_sel; i += nthreads) {  // sampleGenotypesCallback
            const unsigned long long t_in = BTG_NOISE_TIMING && ns.phase_ns ? global_timer_ns() : 0;
#if BTG_NOISE_TIMING
            const bool sub = ns.phase_ns != nullptr;
            long long ck = sub ? clock64() : 0;
            auto tick = [&](int k) { if (sub) { const long long now = clock64(); sub_acc[k] += (unsigned int)(now - ck); ck = now; } };
#else
            auto tick = [](int) {};
#endif
            Cl cl;
            cl.bind(du, sel[i]);
            if (hot && i == n_big + tid && cl.H <= kHotH) {
                cl.bind_hot(hot_smem, threadIdx.x, blockDim.x);
                if (cl.tile_fits_hot()) cl.bind_hot_tile();
            }
            tick(0);
            const uint8_t *ploidy_i = du.group_ploidy + (size_t)cl.g * du.S;
            cl_fill_cache_rows(cl, T, ploidy_i);  // > 4 live haplotypes: entries are filled on demand
            tick(1);
            {
                const uint64_t gidx = group_index(o, cl.g);
                Philox prng, fr;
                prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
                fr.load(cl.rng, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
                tick(2);
                cl_sample_diplotypes<false, true>(cl, T, ploidy_i, joint && it > o.gibbs_burn_in, prng);
                tick(3);
                cl_sample_frequencies(cl, fr);
                tick(4);
                prng.save(cl.rng, kRng0);
                fr.save(cl.rng, kRng1);
                tick(5);
            }
#if BTG_NOISE_TIMING
            if (sub) sub_acc[8]++;
#endif
            if (BTG_NOISE_TIMING && ns.phase_ns) { const unsigned long long v = ((global_timer_ns() - t_in) << 32) | cl.c; atomicMax(ns.phase_ns + 6, v); atomicMax(ns.phase_ns + 8 + 3 * (it - 1) + 2, v); }
            const uint32_t n_sub = cl.misc[kNSub];
            for (uint32_t s = 0; s < cl.S; s++) {  // getNoiseCounts
                const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16;
                uint32_t n0 = 0, c0 = 0;
                const uint32_t g = du.sample_gender[s];
#pragma unroll 4
                for (uint32_t j = 0; j < n_sub; j++) {
                    const uint8_t mm = (uint8_t)(cl.tileDiplMult(j, da, db) + cl.tile_ic[j * 2 + g]), cc = cl.tile_c[j * cl.S + s];
                    n0 += mm == 0; c0 += mm == 0 ? cc : 0u;
                }
                if (s < kAccSamples) { acc_n0[s] += n0; acc_c0[s] += c0; }   // one reduction per phase instead of two shared atomics per cluster
                else if (n0) { atomicAdd(sh_stat + 2 * s, (unsigned long long)n0); atomicAdd(sh_stat + 2 * s + 1, (unsigned long long)c0); }
            }
            tick(6);
            for (uint32_t j = 0; j < cl.S * cl.Dall; j++) cl.ucache[j] = nan;  // clearGenotyperCache
            tick(7);
        }
        for (uint32_t s = 0; s < kAccSamples && s < du.S; s++) {   // the noise counts of this thread's clusters: warp sum, one atomic per warp
            const uint32_t n0 = __reduce_add_sync(0xFFFFFFFFu, acc_n0[s]), c0 = __reduce_add_sync(0xFFFFFFFFu, acc_c0[s]);
            if ((threadIdx.x & 31u) == 0 && n0) { atomicAdd(sh_stat + 2 * s, (unsigned long long)n0); atomicAdd(sh_stat + 2 * s + 1, (unsigned long long)c0); }
        }
        if (spread && n_fill_tasks) grid_barrier(gb);
        lap(0);
        // phase B (first iteration of the chain): sel[0 .. n_big), one WARP each, sample from the caches the grid has filled
        if (spread)
            for (uint32_t i = tid >> 5; i < n_big; i += nthreads >> 5) {
                Cl cl;
                cl.bind(du, sel[i]);
                owner_step(cl, tid & 31u);
            }
        __syncthreads();
        if (threadIdx.x < 2 * du.S && sh_stat[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh_stat[threadIdx.x]);
        grid_barrier(gb);
        lap(1);
        if (blockIdx.x == 0) {
            if (px.world > 1) {  // sharded unit: add up the ranks' statistics over peer memory (comm.cuh) before the draw
                if (threadIdx.x < 2 * du.S) sh_tot[threadIdx.x] = hist[threadIdx.x];
                peer_allreduce_block(px, px.seq0 + it, sh_tot, 2 * du.S);
                if (threadIdx.x < 2 * du.S) hist[threadIdx.x] = sh_tot[threadIdx.x];
                __syncthreads();
            }
            noise_update_block(ns, du.S, prior_shape, prior_scale, o.random_seed, 1, o.gibbs_burn_in < it, (double)chain, (double)it, 1, sh_rates);
        }
        lap(2);
        grid_barrier(gb);
        lap(3);
    }
#if BTG_NOISE_TIMING
    if (ns.phase_ns && sub_acc[8])
        for (int k = 0; k < 9; k++) atomicAdd(ns.phase_ns + 8 + 3 * (size_t)iters + k, sub_acc[k]);
#endif
}

__global__ void k_noise_rng_init(uint32_t *rng, uint32_t seed, uint32_t chain = 0) {
    Philox r;
    r.init(seed, (uint64_t)-1, 0, kRngNoise, chain);
    r.save(rng, 0);
}

// estimateNoise, end: mean of the post-burn-in rates of all chains (added up in chain order) -> setNoiseRates
// (InferenceEngine.cpp:259-264), Poisson rows rebuilt, final trace row "0 0"
__global__ void k_noise_finish(const double *chain_means /* [n_chains][S] sums */, uint32_t n_chains, uint32_t S, double div, double *rates, double *noise_table,
                               double *trace_row) {
    __shared__ double sh[BTG_MAX_SAMPLES];
    if (threadIdx.x < S) {
        double acc = 0;
        for (uint32_t b = 0; b < n_chains; b++) acc += chain_means[(size_t)b * S + threadIdx.x];
        const double r = acc / div;
        rates[threadIdx.x] = r;
        sh[threadIdx.x] = r;
        if (trace_row) { trace_row[0] = 0; trace_row[1] = 0; trace_row[2 + threadIdx.x] = r; }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < S * 256u; i += blockDim.x) noise_table[i] = noiseCountLogPmf(sh[i >> 8], i & 255u);
}

}  // namespace


void btg_unit_free_result(btg_unit *u);

extern "C" {

// ---- count distribution ---------------------------------------------------------------------
btg_count_dist *btg_count_dist_create(uint32_t S, const double *nb_p, const double *nb_size, float prior_shape, float prior_scale) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (S == 0 || S > BTG_MAX_SAMPLES || !nb_p || !nb_size) { set_error("bad count distribution arguments"); return nullptr; }
    for (uint32_t s = 0; s < S; s++)
        if (!(nb_p[s] > 0 && nb_p[s] < 1 && nb_size[s] > 0)) { set_error("negative binomial parameters out of range for sample %u", s); return nullptr; }
    auto *cd = new btg_count_dist();
    cd->S = S;
    cd->prior_shape = prior_shape;
    cd->prior_scale = prior_scale;
    cd->h_p.assign(nb_p, nb_p + S);
    cd->h_size.assign(nb_size, nb_size + S);
    bool ok = true;
    cd->p = upload(nb_p, S, ok);
    cd->size = upload(nb_size, S, ok);
    std::vector<double> ones(S, 1.0);
    cd->rates = upload(ones.data(), S, ok);
    ok = ok && btg::dmalloc(&cd->genomic, (size_t)S * 65536 * sizeof(double)) == cudaSuccess;
    ok = ok && btg::dmalloc(&cd->noise, (size_t)S * 256 * sizeof(double)) == cudaSuccess;
    if (!ok) { set_error("count distribution allocation failed"); btg_count_dist_free(cd); return nullptr; }
    auto s = ctx().stream;
    k_genomic_table<<<(S * 65536 + 255) / 256, 256, 0, s>>>(cd->p, cd->size, S, cd->genomic);
    BTG_LAUNCHED();
    k_noise_table<<<(S * 256 + 255) / 256, 256, 0, s>>>(cd->rates, S, cd->noise);
    BTG_LAUNCHED();
    if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("count table kernels failed: %s", cudaGetErrorString(cudaGetLastError())); btg_count_dist_free(cd); return nullptr; }
    return cd;
}

void btg_nb_moments_to_parameters(double mean, double var, uint32_t multiplicity, double *p_out, double *size_out) {
    const double max_p = 0.99;  // NegativeBinomialDistribution.cpp:38,68-79
    if (max_p < (mean / var)) var = mean / max_p;
    if (p_out) *p_out = mean / var;
    if (size_out) *size_out = std::pow(mean, 2) / (var - mean) / multiplicity;  // CountDistribution.cpp:115-116
}

int btg_count_dist_set_noise_rates(btg_count_dist *cd, const double *rates) {
    BTG_REQUIRE_INIT();
    if (!cd || !rates) { set_error("null argument"); return BTG_EINVAL; }
    auto s = ctx().stream;
    BTG_CUDA(cudaMemcpyAsync(cd->rates, rates, cd->S * sizeof(double), cudaMemcpyHostToDevice, s));
    k_noise_table<<<(cd->S * 256 + 255) / 256, 256, 0, s>>>(cd->rates, cd->S, cd->noise);
    BTG_LAUNCHED();
    BTG_CUDA(cudaStreamSynchronize(s));
    return BTG_OK;
}

int btg_count_dist_get_noise_rates(const btg_count_dist *cd, double *out) {
    BTG_REQUIRE_INIT();
    if (!cd || !out) { set_error("null argument"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy(out, cd->rates, cd->S * sizeof(double), cudaMemcpyDeviceToHost));
    return BTG_OK;
}

int btg_count_dist_tables(const btg_count_dist *cd, double *genomic_out, double *noise_out) {
    BTG_REQUIRE_INIT();
    if (!cd) { set_error("null argument"); return BTG_EINVAL; }
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    if (genomic_out) BTG_CUDA(cudaMemcpy(genomic_out, cd->genomic, (size_t)cd->S * 65536 * sizeof(double), cudaMemcpyDeviceToHost));
    if (noise_out) BTG_CUDA(cudaMemcpy(noise_out, cd->noise, (size_t)cd->S * 256 * sizeof(double), cudaMemcpyDeviceToHost));
    return BTG_OK;
}

void btg_count_dist_free(btg_count_dist *cd) {
    if (!cd) return;
    btg::dfree(cd->p); btg::dfree(cd->size); btg::dfree(cd->rates); btg::dfree(cd->genomic); btg::dfree(cd->noise);
    delete cd;
}

// ---- unit -------------------------------------------------------------------------------------
btg_unit *btg_unit_upload(const btg_unit_desc *d) { return btg_unit_upload_dev(d, nullptr, 0, 0); }

btg_unit *btg_unit_upload_dev(const btg_unit_desc *d, const btg_unit_desc *dev, uint64_t n_vh_dev, uint64_t n_vh_bits_dev) {
    if (!ctx().ready) { set_error("btg_init() has not been called"); return nullptr; }
    if (!d || d->n_samples == 0 || d->n_samples > BTG_MAX_SAMPLES) { set_error("bad unit descriptor"); return nullptr; }
    // row-level arrays may come from the device (dev->field != NULL): device-to-device copy instead of a host round trip
    auto from = [&](auto host_ptr, auto dev_ptr, size_t n, bool &ok_flag) {
        using T = std::remove_cv_t<std::remove_pointer_t<decltype(host_ptr)>>;
        if (!dev_ptr) return upload(host_ptr, n, ok_flag);
        T *p = nullptr;
        if (btg::dmalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) { ok_flag = false; return (T *)nullptr; }
        if (n && cudaMemcpyAsync(p, dev_ptr, n * sizeof(T), cudaMemcpyDeviceToDevice, ctx().stream) != cudaSuccess) ok_flag = false;
        return p;
    };
#define BTG_DEVF(f) (dev ? dev->f : nullptr)
    const uint32_t S = d->n_samples, G = d->n_groups, C = d->n_clusters;
    auto *u = new btg_unit();
    bool ok = true;
    auto keep = [&](auto *p) { u->allocs.push_back((void *)p); return p; };
    const bool up_timing = getenv("BTG_UPLOAD_TIMING") != nullptr;
    auto up_t0 = std::chrono::steady_clock::now();
    auto up_lap = [&](const char *what) {
        if (!up_timing) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[btgpu] unit upload: %s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - up_t0).count());
        up_t0 = now;
    };
    const uint64_t rows = d->cl_kmer_off[C], nvar = d->cl_var_off[C], n_vh = BTG_DEVF(kmer_vh_off) ? n_vh_dev : d->kmer_vh_off[rows];
    DevUnit &du = u->du;
    du.S = S; du.G = G; du.C = C;
    // arena layout: with many samples a cluster is worked on by a warp (lane = sample, gibbs_wide.cu) and gets a dense slot of its own
    du.wide = getenv("BTG_WIDE") ? (atoi(getenv("BTG_WIDE")) != 0) : (S >= kWideMinSamples);
    du.sample_gender = keep(upload(d->sample_gender, S, ok));
    du.group_ploidy = keep(upload(d->group_ploidy, (size_t)G * S, ok));
    du.group_cluster_off = keep(upload(d->group_cluster_off, G + 1, ok));
    du.cluster_idx = keep(upload(d->cluster_idx, C, ok));
    du.cl_nhap = keep(upload(d->cl_nhap, C, ok));
    du.cl_kmer_off = keep(upload(d->cl_kmer_off, C + 1, ok));
    du.cl_var_off = keep(upload(d->cl_var_off, C + 1, ok));
    du.cl_mult_off = keep(upload(d->cl_mult_off, C + 1, ok));
    du.mult = keep(from(d->mult, BTG_DEVF(mult), d->cl_mult_off[C], ok));
    du.k_has_counts = keep(from(d->k_has_counts, BTG_DEVF(k_has_counts), rows, ok));
    du.k_counts = keep(from(d->k_counts, BTG_DEVF(k_counts), rows * S, ok));
    du.k_ic = keep(from(d->k_ic, BTG_DEVF(k_ic), rows * 2, ok));
    du.cl_uniq_off = keep(upload(d->cl_uniq_off, C + 1, ok));
    du.uniq_idx = keep(from(d->uniq_idx, BTG_DEVF(uniq_idx), d->cl_uniq_off[C], ok));
    du.kmer_vh_off = keep(from(d->kmer_vh_off, BTG_DEVF(kmer_vh_off), rows + 1, ok));
    du.vh_var = keep(from(d->vh_var, BTG_DEVF(vh_var), n_vh, ok));
    du.vh_bits_off = keep(from(d->vh_bits_off, BTG_DEVF(vh_bits_off), n_vh + 1, ok));
    du.vh_bits = keep(from(d->vh_bits, BTG_DEVF(vh_bits), BTG_DEVF(vh_bits_off) ? n_vh_bits_dev : d->vh_bits_off[n_vh], ok));
    du.cl_hapvar_off = keep(upload(d->cl_hapvar_off, C + 1, ok));
    du.hap_alleles = keep(from(d->hap_alleles, BTG_DEVF(hap_alleles), d->cl_hapvar_off[C], ok));
    du.var_nalleles = keep(upload(d->var_nalleles, nvar, ok));
    du.var_dep = keep(upload(d->var_dep, nvar, ok));
    // nested groups and multicluster k-mers
    {
        const uint64_t n_multi = d->cl_multi_off[C];
        uint64_t n_hap = 0, n_shared = 0;
        std::vector<uint64_t> hap_start(C + 1, 0);
        for (uint32_t c = 0; c < C; c++) hap_start[c + 1] = hap_start[c] + d->cl_nhap[c];
        n_hap = hap_start[C];
        if (n_multi && (BTG_DEVF(k_shared) || BTG_DEVF(k_has_counts))) {
            set_error("units with multicluster k-mers must pass k_shared and k_has_counts as host arrays (they are validated on the host)");
            ok = false;
        }
        for (uint32_t c = 0; c < C && ok; c++)
            for (uint64_t i = d->cl_multi_off[c]; i < d->cl_multi_off[c + 1]; i++) {
                const uint64_t r = d->cl_kmer_off[c] + d->multi_idx[i];
                if (d->k_shared[r] == 0xFFFFFFFFu || !d->k_has_counts[r]) { set_error("cluster %u: multicluster k-mer row %llu has no shared count record (k_shared)", c, (unsigned long long)r); ok = false; break; }
                n_shared = std::max<uint64_t>(n_shared, (uint64_t)d->k_shared[r] + 1);
            }
        du.k_shared = keep(from(d->k_shared, BTG_DEVF(k_shared), rows, ok));
        du.cl_multi_off = keep(upload(d->cl_multi_off, C + 1, ok));
        du.multi_idx = keep(upload(d->multi_idx, n_multi, ok));
        du.hap_start = keep(upload(hap_start.data(), C + 1, ok));
        du.hap_nested_off = keep(upload(d->hap_nested_off, n_hap + 1, ok));
        du.hap_nested = keep(upload(d->hap_nested, d->hap_nested_off[n_hap], ok));
        du.cl_dep_off = keep(upload(d->cl_dep_off, C + 1, ok));
        const uint64_t n_dep = d->cl_dep_off[C];
        du.dep_cluster = keep(upload(d->dep_cluster, n_dep, ok));
        du.dep_var_off = keep(upload(d->dep_var_off, n_dep + 1, ok));
        du.dep_var = keep(upload(d->dep_var, d->dep_var_off[n_dep], ok));
        du.group_src_off = keep(upload(d->group_src_off, G + 1, ok));
        du.group_src = keep(upload(d->group_src, d->group_src_off[G], ok));
        // out-edges as a CSR over clusters, each vertex's targets in the order given (VariantClusterGroup.cpp:94-104)
        std::vector<uint64_t> cl_edge_off(C + 1, 0);
        const uint64_t n_edges = d->group_edge_off[G];
        std::vector<uint32_t> edge_dst(n_edges);
        std::vector<uint32_t> nested_groups, nest_slot(C, 0xFFFFFFFFu);
        uint32_t n_nest = 0;
        for (uint32_t g = 0; g < G && ok; g++) {
            const uint64_t c0 = d->group_cluster_off[g], n = d->group_cluster_off[g + 1] - c0;
            if (n == 0) { set_error("group %u has no cluster", g); ok = false; break; }
            for (uint64_t e = d->group_edge_off[g]; e < d->group_edge_off[g + 1]; e++) {
                if (d->group_edge_src[e] >= n || d->group_edge_dst[e] >= n) { set_error("group %u: edge %llu out of range", g, (unsigned long long)e); ok = false; break; }
                cl_edge_off[c0 + d->group_edge_src[e] + 1]++;
            }
            for (uint64_t e = d->group_src_off[g]; e < d->group_src_off[g + 1]; e++)
                if (d->group_src[e] >= n) { set_error("group %u: source vertex out of range", g); ok = false; break; }
            if (n > 1) {
                nested_groups.push_back(g);
                for (uint64_t c = c0; c < c0 + n; c++) nest_slot[c] = n_nest++;
                if (d->group_edge_off[g + 1] - d->group_edge_off[g] + (d->group_src_off[g + 1] - d->group_src_off[g]) != n) {
                    set_error("group %u: %llu clusters need a forest of %llu sources + edges", g, (unsigned long long)n, (unsigned long long)n); ok = false; break;
                }
            }
        }
        for (uint32_t c = 0; c < C; c++) cl_edge_off[c + 1] += cl_edge_off[c];
        if (ok) {
            std::vector<uint64_t> fill(cl_edge_off.begin(), cl_edge_off.end() - 1);
            for (uint32_t g = 0; g < G; g++) {
                const uint64_t c0 = d->group_cluster_off[g];
                for (uint64_t e = d->group_edge_off[g]; e < d->group_edge_off[g + 1]; e++) edge_dst[fill[c0 + d->group_edge_src[e]]++] = d->group_edge_dst[e];
            }
        }
        du.cl_edge_off = keep(upload(cl_edge_off.data(), C + 1, ok));
        du.edge_dst = keep(upload(edge_dst.data(), n_edges, ok));
        du.nested_groups = keep(upload(nested_groups.data(), nested_groups.size(), ok));
        du.n_nested_groups = (uint32_t)nested_groups.size();
        du.nest_slot = keep(upload(nest_slot.data(), C, ok));
        auto dmalloc = [&](auto *&dst, size_t n) {
            void *p = nullptr;
            if (btg::dmalloc(&p, (n ? n : 1) * sizeof(*dst)) != cudaSuccess) { ok = false; dst = nullptr; return; }
            u->allocs.push_back(p);
            dst = static_cast<std::remove_reference_t<decltype(dst)>>(p);
        };
        dmalloc(du.src_mut, d->group_src_off[G]); dmalloc(du.edge_mut, n_edges);
        dmalloc(du.dfs_order, nested_groups.empty() ? 0 : C); dmalloc(du.dfs_stack, nested_groups.empty() ? 0 : C);
        dmalloc(du.shared_mult, n_shared * S);
        dmalloc(du.nest_pl, (size_t)n_nest * S); dmalloc(du.nest_k, (size_t)n_nest * S);
        dmalloc(du.nest_n, (size_t)n_nest * S * 2); dmalloc(du.nest_f, (size_t)n_nest * S * 4);
        u->n_shared = n_shared;
    }
    // result offsets
    u->n_variants = nvar;
    u->h_valt_off.assign(nvar + 1, 0);
    u->h_allele_off.assign(nvar + 1, 0);
    u->h_geno_off.assign(nvar + 1, 0);
    for (uint64_t v = 0; v < nvar; v++) {
        const uint64_t nA = d->var_nalleles[v];
        u->h_valt_off[v + 1] = u->h_valt_off[v] + nA;
        u->h_allele_off[v + 1] = u->h_allele_off[v] + S * nA;
        u->h_geno_off[v + 1] = u->h_geno_off[v] + S * nA * (nA + 1) / 2;
    }
    u->n_alleles_total = u->h_valt_off[nvar];
    du.valt_off = keep(upload(u->h_valt_off.data(), nvar + 1, ok));
    // arena layout + cost order
    u->h_layout.resize(C);
    u->h_fill_cost.assign(C, 0);
    u->h_nhap.assign(d->cl_nhap, d->cl_nhap + C);
    u->h_group_cluster_off.assign(d->group_cluster_off, d->group_cluster_off + G + 1);
    u->h_cl_var_off.assign(d->cl_var_off, d->cl_var_off + C + 1);
    struct Dims { uint32_t H, K, nv, nu, nal, Dall, nm; };
    up_lap("descriptor arrays -> device");
    std::vector<Dims> dims(C);
    std::vector<uint64_t> cost(C);
    for (uint32_t g = 0; g < G; g++) {
        for (uint64_t c = d->group_cluster_off[g]; c < d->group_cluster_off[g + 1]; c++) {
            const uint32_t H = d->cl_nhap[c];
            const uint32_t K = (uint32_t)(d->cl_kmer_off[c + 1] - d->cl_kmer_off[c]);
            const uint32_t nv = (uint32_t)(d->cl_var_off[c + 1] - d->cl_var_off[c]);
            const uint32_t nu = (uint32_t)(d->cl_uniq_off[c + 1] - d->cl_uniq_off[c]);
            const uint32_t nal = (uint32_t)(u->h_valt_off[d->cl_var_off[c + 1]] - u->h_valt_off[d->cl_var_off[c]]);
            if (H == 0 || H >= 0xFFFE) { set_error("cluster %llu: invalid number of haplotypes %u", (unsigned long long)c, H); ok = false; break; }
            ClusterLayout &L = u->h_layout[c];
            L.group = g;
            L.n_alleles = nal;
            L.Dall = (H + 1) * (H + 2) / 2;
            const uint32_t nm = (uint32_t)(d->cl_multi_off[c + 1] - d->cl_multi_off[c]);
            dims[c] = Dims{H, K, nv, nu, nal, L.Dall, nm};
            u->h_fill_cost[c] = (uint32_t)std::min<uint64_t>(0xFFFFFFFFu, (uint64_t)S * ((uint64_t)H * (H + 1) / 2) * (nu / 10 + 1));
            cost[c] = (uint64_t)S * ((uint64_t)H * (H + 1) / 2) * 8 + nu + (uint64_t)H * K / 16;
            u->max_h = std::max(u->max_h, H);
        }
    }
    std::vector<uint32_t> order(C);
    std::iota(order.begin(), order.end(), 0u);
    // clusters of single-cluster groups first (one thread each), then the clusters of nested groups (one thread per group)
    auto is_nested = [&](uint32_t c) { const uint32_t g = u->h_layout[c].group; return d->group_cluster_off[g + 1] - d->group_cluster_off[g] > 1; };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const bool na = is_nested(a), nb = is_nested(b);
        return na != nb ? nb : cost[a] > cost[b];
    });
    du.n_regular = 0;
    while (du.n_regular < C && !is_nested(order[du.n_regular])) du.n_regular++;
    // chain-split clusters (default mode): large clusters of single-cluster groups get kChainSplit arena positions; the extra
    // positions follow the C regular ones
    const uint32_t split_cost = getenv("BTG_SPLIT_COST") ? (uint32_t)strtoul(getenv("BTG_SPLIT_COST"), nullptr, 10) : kSplitFillCost;  // tests: 0 splits everything
    std::vector<uint32_t> split_cluster, split_of(C ? C : 1, NONE32);
    for (uint32_t i = 0; i < du.n_regular; i++)
        if (u->h_fill_cost[order[i]] > split_cost) { split_of[order[i]] = (uint32_t)split_cluster.size(); split_cluster.push_back(order[i]); }
    const uint32_t n_split = (uint32_t)split_cluster.size();
    std::vector<uint32_t> ext_cluster(order);  // cluster at every arena position
    std::vector<uint32_t> split_pos((size_t)n_split * kChainSplit + 1, NONE32);
    for (uint32_t j = 0; j < n_split; j++)
        for (uint32_t v = 1; v < kChainSplit; v++) { split_pos[(size_t)j * kChainSplit + v] = (uint32_t)ext_cluster.size(); ext_cluster.push_back(split_cluster[j]); }
    const uint32_t n_pos = (uint32_t)ext_cluster.size();
    // one arena slot per warp of the position order, sized by the largest cluster in it (wide layout: one slot per position)
    const uint32_t per_slot = du.wide ? 1u : 32u;
    // dense diplotype caches of the wide layout: 16 B x S x (H+1)(H+2)/2 per cluster.  Every cluster gets them as long as the total stays within a
    // quarter of the free HBM; otherwise the cap on S x Dall is halved until it does (clusters above the cap recompute their sums, which is
    // correct but costs them two walks of the live diplotypes per iteration).  BTG_WIDE_CACHE_CAP fixes the cap (tests lower it).
    uint64_t cache_cap = kWideCacheCap;
    if (getenv("BTG_WIDE_CACHE_CAP")) cache_cap = strtoull(getenv("BTG_WIDE_CACHE_CAP"), nullptr, 10);
    else if (du.wide) {
        const size_t free_b = btg::free_device_memory();
        cache_cap = 1ull << 24;
        for (;;) {
            uint64_t bytes = 0;
            for (uint32_t c = 0; c < C; c++) { const uint64_t e = (uint64_t)S * dims[c].Dall; if (e <= cache_cap || dims[c].nm) bytes += 16 * e; }
            if (bytes <= free_b / 4 || cache_cap <= kWideCacheCap) break;
            cache_cap >>= 1;
        }
    }
    const uint32_t n_slots = (n_pos + per_slot - 1) / per_slot;
    u->h_slots.assign(n_slots ? n_slots : 1, SlotLayout{});
    uint64_t f64_total = 0, u32_total = 0, u8_total = 0;
    for (uint32_t w = 0; w < n_slots; w++) {
        SlotLayout &SL = u->h_slots[w];
        for (uint64_t i = (uint64_t)w * per_slot; i < std::min<uint64_t>(n_pos, (uint64_t)w * per_slot + per_slot); i++) {
            const Dims &D = dims[ext_cluster[i]];
            if (i < C) u->h_layout[order[i]].pos = i;
            SL.H = std::max(SL.H, D.H); SL.K = std::max(SL.K, D.K); SL.nvar = std::max(SL.nvar, D.nv);
            SL.n_uniq = std::max(SL.n_uniq, D.nu); SL.n_alleles = std::max(SL.n_alleles, D.nal); SL.Dall = std::max(SL.Dall, D.Dall);
            SL.n_multi = std::max(SL.n_multi, D.nm);
            // a unit with nested groups runs its lock-step joint mode on the warp-per-group kernel (lane = sample) whatever the layout
            if (du.wide || du.n_regular < C) SL.cum_rows = 1;
        }
        const ArenaSizes a = arena_sizes(S, SL.H, SL.K, SL.nvar, SL.n_uniq, SL.n_alleles, SL.Dall, SL.n_multi, du.wide, cache_cap, SL.cum_rows);
        SL.has_cache = arena_has_cache(S, SL.Dall, SL.n_multi, du.wide, cache_cap);
        SL.f64_off = f64_total; SL.u32_off = u32_total; SL.u8_off = u8_total;
        f64_total += a.f64 * per_slot; u32_total += a.u32 * per_slot; u8_total += a.u8 * per_slot;
    }
    std::vector<uint64_t> tile_off(C ? C : 1, ~0ull);
    uint64_t tile_total = 0;
    for (uint32_t c = 0; c < C; c++)
        if (!du.wide && u->h_fill_cost[c] > (uint64_t)big_fill_cost() * S) { tile_off[c] = tile_total; tile_total += ((uint64_t)dims[c].nu * (dims[c].H + S + 2) + 31) & ~31ull; }
    du.big_tile_off = keep(upload(tile_off.data(), C, ok));
    uint8_t *tile_pool = nullptr;
    ok = ok && btg::dmalloc(&tile_pool, tile_total + 32) == cudaSuccess;
    keep(tile_pool);
    du.big_tile_pool = tile_pool;
    up_lap("host layout");
    for (uint32_t j = 0; j < n_split; j++) split_pos[(size_t)j * kChainSplit] = u->h_layout[split_cluster[j]].pos;
    du.n_split = n_split;
    du.split_cluster = keep(upload(split_cluster.data(), n_split, ok));
    du.split_pos = keep(upload(split_pos.data(), (size_t)n_split * kChainSplit, ok));
    du.split_of = keep(upload(split_of.data(), C, ok));
    du.layout = keep(upload(u->h_layout.data(), C, ok));
    du.slots = keep(upload(u->h_slots.data(), u->h_slots.size(), ok));
    du.order = keep(upload(order.data(), C, ok));
    double *f64_pool = nullptr; uint32_t *u32_pool = nullptr; uint8_t *u8_pool = nullptr; double *lg = nullptr;
    ok = ok && btg::dmalloc(&f64_pool, (f64_total + 1) * sizeof(double)) == cudaSuccess;
    ok = ok && btg::dmalloc(&u32_pool, (u32_total + 1) * sizeof(uint32_t)) == cudaSuccess;
    ok = ok && btg::dmalloc(&u8_pool, u8_total + 8) == cudaSuccess;
    const uint32_t n_lg = u->max_h + 2 * S + 4;
    ok = ok && btg::dmalloc(&lg, n_lg * sizeof(double)) == cudaSuccess;
    keep(f64_pool); keep(u32_pool); keep(u8_pool); keep(lg);
    u->f64_total = f64_total; u->u32_total = u32_total; u->u8_total = u8_total; u->tile_total = tile_total;
    du.f64_pool = f64_pool; du.u32_pool = u32_pool; du.u8_pool = u8_pool; du.lgamma_int = lg;
    up_lap("arena allocation");
    if (ok) {
        k_lgamma_int<<<(n_lg + 127) / 128, 128, 0, ctx().stream>>>(lg, n_lg);
        BTG_LAUNCHED();
        ok = cudaStreamSynchronize(ctx().stream) == cudaSuccess;
    }
    if (!ok) {
        if (!*btg_last_error()) set_error("unit upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        btg_unit_free(u);
        return nullptr;
    }
    return u;
}

void btg_unit_free(btg_unit *u) {
    if (!u) return;
    cudaStreamSynchronize(ctx().stream);
    for (void *p : u->allocs) btg::dfree(p);
    btg_unit_free_result(u);
    delete u;
}

}  // extern "C"

namespace {
// device-side result arrays owned by the unit (allocated on first use)
struct DevResult {
    ResultView R{};
    std::vector<void *> allocs;
    uint64_t nv = 0, nall = 0, ngen = 0, nalt = 0;
    ~DevResult() { for (void *p : allocs) btg::dfree(p); }
};

DevResult *unit_result(btg_unit *u) {
    if (u->res_view) return static_cast<DevResult *>(u->res_view);
    auto *dr = new DevResult();
    const uint32_t S = u->du.S;
    const uint64_t nv = u->n_variants, nall = u->h_allele_off[nv], ngen = u->h_geno_off[nv], nalt = u->h_valt_off[nv];
    dr->nv = nv; dr->nall = nall; dr->ngen = ngen; dr->nalt = nalt;
    bool ok = true;
    auto mk = [&](auto *&dst, size_t n) {
        void *d = nullptr;
        if (btg::dmalloc(&d, (n ? n : 1) * sizeof(*dst)) != cudaSuccess) { ok = false; return; }
        dr->allocs.push_back(d);
        dst = static_cast<std::remove_reference_t<decltype(dst)>>(d);
    };
    dr->R.allele_off = upload(u->h_allele_off.data(), nv + 1, ok); dr->allocs.push_back((void *)dr->R.allele_off);
    dr->R.geno_off = upload(u->h_geno_off.data(), nv + 1, ok); dr->allocs.push_back((void *)dr->R.geno_off);
    dr->R.valt_off = u->du.valt_off;
    mk(dr->R.gt, nv * S * 2); mk(dr->R.gq, nv * S); mk(dr->R.gpp, ngen); mk(dr->R.app, nall);
    mk(dr->R.nak, nall); mk(dr->R.fak, nall); mk(dr->R.mac, nall); mk(dr->R.saf, nall); mk(dr->R.ploidy, nv * S);
    mk(dr->R.an, nv); mk(dr->R.ac, nalt); mk(dr->R.af, nalt); mk(dr->R.acp, nalt); mk(dr->R.anc, nalt); mk(dr->R.hc, nv);
    if (!ok) { delete dr; return nullptr; }
    u->res_view = dr;
    return dr;
}
}  // namespace

extern "C" {

int btg_estimate_genotypes_async(btg_unit *u, const btg_count_dist *cd, const btg_gibbs_opts *opts, void *stream) {
    BTG_REQUIRE_INIT();
    if (!u || !cd || !opts) { set_error("null argument"); return BTG_EINVAL; }
    if (cd->S != u->du.S) { set_error("count distribution has %u samples, unit has %u", cd->S, u->du.S); return BTG_EINVAL; }
    DevResult *dr = unit_result(u);
    if (!dr) { set_error("result allocation failed"); return BTG_ENOMEM; }
    Tables T{cd->genomic, cd->noise};
    if (u->du.n_nested_groups)  // KmerCounts::multiplicities start at zero in a fresh run (KmerCounts.hpp:100)
        BTG_CUDA(cudaMemsetAsync(u->du.shared_mult, 0, (size_t)u->n_shared * u->du.S, pick_stream(stream)));
    if (u->du.wide) {
        BTG_CUDA(wide_estimate_genotypes(u->du, T, *opts, dr->R, pick_stream(stream)));
        if (u->du.n_split) {
            k_merge_split<<<(u->du.n_split + 63) / 64, 64, 0, pick_stream(stream)>>>(u->du, *opts, dr->R);
            BTG_LAUNCHED();
            BTG_CUDA(cudaGetLastError());
        }
        return BTG_OK;
    }
    if (u->du.n_nested_groups) {
        k_estimate_genotypes_nested<<<(u->du.n_nested_groups + 63) / 64, 64, 0, pick_stream(stream)>>>(u->du, T, *opts, dr->R);
        BTG_LAUNCHED();
        BTG_CUDA(cudaGetLastError());
    }
    if (u->du.n_regular) {
        static const int occ = getenv("BTG_GIBBS_OCC") ? atoi(getenv("BTG_GIBBS_OCC")) : 8;
        const int reconverge = getenv("BTG_GIBBS_SYNC") ? atoi(getenv("BTG_GIBBS_SYNC")) : 1;
        const unsigned grid = (u->du.n_regular + u->du.n_split * kChainSplit + 63) / 64;
        unsigned long long *dbg = nullptr;
        if (getenv("BTG_GIBBS_TIMING")) { btg::dmalloc(&dbg, 32); cudaMemset(dbg, 0, 32); }
        // hot state of the small clusters in shared memory (gibbs_core.cuh): hot_bytes(S) per thread, as long as the blocks of an SM still fit
        const size_t hot_smem = (size_t)hot_bytes(u->du.S) * 64;
        static const int hot_env = getenv("BTG_HOT") ? atoi(getenv("BTG_HOT")) : 1;
        const int hot = hot_env && hot_smem * 6 <= 220 * 1024;
        const size_t smem = hot ? hot_smem : 0;
        if (hot) {
            cudaFuncSetAttribute(k_estimate_genotypes<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_estimate_genotypes<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_estimate_genotypes<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        if (occ >= 16) k_estimate_genotypes<16><<<grid, 64, smem, pick_stream(stream)>>>(u->du, T, *opts, dr->R, reconverge, dbg, hot);
        else if (occ >= 12) k_estimate_genotypes<12><<<grid, 64, smem, pick_stream(stream)>>>(u->du, T, *opts, dr->R, reconverge, dbg, hot);
        else k_estimate_genotypes<8><<<grid, 64, smem, pick_stream(stream)>>>(u->du, T, *opts, dr->R, reconverge, dbg, hot);
        BTG_LAUNCHED();
        BTG_CUDA(cudaGetLastError());
        if (u->du.n_split) {
            k_merge_split<<<(u->du.n_split + 63) / 64, 64, 0, pick_stream(stream)>>>(u->du, *opts, dr->R);
            BTG_LAUNCHED();
            BTG_CUDA(cudaGetLastError());
        }
        if (dbg) {
            unsigned long long h[4];
            cudaStreamSynchronize(pick_stream(stream));
            cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost);
            btg::dfree(dbg);
            const uint32_t c = (uint32_t)(h[0] & 0xFFFFFFFFu);
            fprintf(stderr, "[btgpu] k_estimate_genotypes: slowest thread %.1f ms (cluster %u, H %u, fill cost %u); mean thread %.2f ms over %u clusters; clusters with H > 4 hold %.1f %% of the thread time\n",
                    (h[0] >> 32) / 1e6, c, c < u->du.C ? u->h_nhap[c] : 0, c < u->du.C ? u->h_fill_cost[c] : 0, h[1] / 1e6 / std::max(1u, u->du.n_regular), u->du.n_regular, 100.0 * h[2] / std::max<unsigned long long>(1, h[1]));
        }
    }
    return BTG_OK;
}

int btg_unit_download_result(btg_unit *u, btg_genotype_result *out, void *stream) {
    BTG_REQUIRE_INIT();
    if (!u || !out) { set_error("null argument"); return BTG_EINVAL; }
    if (out->n_variants != u->n_variants) { set_error("result sized for %llu variants, unit has %llu", (unsigned long long)out->n_variants, (unsigned long long)u->n_variants); return BTG_EINVAL; }
    DevResult *dr = unit_result(u);
    if (!dr) { set_error("result allocation failed"); return BTG_ENOMEM; }
    const uint32_t S = u->du.S;
    auto s = pick_stream(stream);
    const ResultView &R = dr->R;
#define BTG_DL(field, n) BTG_CUDA(cudaMemcpyAsync(out->field, R.field, (n) * sizeof(*R.field), cudaMemcpyDeviceToHost, s))
    BTG_DL(gt, dr->nv * S * 2); BTG_DL(gq, dr->nv * S); BTG_DL(gpp, dr->ngen); BTG_DL(app, dr->nall);
    BTG_DL(nak, dr->nall); BTG_DL(fak, dr->nall); BTG_DL(mac, dr->nall); BTG_DL(saf, dr->nall); BTG_DL(ploidy, dr->nv * S);
    BTG_DL(an, dr->nv); BTG_DL(ac, dr->nalt); BTG_DL(af, dr->nalt); BTG_DL(acp, dr->nalt); BTG_DL(anc, dr->nalt); BTG_DL(hc, dr->nv);
#undef BTG_DL
    BTG_CUDA(cudaStreamSynchronize(s));
    return BTG_OK;
}

int btg_estimate_genotypes(btg_unit *u, const btg_count_dist *cd, const btg_gibbs_opts *opts, btg_genotype_result *out) {
    if (!out) { set_error("null argument"); return BTG_EINVAL; }
    int rc = btg_estimate_genotypes_async(u, cd, opts, nullptr);
    if (rc != BTG_OK) return rc;
    return btg_unit_download_result(u, out, nullptr);
}

}  // extern "C"
void btg_unit_free_result(btg_unit *u) {
    if (u->res_view) { delete static_cast<DevResult *>(u->res_view); u->res_view = nullptr; }
}
extern "C" {

int btg_unit_cluster_tally(const btg_unit *u, uint32_t cluster, uint32_t *tally_out, uint64_t n) {
    BTG_REQUIRE_INIT();
    if (!u || cluster >= u->du.C || !tally_out) { set_error("bad argument"); return BTG_EINVAL; }
    const ClusterLayout &L = u->h_layout[cluster];
    const uint32_t H = u->h_nhap[cluster], S = u->du.S;
    const uint64_t need = (uint64_t)L.Dall * S;
    if (n < need) { set_error("tally buffer too small"); return BTG_EINVAL; }
    // tally sits after obs[H], uniq[2*n_uniq], cnt[H*nvar] in the slot's u32 arrays (see Cl::bind), lane-interleaved
    (void)H;
    const uint32_t st = u->du.wide ? 1u : 32u;
    const SlotLayout &SL = u->h_slots[u->du.wide ? L.pos : L.pos >> 5];
    const uint64_t off = SL.u32_off + (u->du.wide ? 0u : L.pos & 31u) + ((uint64_t)SL.H + 2ull * SL.n_uniq + (uint64_t)SL.H * SL.nvar) * st;
    BTG_CUDA(cudaStreamSynchronize(ctx().stream));
    BTG_CUDA(cudaMemcpy2D(tally_out, sizeof(uint32_t), u->du.u32_pool + off, st * sizeof(uint32_t), sizeof(uint32_t), need, cudaMemcpyDeviceToHost));
    return BTG_OK;
}

// large clusters of the lock-step chains: their cache fill is spread over the grid as fill tasks.  One-thread kernel (gibbs.cu): the
// clusters with a dense tile, cost PER SAMPLE (a one-thread cluster walks its samples in turn).  Warp kernel (gibbs_wide.cu, lane =
// sample): the same per-lane cost bound, for single-cluster groups whose slot holds the dense caches.
static bool lockstep_is_big(const btg_unit *u, uint32_t c, bool warp_kernel) {
    const uint32_t S = u->du.S;
    if (!(u->h_fill_cost[c] > (uint64_t)(warp_kernel ? kBigFillCostWarp : big_fill_cost()) * S)) return false;
    if (!warp_kernel) return true;
    const uint32_t g = u->h_layout[c].group;
    if (u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] != 1) return false;
    return u->h_slots[u->du.wide ? u->h_layout[c].pos : u->h_layout[c].pos >> 5].has_cache != 0;
}
// fill tasks of one large cluster: (index in sel, part, parts).  One-thread kernel: one warp per 32 cache entries (S x diplotypes,
// upper bound; 4 with >= 16 k-mers per entry), at most 64.  Warp kernel: a task takes every parts-th PAIR of live haplotypes for
// all samples at once (lane = sample): one part per two pairs of the full enumeration, at most 64.
static void lockstep_fill_tasks(const btg_unit *u, uint32_t c, uint32_t sel_idx, bool warp_kernel, std::vector<uint32_t> &tasks) {
    const uint64_t H = u->h_nhap[c], pairs = H * (H + 1) / 2, entries = (uint64_t)u->du.S * pairs;
    uint32_t parts;
    if (warp_kernel) {
        parts = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, (pairs + 1) / 2));
    } else {
        const bool by_terms = (u->h_fill_cost[c] / std::max<uint64_t>(1, entries)) >= 16;
        parts = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, by_terms ? (entries + 3) / 4 : (entries + 31) / 32));
    }
    for (uint32_t p = 0; p < parts; p++) { tasks.push_back(sel_idx); tasks.push_back(p); tasks.push_back(parts); }
}

// dynamic shared memory of k_noise_chain: the hot state of each thread's first cluster (gibbs_core.cuh), as long as two blocks still fit an SM
static size_t noise_chain_hot_smem(uint32_t S, uint32_t bs, int *hot_out) {
    static const int hot_env = getenv("BTG_HOT") ? atoi(getenv("BTG_HOT")) : 1;
    const size_t bytes = (size_t)hot_bytes(S) * bs;
    const int hot = hot_env && bytes * 2 <= 220 * 1024;
    if (hot) cudaFuncSetAttribute(k_noise_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    *hot_out = hot;
    return hot ? bytes : 0;
}

static int noise_chains(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out, int joint);
static int estimate_noise_concurrent(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out,
                                     uint32_t chain_first = 0, uint32_t chain_stride = 1, double *chain_sums_out = nullptr);

int btg_estimate_noise(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, double *trace_out) {
    return estimate_noise_concurrent(u, cd, opts, nullptr, trace_out);
}

int btg_estimate_noise_sharded(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out) {
    return estimate_noise_concurrent(u, cd, opts, sh, trace_out);
}

int btg_estimate_noise_chains(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, uint32_t chain_first, uint32_t chain_stride, double *chain_sums_out,
                              double *trace_out) {
    if (!chain_sums_out || chain_stride == 0 || chain_first >= chain_stride) { set_error("bad chain partition (first %u, stride %u)", chain_first, chain_stride); return BTG_EINVAL; }
    return estimate_noise_concurrent(u, cd, opts, nullptr, trace_out, chain_first, chain_stride, chain_sums_out);
}

int btg_count_dist_finish_noise(btg_count_dist *cd, const double *chain_sums, uint32_t n_chains, uint32_t gibbs_samples) {
    BTG_REQUIRE_INIT();
    if (!cd || !chain_sums || n_chains == 0 || gibbs_samples == 0) { set_error("bad argument"); return BTG_EINVAL; }
    double *d = nullptr;
    const size_t bytes = (size_t)n_chains * cd->S * sizeof(double);
    if (btg::dmalloc(&d, bytes) != cudaSuccess) { set_error("allocation failed"); return BTG_ENOMEM; }
    auto s0 = ctx().stream;
    cudaMemcpyAsync(d, chain_sums, bytes, cudaMemcpyHostToDevice, s0);
    k_noise_finish<<<1, 256, 0, s0>>>(d, n_chains, cd->S, (double)gibbs_samples * n_chains, cd->rates, cd->noise, nullptr);
    BTG_LAUNCHED();
    const cudaError_t e = cudaStreamSynchronize(s0);
    btg::dfree(d);
    if (e != cudaSuccess) { set_error("noise finish failed: %s", cudaGetErrorString(e)); return BTG_ECUDA; }
    return BTG_OK;
}

int btg_estimate_noise_and_genotypes(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, btg_genotype_result *out, double *trace_out) {
    return btg_estimate_noise_and_genotypes_sharded(u, cd, opts, nullptr, out, trace_out);
}

int btg_estimate_noise_and_genotypes_sharded(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, btg_genotype_result *out,
                                             double *trace_out) {
    if (!out) { set_error("null argument"); return BTG_EINVAL; }
    int rc = noise_chains(u, cd, opts, sh, trace_out, 1);
    if (rc != BTG_OK) return rc;
    DevResult *dr = unit_result(u);
    if (!dr) { set_error("result allocation failed"); return BTG_ENOMEM; }
    if (u->du.C) {
        k_summarise<<<(u->du.C + 63) / 64, 64, 0, ctx().stream>>>(u->du, *opts, dr->R);
        BTG_LAUNCHED();
        BTG_CUDA(cudaGetLastError());
    }
    return btg_unit_download_result(u, out, nullptr);
}


// the engine's own stream (InferenceEngine.cpp:174) lives on the host: Fisher-Yates with the same Philox recipe as the device
struct HostEnginePhilox {
    uint32_t key[2], ctr[4], buf[4]; int pos;
    void init(uint32_t seed) { key[0] = seed; key[1] = 0; ctr[0] = ctr[1] = ctr[2] = 0; ctr[3] = kRngEngine; pos = 4; }
    uint32_t next() {
        if (pos == 4) {
            uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
            for (int r = 0; r < 10; r++) {
                const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
                c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
                k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
            }
            for (int i = 0; i < 4; i++) buf[i] = c[i];
            if (++ctr[0] == 0) ++ctr[1];
            pos = 0;
        }
        return buf[pos++];
    }
    uint32_t uniform_int(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};

// shadow arenas for chains that run at the same time: shadow 0 is the unit's own arena
static uint32_t ensure_shadows(btg_unit *u, uint32_t want) {
    if (u->shadow_du.empty()) u->shadow_du.push_back(u->du);
    while (u->shadow_du.size() < want) {
        double *f = nullptr; uint32_t *w = nullptr; uint8_t *b = nullptr, *t = nullptr;
        const bool ok = btg::dmalloc(&f, (u->f64_total + 1) * sizeof(double)) == cudaSuccess && btg::dmalloc(&w, (u->u32_total + 1) * sizeof(uint32_t)) == cudaSuccess &&
                        btg::dmalloc(&b, u->u8_total + 8) == cudaSuccess && btg::dmalloc(&t, u->tile_total + 32) == cudaSuccess;
        if (!ok) { btg::dfree(f); btg::dfree(w); btg::dfree(b); btg::dfree(t); cudaGetLastError(); break; }  // fewer concurrent chains
        for (void *p : {(void *)f, (void *)w, (void *)b, (void *)t}) u->allocs.push_back(p);
        DevUnit d = u->du;
        d.f64_pool = f; d.u32_pool = w; d.u8_pool = b; d.big_tile_pool = t;
        u->shadow_du.push_back(d);
    }
    return (uint32_t)std::min<size_t>(want, u->shadow_du.size());
}

// InferenceEngine::estimateNoise (InferenceEngine.cpp:135-276).  The chains are independent under this library's stream contract
// (fresh genotypers per chain as in the reference, InferenceEngine.cpp:240-251; the noise rates of chain c come from their own
// stream, kind 4 / chain c + 1, starting with the prior draw), so several chains run AT THE SAME TIME: each is one persistent
// cooperative k_noise_chain on its own CUDA stream, arena and noise state, with a share of the SMs (BTG_NOISE_CONCURRENCY = K).
// Measured (profiles/r1_noise_chain_phases.txt): K = 4 or 8 chains side by side take as long as one after the other — an iteration
// costs (clusters per thread) x (slowest lane's step), so a chain on 1/K of the SMs is K times slower — hence the default K = 1;
// the per-chain contract is what makes the results independent of K.
static int estimate_noise_concurrent(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out,
                                     uint32_t chain_first, uint32_t chain_stride, double *chain_sums_out) {
    BTG_REQUIRE_INIT();
    if (!u || !cd || !opts) { set_error("null argument"); return BTG_EINVAL; }
    btg_comm *comm = sh ? sh->comm : nullptr;
    const uint32_t world = comm ? comm->world : 1;
    if (sh && (!sh->group_n_clusters || !sh->group_n_variants || (u->du.G && group_index(*opts, u->du.G - 1) >= sh->n_groups_total))) {
        set_error("bad shard descriptor: this rank's groups (first %llu, stride %u, %u of them) do not fit the %llu groups of the unit", (unsigned long long)opts->group_index_base,
                  (unsigned)opts->group_index_stride, u->du.G, (unsigned long long)(sh ? sh->n_groups_total : 0));
        return BTG_EINVAL;
    }
    if (comm && !comm->connected) { set_error("communicator is not connected (btg_comm_connect)"); return BTG_ESTATE; }
    if (cd->S != u->du.S) { set_error("count distribution / unit sample mismatch"); return BTG_EINVAL; }
    const uint32_t S = u->du.S, G = u->du.G, n_chains = opts->n_chains;
    const uint32_t iters = (uint32_t)opts->gibbs_burn_in + opts->gibbs_samples;
    const size_t row_len = 2 + S, trace_rows = (size_t)n_chains * (iters + 1) + 1;
    auto s0 = ctx().stream;
    const bool want_phases = getenv("BTG_NOISE_PHASES") && atoi(getenv("BTG_NOISE_PHASES"));
    // one mailbox per communicator: a sharded unit runs its chains one after the other (the exchange is inside the kernel)
    uint32_t K = world > 1 || want_phases ? 1u : (uint32_t)std::max(1, getenv("BTG_NOISE_CONCURRENCY") ? atoi(getenv("BTG_NOISE_CONCURRENCY")) : 1);
    K = ensure_shadows(u, std::min(K, std::max(1u, n_chains)));
    int rc = BTG_OK;
    std::vector<void *> tmp;
    auto dalloc = [&](size_t bytes) { void *p = nullptr; if (btg::dmalloc(&p, bytes ? bytes : 8) != cudaSuccess) { rc = BTG_ENOMEM; return (void *)nullptr; } tmp.push_back(p); cudaMemsetAsync(p, 0, bytes ? bytes : 8, s0); return p; };
    // per-chain state, one allocation each: [n_chains][...]
    const size_t nc = std::max(1u, n_chains);
    auto *d_hist = (unsigned long long *)dalloc(nc * 2 * S * 8);
    auto *d_rates = (double *)dalloc(nc * S * 8);
    auto *d_tables = (double *)dalloc(nc * S * 256 * 8);
    auto *d_means = (double *)dalloc(nc * S * 8);
    auto *d_rng = (uint32_t *)dalloc(nc * 16 * 4);
    auto *d_rows = (uint32_t *)dalloc(nc * 4);
    auto *d_bar = (unsigned int *)dalloc(nc * 256);
    auto *d_trace = trace_out ? (double *)dalloc(trace_rows * row_len * 8) : nullptr;
    const uint32_t n_lg = 1024;
    auto *lg_tab = (double *)dalloc(n_lg * 8);
    unsigned long long *d_phase = want_phases ? (unsigned long long *)dalloc((8 + 3 * (size_t)iters + 16) * 8) : nullptr;
    std::vector<cudaStream_t> streams(K, nullptr);
    cudaEvent_t ev_ready = nullptr;
    std::vector<std::vector<uint32_t>> sels(n_chains), tasks(n_chains);
    std::vector<uint32_t> n_bigs(n_chains, 0);
    uint32_t *d_sel = nullptr, *d_tasks = nullptr;
    size_t sel_cap = 0, task_cap = 0;
    std::function<void(uint32_t)> select_chain;  // group selection of chain b (host; chains in order: the engine stream is sequential)
    if (rc == BTG_OK) {
        // ---- group selection of every chain (InferenceEngine.cpp:174-189), on the host, in chain order ----
        HostEnginePhilox engine;
        engine.init(opts->random_seed);
        const uint64_t base = sh ? opts->group_index_base : 0, stride = sh && opts->group_index_stride ? opts->group_index_stride : 1;
        std::vector<uint32_t> noise_groups;  // single-cluster groups of the WHOLE unit (InferenceEngine.cpp:144-151)
        if (sh) { for (uint64_t g = 0; g < sh->n_groups_total; g++) if (sh->group_n_clusters[g] == 1) noise_groups.push_back((uint32_t)g); }
        else { for (uint32_t g = 0; g < G; g++) if (u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] == 1) noise_groups.push_back(g); }
        auto group_variants = [&](uint32_t g) {
            if (sh) return sh->group_n_variants[g];
            uint64_t n = 0;
            for (uint64_t c = u->h_group_cluster_off[g]; c < u->h_group_cluster_off[g + 1]; c++) n += u->h_cl_var_off[c + 1] - u->h_cl_var_off[c];
            return (uint32_t)n;
        };
        const uint32_t noise_variants_batch_size = 100000;  // InferenceEngine.cpp:50
        auto is_big = [&](uint32_t c) { return lockstep_is_big(u, c, u->du.wide); };
        // upper bounds per chain: every local single-cluster group selected; every large cluster with its maximal number of fill tasks
        for (uint32_t g = 0; g < G; g++) {
            if (u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] != 1) continue;
            const uint32_t c = (uint32_t)u->h_group_cluster_off[g];
            sel_cap++;
            if (is_big(c)) task_cap += 3 * 64;
        }
        d_sel = (uint32_t *)dalloc(std::max<size_t>(1, sel_cap) * nc * 4);
        d_tasks = (uint32_t *)dalloc(std::max<size_t>(1, task_cap) * nc * 4);
        select_chain = [&, base, stride, noise_groups, engine, group_variants, is_big, noise_variants_batch_size](uint32_t b) mutable {
            uint32_t end = 0, nvv = 0;
            for (size_t i = noise_groups.size(); i > 1; i--) std::swap(noise_groups[i - 1], noise_groups[engine.uniform_int((uint32_t)i)]);
            while (nvv < noise_variants_batch_size && end < noise_groups.size()) { nvv += group_variants(noise_groups[end]); end++; }
            std::sort(noise_groups.begin(), noise_groups.begin() + end);
            auto &sel = sels[b];
            for (uint32_t i = 0; i < end; i++) {
                const uint64_t g = noise_groups[i];
                if (g >= base && (g - base) % stride == 0 && (g - base) / stride < G) sel.push_back((uint32_t)u->h_group_cluster_off[(g - base) / stride]);  // this rank's share
            }
            // large clusters first (fill tasks + one warp each), then by position in the cost order (neighbours share arena slots)
            std::sort(sel.begin(), sel.end(), [&](uint32_t a, uint32_t c) {
                const bool ba = is_big(a), bc = is_big(c);
                return ba != bc ? ba : u->h_layout[a].pos < u->h_layout[c].pos;
            });
            uint32_t n_big = 0;
            while (n_big < sel.size() && is_big(sel[n_big])) n_big++;
            n_bigs[b] = n_big;
            for (uint32_t i = 0; i < n_big; i++) lockstep_fill_tasks(u, sel[i], i, u->du.wide, tasks[b]);
        };
    }
    if (rc == BTG_OK) {
        for (auto &st : streams) if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) rc = BTG_ECUDA;
        if (cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming) != cudaSuccess) rc = BTG_ECUDA;
    }
    if (rc == BTG_OK) {
        k_lgamma_int<<<(n_lg + 127) / 128, 128, 0, s0>>>(lg_tab, n_lg);
        BTG_LAUNCHED();
        cudaEventRecord(ev_ready, s0);  // allocations zeroed, tables of cd complete
        PeerExchange px{};
        px.world = world; px.rank = comm ? comm->rank : 0;
        if (comm) {
            for (uint32_t r = 0; r < world; r++) px.mail[r] = comm->peers[r];
            px.error = comm->error;
            const char *tmo = getenv("BTG_PEER_TIMEOUT_MS");
            px.timeout_ns = (tmo ? strtoull(tmo, nullptr, 10) : 20000ull) * 1000000ull;
        }
        const uint32_t bs = 256;
        int per_sm = 0, hot = 0;
        const size_t hot_smem = noise_chain_hot_smem(S, bs, &hot);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_noise_chain, bs, hot_smem);
        const uint32_t capacity = (uint32_t)std::max(1, per_sm) * (uint32_t)ctx().sm_count;
        const uint32_t max_blocks = std::max(1u, (K > 1 ? capacity - capacity / 16 : capacity) / K);  // this chain's share of the SMs (a few block slots stay free)
        for (auto &st : streams) cudaStreamWaitEvent(st, ev_ready, 0);
        uint32_t launched = 0;
        for (uint32_t b = 0; b < n_chains && rc == BTG_OK; b++) {
            select_chain(b);  // while the previous chains run on the device (every chain's selection: the engine stream is sequential)
            if (b % chain_stride != chain_first) continue;     // a chain of another rank (btg_estimate_noise_chains)
            const uint32_t k = launched++ % K;
            cudaStream_t st = streams[k];
            NoiseState ns{};
            ns.hist = (uint64_t *)(d_hist + (size_t)b * 2 * S);
            ns.rates = d_rates + (size_t)b * S;
            ns.noise_table = d_tables + (size_t)b * S * 256;
            ns.mean_rates = d_means + (size_t)b * S;
            ns.trace = d_trace ? d_trace + (size_t)b * (iters + 1) * row_len : nullptr;
            ns.rng = d_rng + (size_t)b * 16;
            ns.trace_row = d_rows + b;
            ns.lg = lg_tab; ns.n_lg = n_lg;
            ns.phase_ns = d_phase;
            GridBarrier gb{d_bar + (size_t)b * 64, d_bar + (size_t)b * 64 + 32};
            uint32_t *sel_b = d_sel + (size_t)b * std::max<size_t>(1, sel_cap), *tasks_b = d_tasks + (size_t)b * std::max<size_t>(1, task_cap);
            if (!sels[b].empty() && cudaMemcpyAsync(sel_b, sels[b].data(), sels[b].size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = BTG_ECUDA; break; }
            if (!tasks[b].empty() && cudaMemcpyAsync(tasks_b, tasks[b].data(), tasks[b].size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = BTG_ECUDA; break; }
            // this chain's noise stream, and its first draw: the rates of the prior (CountDistribution.cpp:62,163-171)
            k_noise_rng_init<<<1, 1, 0, st>>>(ns.rng, opts->random_seed, b + 1);
            BTG_LAUNCHED();
            NoiseState ns_quiet = ns;
            ns_quiet.trace = nullptr;
            k_noise_update<<<1, 256, 0, st>>>(ns_quiet, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 0, 0, 0, 0, 1);
            BTG_LAUNCHED();
            Tables T{cd->genomic, ns.noise_table};
            uint32_t n_sel = (uint32_t)sels[b].size(), n_big = n_bigs[b], n_tasks = (uint32_t)(tasks[b].size() / 3), chain_id = b + 1, iters_arg = iters;
            const uint32_t want_threads = std::max<uint32_t>({n_big * 32u, n_sel - n_big, n_tasks * 32u, 1u});
            const uint32_t grid = std::max(1u, std::min((want_threads + bs - 1) / bs, max_blocks));
            float ps = cd->prior_shape, pc = cd->prior_scale;
            btg_gibbs_opts o = *opts;
            int joint = 0;
            unsigned long long *hist = (unsigned long long *)ns.hist;
            if (comm) { px.seq0 = comm->seq; comm->seq += iters; }
            DevUnit du_k = u->shadow_du[k];
            void *args[] = {&du_k, &T, &o, &sel_b, &n_sel, &n_big, &chain_id, &iters_arg, &ns, &ps, &pc, &hist, &joint, &px, &tasks_b, &n_tasks, &gb, &hot};
            cudaError_t e;
            if (u->du.wide) {  // warp per cluster, lane = sample (gibbs_wide.cu)
                e = wide_noise_chain(du_k, T, o, sel_b, n_sel, n_big, tasks_b, n_tasks, chain_id, iters_arg, ns, ps, pc, hist, joint, px, gb, K, ctx().sm_count, st);
            } else {
                e = cudaLaunchCooperativeKernel((void *)k_noise_chain, dim3(grid), dim3(bs), args, hot_smem, st);
                BTG_LAUNCHED();
            }
            if (e != cudaSuccess) { set_error("cooperative launch failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; break; }
        }
        // join the chain streams, then: mean of the post-burn-in rates -> setNoiseRates, final trace row "0 0"
        for (auto &st : streams) { cudaEvent_t ev; if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) { cudaEventRecord(ev, st); cudaStreamWaitEvent(s0, ev, 0); cudaEventDestroy(ev); } }
        if (rc == BTG_OK) {
            if (chain_sums_out) {   // this rank's chains only: the caller adds the ranks' rows up and ends with btg_count_dist_finish_noise
                cudaMemcpyAsync(chain_sums_out, d_means, (size_t)n_chains * S * 8, cudaMemcpyDeviceToHost, s0);
            } else {
                k_noise_finish<<<1, 256, 0, s0>>>(d_means, n_chains, S, (double)opts->gibbs_samples * n_chains, cd->rates, cd->noise,
                                                  d_trace ? d_trace + (trace_rows - 1) * row_len : nullptr);
                BTG_LAUNCHED();
            }
            if (trace_out) cudaMemcpyAsync(trace_out, d_trace, trace_rows * row_len * 8, cudaMemcpyDeviceToHost, s0);
        }
        cudaError_t e = cudaStreamSynchronize(s0);
        for (auto &st : streams) if (st) { cudaError_t e2 = cudaStreamSynchronize(st); if (e == cudaSuccess) e = e2; }
        if (e != cudaSuccess && rc == BTG_OK) { set_error("noise estimation failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; }
        if (rc == BTG_OK && comm && world > 1) {
            uint32_t flag = 0;
            cudaMemcpy(&flag, comm->error, 4, cudaMemcpyDeviceToHost);
            if (flag) { set_error("peer exchange timed out waiting for rank %u (a rank failed or left the lock-step)", flag - 1); cudaMemset(comm->error, 0, 4); rc = BTG_ECUDA; }
        }
        if (rc == BTG_OK) {
            std::vector<unsigned int> bar(nc * 64);
            cudaMemcpy(bar.data(), d_bar, bar.size() * 4, cudaMemcpyDeviceToHost);
            for (size_t b = 0; b < nc; b++) if (bar[b * 64 + 33]) { set_error("grid barrier of chain %zu timed out (grid not co-resident)", b); rc = BTG_ECUDA; break; }
        }
        if (rc == BTG_OK && d_phase) {
            unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            cudaMemcpy(ph, d_phase, sizeof ph, cudaMemcpyDeviceToHost);
            const double n_it = (double)n_chains * iters;
            fprintf(stderr, "[btgpu] noise chain phases, us per iteration (block 0, chains one after the other): fill+one-thread %.1f  large owners %.1f  exchange+update %.1f  release %.1f; construct+reset %.2f ms per chain\n",
                    ph[0] / n_it / 1e3, ph[1] / n_it / 1e3, ph[2] / n_it / 1e3, ph[3] / n_it / 1e3, ph[7] / 1e6 / std::max(1u, n_chains));
            const char *what[3] = {"fill task", "large-cluster owner", "one-thread cluster"};
            for (int k = 0; k < 3 && BTG_NOISE_TIMING; k++) {
                const uint32_t c = (uint32_t)(ph[4 + k] & 0xFFFFFFFFu);
                if (ph[4 + k] && c < u->du.C)
                    fprintf(stderr, "[btgpu]   slowest %s: %.1f us, cluster %u (H %u, variants %llu, fill cost %u)\n", what[k], (ph[4 + k] >> 32) / 1e3, c, u->h_nhap[c],
                            (unsigned long long)(u->h_cl_var_off[c + 1] - u->h_cl_var_off[c]), u->h_fill_cost[c]);
            }
            if (BTG_NOISE_TIMING) {   // per-iteration maxima: median over iterations
                std::vector<unsigned long long> it_max(3 * (size_t)iters);
                cudaMemcpy(it_max.data(), d_phase + 8, it_max.size() * 8, cudaMemcpyDeviceToHost);
                for (int k = 0; k < 3; k++) {
                    std::vector<double> us;
                    for (uint32_t i = 10; i < iters; i++) us.push_back((it_max[3 * i + k] >> 32) / 1e3);
                    if (us.empty()) continue;
                    std::sort(us.begin(), us.end());
                    fprintf(stderr, "[btgpu]   per-iteration slowest %s: median %.1f us, p90 %.1f us\n", what[k], us[us.size() / 2], us[us.size() * 9 / 10]);
                }
                unsigned long long sub[16];
                cudaMemcpy(sub, d_phase + 8 + 3 * (size_t)iters, sizeof sub, cudaMemcpyDeviceToHost);
                const char *nm[8] = {"bind", "fill rows", "rng load", "sample diplotypes", "sample frequencies", "rng save", "noise counts", "clear cache"};
                if (sub[8]) {
                    fprintf(stderr, "[btgpu]   one-thread clusters, mean cycles per cluster-iteration:");
                    for (int k = 0; k < 8; k++) fprintf(stderr, " %s %.0f;", nm[k], (double)sub[k] / (double)sub[8]);
                    fprintf(stderr, "\n");
                }
            }
        }
    } else if (rc != BTG_OK && !*btg_last_error()) {
        set_error("noise estimation allocation failed");
    }
    cudaStreamSynchronize(s0);
    for (auto &st : streams) if (st) cudaStreamDestroy(st);
    if (ev_ready) cudaEventDestroy(ev_ready);
    for (void *p : tmp) btg::dfree(p);
    return rc;
}

static int noise_chains(btg_unit *u, btg_count_dist *cd, const btg_gibbs_opts *opts, const btg_shard_desc *sh, double *trace_out, int joint) {
    BTG_REQUIRE_INIT();
    if (!u || !cd || !opts) { set_error("null argument"); return BTG_EINVAL; }
    btg_comm *comm = sh ? sh->comm : nullptr;
    const uint32_t world = comm ? comm->world : 1;
    if (sh && (!sh->group_n_clusters || !sh->group_n_variants || (u->du.G && group_index(*opts, u->du.G - 1) >= sh->n_groups_total))) {
        set_error("bad shard descriptor: this rank's groups (first %llu, stride %u, %u of them) do not fit the %llu groups of the unit", (unsigned long long)opts->group_index_base,
                  (unsigned)opts->group_index_stride, u->du.G, (unsigned long long)(sh ? sh->n_groups_total : 0));
        return BTG_EINVAL;
    }
    if (comm && !comm->connected) { set_error("communicator is not connected (btg_comm_connect)"); return BTG_ESTATE; }
    if (cd->S != u->du.S) { set_error("count distribution / unit sample mismatch"); return BTG_EINVAL; }
    const uint32_t S = u->du.S, G = u->du.G;
    // units with many samples, and any unit with nested groups in the joint mode, take the warp-per-group kernel (gibbs_wide.cu)
    const bool use_wide = u->du.wide || (joint && u->du.n_nested_groups);
    const uint32_t iters = (uint32_t)opts->gibbs_burn_in + opts->gibbs_samples;
    const size_t trace_rows = (size_t)opts->n_chains * (iters + 1) + (joint ? 0 : 1);
    auto s = ctx().stream;
    NoiseState ns{};
    unsigned long long *hist = nullptr;
    uint32_t *d_sel = nullptr, *d_tasks = nullptr;
    size_t tasks_cap = 0;
    std::vector<void *> tmp;
    auto dalloc = [&](size_t bytes) { void *p = nullptr; if (btg::dmalloc(&p, bytes ? bytes : 8) != cudaSuccess) return (void *)nullptr; tmp.push_back(p); cudaMemsetAsync(p, 0, bytes ? bytes : 8, s); return p; };
    hist = (unsigned long long *)dalloc((size_t)S * 2 * 8);
    ns.hist = (uint64_t *)hist;
    ns.rates = cd->rates;
    ns.noise_table = cd->noise;
    ns.mean_rates = (double *)dalloc(S * 8);
    ns.trace = trace_out ? (double *)dalloc(trace_rows * (2 + S) * 8) : nullptr;
    ns.rng = (uint32_t *)dalloc(16 * 4);
    GridBarrier gb{};
    gb.count = (unsigned int *)dalloc(256);   // count and generation in separate 128-byte lines
    gb.gen = gb.count ? gb.count + 32 : nullptr;
    const bool want_phases = getenv("BTG_NOISE_PHASES") && atoi(getenv("BTG_NOISE_PHASES"));
    ns.phase_ns = want_phases ? (unsigned long long *)dalloc((8 + 3 * (size_t)iters + 16) * 8) : nullptr;  // [4 phases, 3 maxima, spare][iteration][3 maxima]
    const uint32_t n_lg = 1024;
    double *lg_tab = (double *)dalloc(n_lg * 8);
    ns.lg = lg_tab; ns.n_lg = n_lg;
    ns.trace_row = (uint32_t *)dalloc(4);
    d_sel = (uint32_t *)dalloc((size_t)G * 4);
    int rc = BTG_OK;
    if (!hist || !ns.mean_rates || !ns.rng || !ns.trace_row || !d_sel || (trace_out && !ns.trace)) { set_error("noise estimation allocation failed"); rc = BTG_ENOMEM; }
    if (rc == BTG_OK) {
        // the engine's own stream (InferenceEngine.cpp:174) lives on the host: Fisher-Yates with the same Philox recipe
        struct HostPhilox {
            uint32_t key[2], ctr[4], buf[4]; int pos;
            void init(uint32_t seed) { key[0] = seed; key[1] = 0; ctr[0] = ctr[1] = ctr[2] = 0; ctr[3] = kRngEngine; pos = 4; }
            uint32_t next() {
                if (pos == 4) {
                    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
                    for (int r = 0; r < 10; r++) {
                        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
                        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
                        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
                        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
                    }
                    for (int i = 0; i < 4; i++) buf[i] = c[i];
                    if (++ctr[0] == 0) ++ctr[1];
                    pos = 0;
                }
                return buf[pos++];
            }
            uint32_t uniform_int(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
        } engine;
        engine.init(opts->random_seed);
        // single-cluster groups (InferenceEngine.cpp:144-151) of the WHOLE unit: with a shard descriptor every rank walks the
        // same global list with the same engine stream, so the selection does not depend on the sharding
        const uint64_t base = sh ? opts->group_index_base : 0, stride = sh && opts->group_index_stride ? opts->group_index_stride : 1;
        std::vector<uint32_t> noise_groups;
        // estimateNoise: groups of one cluster (InferenceEngine.cpp:144-151); estimateNoiseAndGenotypes: every group (:407-408)
        if (sh) {
            for (uint64_t g = 0; g < sh->n_groups_total; g++) if (joint || sh->group_n_clusters[g] == 1) noise_groups.push_back((uint32_t)g);
        } else {
            for (uint32_t g = 0; g < G; g++)
                if (joint || u->h_group_cluster_off[g + 1] - u->h_group_cluster_off[g] == 1) noise_groups.push_back(g);
        }
        if (joint && u->du.n_nested_groups) cudaMemsetAsync(u->du.shared_mult, 0, (size_t)u->n_shared * S, s);  // KmerCounts::multiplicities of a fresh run
        auto group_variants = [&](uint32_t g) {
            if (sh) return sh->group_n_variants[g];
            uint64_t n = 0;
            for (uint64_t c = u->h_group_cluster_off[g]; c < u->h_group_cluster_off[g + 1]; c++) n += u->h_cl_var_off[c + 1] - u->h_cl_var_off[c];
            return (uint32_t)n;
        };
        PeerExchange px{};
        px.world = world; px.rank = comm ? comm->rank : 0;
        if (comm) {
            for (uint32_t r = 0; r < world; r++) px.mail[r] = comm->peers[r];
            px.error = comm->error;
            const char *tmo = getenv("BTG_PEER_TIMEOUT_MS");
            px.timeout_ns = (tmo ? strtoull(tmo, nullptr, 10) : 20000ull) * 1000000ull;
        }
        Tables T{cd->genomic, cd->noise};
        k_noise_rng_init<<<1, 1, 0, s>>>(ns.rng, opts->random_seed);
        BTG_LAUNCHED();
        if (lg_tab) { k_lgamma_int<<<(n_lg + 127) / 128, 128, 0, s>>>(lg_tab, n_lg); BTG_LAUNCHED(); } else ns.lg = nullptr;
        NoiseState ns_quiet = ns;  // same state, no trace row
        ns_quiet.trace = nullptr;
        // CountDistribution ctor draws the initial rates (CountDistribution.cpp:62)
        k_noise_update<<<1, 256, 0, s>>>(ns_quiet, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 0, 0, 0, 0, 1);
        BTG_LAUNCHED();
        const uint32_t noise_variants_batch_size = 100000;  // InferenceEngine.cpp:50
        std::vector<uint32_t> sel, tasks;
        for (uint32_t chain = 0; chain < opts->n_chains && rc == BTG_OK; chain++) {
            uint32_t end = 0, nvv = 0;
            if (joint) {
                end = (uint32_t)noise_groups.size();  // every group, every chain (InferenceEngine.cpp:407-408)
            } else {
                for (size_t i = noise_groups.size(); i > 1; i--) std::swap(noise_groups[i - 1], noise_groups[engine.uniform_int((uint32_t)i)]);
                while (nvv < noise_variants_batch_size && end < noise_groups.size()) { nvv += group_variants(noise_groups[end]); end++; }
                std::sort(noise_groups.begin(), noise_groups.begin() + end);
            }
            sel.clear();
            for (uint32_t i = 0; i < end; i++) {
                const uint64_t g = noise_groups[i];
                if (g >= base && (g - base) % stride == 0 && (g - base) / stride < G) sel.push_back((uint32_t)u->h_group_cluster_off[(g - base) / stride]);  // this rank's share
            }
            // large clusters first (one warp each in the chain kernel), then by position in the cost order (neighbours share arena slots)
            auto is_big = [&](uint32_t c) { return lockstep_is_big(u, c, use_wide); };
            std::sort(sel.begin(), sel.end(), [&](uint32_t a, uint32_t b) {
                const bool ba = is_big(a), bb = is_big(b);
                return ba != bb ? ba : u->h_layout[a].pos < u->h_layout[b].pos;
            });
            uint32_t n_big = 0;
            while (n_big < sel.size() && is_big(sel[n_big])) n_big++;
            // fill tasks: a large cluster gets one warp per 32 cache entries (S x diplotypes, upper bound), at most 64
            tasks.clear();
            for (uint32_t i = 0; i < n_big; i++) lockstep_fill_tasks(u, sel[i], i, use_wide, tasks);
            if (tasks.size() > tasks_cap) {
                if (d_tasks) btg::dfree(d_tasks);
                tasks_cap = tasks.size() * 2;
                if (btg::dmalloc(&d_tasks, tasks_cap * 4) != cudaSuccess) { d_tasks = nullptr; tasks_cap = 0; rc = BTG_ENOMEM; set_error("fill task allocation failed"); break; }
            }
            if (!tasks.empty() && cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BTG_ECUDA; break; }
            if (cudaMemcpyAsync(d_sel, sel.data(), sel.size() * 4, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = BTG_ECUDA; break; }
            cudaStreamSynchronize(s);  // sel is reused by the host next chain
            const uint32_t n_sel = (uint32_t)sel.size();
            if (n_sel || world > 1) {  // a rank with nothing selected still takes part in every exchange
                // (a shared-memory window of the log-pmf tables was tried and made the fill slower: the chain is bound by instruction issue
                //  at 16 warps/SM, not by the gathers — profiles/r1_noise_chain_phases.txt)
                const uint32_t bs = 256;
                int per_sm = 0, hot = 0;
                const size_t smem = use_wide || joint ? 0 : noise_chain_hot_smem(S, bs, &hot);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_noise_chain, bs, smem);
                const uint32_t max_blocks = (uint32_t)std::max(1, per_sm) * (uint32_t)ctx().sm_count;
                const uint32_t want_threads = std::max<uint32_t>({n_big * 32u, n_sel - n_big, (uint32_t)(tasks.size() / 3) * 32u});
                const uint32_t grid = std::max(1u, std::min((want_threads + bs - 1) / bs, max_blocks));
                uint32_t chain_id = chain + 1, n_sel_arg = n_sel, iters_arg = iters;
                float ps = cd->prior_shape, pc = cd->prior_scale;
                btg_gibbs_opts o = *opts;
                if (comm) { px.seq0 = comm->seq; comm->seq += iters; }
                uint32_t n_tasks = (uint32_t)(tasks.size() / 3);
                void *args[] = {&u->du, &T, &o, &d_sel, &n_sel_arg, &n_big, &chain_id, &iters_arg, &ns, &ps, &pc, &hist, &joint, &px, &d_tasks, &n_tasks, &gb, &hot};
                cudaError_t e;
                if (use_wide) {
                    e = wide_noise_chain(u->du, T, o, d_sel, n_sel_arg, n_big, d_tasks, n_tasks, chain_id, iters_arg, ns, ps, pc, hist, joint, px, gb, 1, ctx().sm_count, s);
                } else {
                    e = cudaLaunchCooperativeKernel((void *)k_noise_chain, dim3(grid), dim3(bs), args, smem, s);
                    BTG_LAUNCHED();
                }
                if (e != cudaSuccess) { set_error("cooperative launch failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; break; }
            } else {
                if (ns.trace) { k_noise_update<<<1, 256, 0, s>>>(ns, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 3, 0, chain + 1, 0, 1); BTG_LAUNCHED(); }
                for (uint32_t it = 1; it <= iters; it++) {
                    k_noise_update<<<1, 256, 0, s>>>(ns, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 1, opts->gibbs_burn_in < it, chain + 1, it, 1);
                    BTG_LAUNCHED();
                }
            }
            // resetNoiseRates at the end of the chain (InferenceEngine.cpp:253); not a trace row
            k_noise_update<<<1, 256, 0, s>>>(ns_quiet, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 0, 0, 0, 0, 1);
            BTG_LAUNCHED();
            if (cudaGetLastError() != cudaSuccess) rc = BTG_ECUDA;
        }
        if (rc == BTG_OK) {
            // mean of the post-burn-in rates -> setNoiseRates (InferenceEngine.cpp:259-264); final trace row "0 0"
            if (!joint) k_noise_update<<<1, 256, 0, s>>>(ns, S, cd->prior_shape, cd->prior_scale, opts->random_seed, 2, 0, 0, 0, (double)opts->gibbs_samples * opts->n_chains);
            BTG_LAUNCHED();
            if (trace_out) cudaMemcpyAsync(trace_out, ns.trace, trace_rows * (2 + S) * 8, cudaMemcpyDeviceToHost, s);
            cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { set_error("noise estimation failed: %s", cudaGetErrorString(e)); rc = BTG_ECUDA; }
            if (rc == BTG_OK && ns.phase_ns) {
                unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                cudaMemcpy(ph, ns.phase_ns, sizeof ph, cudaMemcpyDeviceToHost);
                if (BTG_NOISE_TIMING) {
                    unsigned long long sub[16];
                    cudaMemcpy(sub, ns.phase_ns + 8 + 3 * (size_t)iters, sizeof sub, cudaMemcpyDeviceToHost);
                    const char *nm[8] = {"bind", "fill rows", "rng load", "sample diplotypes", "sample frequencies", "rng save", "noise counts", "clear cache"};
                    if (sub[8]) {
                        fprintf(stderr, "[btgpu]   one-thread clusters, mean cycles per cluster-iteration:");
                        for (int k = 0; k < 8; k++) fprintf(stderr, " %s %.0f;", nm[k], (double)sub[k] / (double)sub[8]);
                        fprintf(stderr, "\n");
                    }
                }
                const char *what[3] = {"fill task", "large-cluster owner", "one-thread cluster"};
                if (BTG_NOISE_TIMING) {   // per-iteration maxima (accumulated over the chains by atomicMax): median over iterations and the cluster that is most often the slowest
                    std::vector<unsigned long long> it_max(3 * (size_t)iters);
                    cudaMemcpy(it_max.data(), ns.phase_ns + 8, it_max.size() * 8, cudaMemcpyDeviceToHost);
                    for (int k = 0; k < 3; k++) {
                        std::vector<double> us;
                        std::vector<uint32_t> who;
                        for (uint32_t i = 10; i < iters; i++) { us.push_back((it_max[3 * i + k] >> 32) / 1e3); who.push_back((uint32_t)(it_max[3 * i + k] & 0xFFFFFFFFu)); }
                        if (us.empty()) continue;
                        std::vector<double> sorted_us = us;
                        std::sort(sorted_us.begin(), sorted_us.end());
                        std::sort(who.begin(), who.end());
                        uint32_t best = who[0], best_n = 0, run = 0;
                        for (size_t i = 0; i < who.size(); i++) { run = (i && who[i] == who[i - 1]) ? run + 1 : 1; if (run > best_n) { best_n = run; best = who[i]; } }
                        fprintf(stderr, "[btgpu]   per-iteration slowest %s: median %.1f us, p90 %.1f us; most often cluster %u (%u of %zu iterations; H %u, variants %llu, fill cost %u)\n",
                                what[k], sorted_us[sorted_us.size() / 2], sorted_us[sorted_us.size() * 9 / 10], best, best_n, who.size(), best < u->du.C ? u->h_nhap[best] : 0,
                                best < u->du.C ? (unsigned long long)(u->h_cl_var_off[best + 1] - u->h_cl_var_off[best]) : 0ull, best < u->du.C ? u->h_fill_cost[best] : 0);
                    }
                }
                for (int k = 0; k < 3 && (BTG_NOISE_TIMING || use_wide); k++) {
                    const uint32_t c = (uint32_t)(ph[4 + k] & 0xFFFFFFFFu);
                    if (ph[4 + k] && c < u->du.C)
                        fprintf(stderr, "[btgpu]   slowest %s: %.1f us, cluster %u (H %u, variants %llu, fill cost %u)\n", what[k], (ph[4 + k] >> 32) / 1e3, c, u->h_nhap[c],
                                (unsigned long long)(u->h_cl_var_off[c + 1] - u->h_cl_var_off[c]), u->h_fill_cost[c]);
                }
                const double n_it = (double)opts->n_chains * iters;
                fprintf(stderr, "[btgpu] noise chain phases, us per iteration (block 0): fill %.1f  sample %.1f  exchange+update %.1f  release %.1f\n",
                        ph[0] / n_it / 1e3, ph[1] / n_it / 1e3, ph[2] / n_it / 1e3, ph[3] / n_it / 1e3);
            }
            if (rc == BTG_OK && comm && world > 1) {
                uint32_t flag = 0;
                cudaMemcpy(&flag, comm->error, 4, cudaMemcpyDeviceToHost);
                if (flag) { set_error("peer exchange timed out waiting for rank %u (a rank failed or left the lock-step)", flag - 1); cudaMemset(comm->error, 0, 4); rc = BTG_ECUDA; }
            }
        } else {
            set_error("noise estimation failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    cudaStreamSynchronize(s);
    for (void *p : tmp) btg::dfree(p);
    if (d_tasks) btg::dfree(d_tasks);
    return rc;
}

}  // extern "C"
