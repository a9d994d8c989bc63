// gibbs_wide.cu — the Gibbs sampler with ONE WARP PER CLUSTER and ONE LANE PER SAMPLE (sm_100a).
//
// gibbs.cu gives every cluster (or nested group) one thread, which walks the samples of an iteration in turn — the right shape
// for 1..3 samples and ~10^5 clusters per launch.  With 30 samples (BASELINE configs[3], --noise-genotyping) that walk is the
// whole cost of a lock-step iteration (7-15 ms, profiles/r1_joint_30samples.txt).  The model allows more: inside one
// sampleDiplotypes call the haplotype frequencies are fixed, so the S diplotype draws are conditionally independent
// (VariantClusterGenotyper.cpp:668-705 — the loop over samples only adds up haplotype counts; the multicluster multiplicities it
// updates are per-sample records, KmerCounts.cpp:205-224).  Here lane s of the cluster's warp
//   * draws the diplotype of sample s from the counter block that (call, sample) owns in the genotyper's stream (gibbs_rng.cuh),
//   * books its tally, its k-mer statistics and its share of the noise counts (all per-sample state),
// the haplotype counts are added up with atomics, and lane 0 draws the haplotype frequencies (a sequential stream).
// The arena of such a unit is "wide": one dense slot per cluster, so the lanes of a warp read neighbouring bytes of the k-mer
// tile (counts of the 30 samples of one k-mer = one 32-byte sector).
//
// Groups with nested clusters run on one warp per GROUP, which walks the group's clusters in the reference's depth-first order
// (VariantClusterGroup::runGibbsSample, VariantClusterGroup.cpp:236-250) — also in the lock-step modes, so
// InferenceEngine::estimateNoiseAndGenotypes (InferenceEngine.cpp:384-472) accepts nested groups.
//
// Every function here works on either arena layout (the accessors carry the stride); the layout only decides what is coalesced.
#include "gibbs_core.cuh"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;

// VariantClusterGenotyper::reset + VariantClusterHaplotypes::sampleKmerSubset by a warp: lane 0 takes the sequential draws
// (shuffle, Bernoulli per k-mer), all lanes build the k-mer tile and clear the caches
template <bool MC>
__device__ __forceinline__ void clw_reset(Cl &cl, const btg_gibbs_opts &o, Philox &prng, uint32_t lane, bool fresh) {
    const uint32_t H = cl.H, S = cl.S;
    if (fresh) {  // chains are independent in the default mode: every chain shuffles the original order
        const uint32_t *src = cl.u->uniq_idx + cl.u->cl_uniq_off[cl.c];
        for (uint32_t i = lane; i < cl.n_uniq; i += 32) cl.uniq[i] = src[i];
    }
    for (uint32_t i = lane; i < H * cl.nvar; i += 32) cl.cnt[i] = 0;
    __syncwarp();
    if (lane == 0) {
        const double rate = (double)o.kmer_subsampling_rate;
        for (uint32_t i = cl.n_uniq; i > 1; i--) {  // Fisher-Yates from the back
            const uint32_t j = prng.uniform_int(i);
            const uint32_t t = cl.uniq[i - 1]; cl.uniq[i - 1] = cl.uniq[j]; cl.uniq[j] = t;
        }
        uint32_t n_sub = 0;
        for (uint32_t i = 0; i < cl.n_uniq; i++) {
            const uint32_t k = cl.uniq[i];
            if (prng.u01() < rate)
                if (!cl_is_max_hap_var_kmer(cl, k, o.max_haplotype_variant_kmers)) cl.uniq_sub[n_sub++] = k;
        }
        cl.misc[kNSub] = n_sub;
        if constexpr (MC) {
            if (cl.n_multi) {
                for (uint32_t i = cl.n_multi; i > 1; i--) {
                    const uint32_t j = prng.uniform_int(i);
                    const uint32_t t = cl.multi[i - 1]; cl.multi[i - 1] = cl.multi[j]; cl.multi[j] = t;
                }
                uint32_t n_msub = 0;
                for (uint32_t i = 0; i < cl.n_multi; i++) {
                    const uint32_t k = cl.multi[i];
                    if (prng.u01() < rate)
                        if (!cl_is_max_hap_var_kmer(cl, k, o.max_haplotype_variant_kmers)) cl.multi_sub[n_msub++] = k;
                }
                cl.misc[kNMultiSub] = n_msub;
            }
            cl.misc[kUseMulti] = 0;
        }
    }
    __syncwarp();
    const uint32_t n_sub = cl.misc[kNSub];
    for (uint32_t e = lane; e < n_sub * H; e += 32) { const uint32_t i = e / H, h = e - i * H; cl.tile_m[e] = cl.m(cl.uniq_sub[i], h); }
    for (uint32_t e = lane; e < n_sub * S; e += 32) {
        const uint32_t i = e / S, s = e - i * S, k = cl.uniq_sub[i];
        cl.tile_c[e] = cl.u->k_has_counts[cl.row0 + k] ? cl.u->k_counts[(cl.row0 + k) * S + s] : 0;
    }
    for (uint32_t e = lane; e < n_sub * 2; e += 32) {
        const uint32_t k = cl.uniq_sub[e >> 1];
        cl.tile_ic[e] = cl.u->k_has_counts[cl.row0 + k] ? cl.u->k_ic[(cl.row0 + k) * 2 + (e & 1u)] : 0;
    }
    for (uint32_t s = lane; s < S; s += 32) cl.stats_update[s] = 1;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (cl.has_cache)
        for (uint32_t i = lane; i < S * cl.Dall; i += 32) cl.ucache[i] = nan;
    if constexpr (MC) {
        if (cl.n_multi) {
            const uint32_t n_msub = cl.misc[kNMultiSub];
            for (uint32_t i = lane; i < n_msub * S; i += 32) cl.sample_multi[i] = 0;
            for (uint32_t i = lane; i < S * cl.Dall; i += 32) cl.mcache[i] = nan;
        }
    }
    const double f0 = 1 / static_cast<double>(H);
    for (uint32_t h = lane; h < H; h += 32) { cl.obs[h] = 0; cl.freq[h] = f0; cl.nz[h] = 1; }
    __syncwarp();
}

// VariantClusterGenotyper::sampleDiplotypes (VariantClusterGenotyper.cpp:668-705), lane = sample.  `prng` is the genotyper's
// stream (every lane holds a copy: the draws of this call are addressed by (t_draw, sample), see gibbs_rng.cuh).
template <bool MC>
__device__ __forceinline__ void clw_sample_diplotypes(Cl &cl, const Tables &T, const uint8_t *ploidy, bool collect, Philox &prng, uint32_t lane) {
    const uint32_t H = cl.H, S = cl.S;
    uint64_t nzm = 0;  // non-zero flags of the first 64 haplotypes as a bit mask
    for (uint32_t hb = 0; hb < H; hb += 32) {
        const uint32_t h = hb + lane;
        const bool live = h < H && cl.nz[h];
        if (live) cl.logf[h] = m_log(cl.freq[h]);  // one log per haplotype per iteration
        const uint32_t bal = __ballot_sync(FULL, live);
        if (hb < 64) nzm |= (uint64_t)bal << hb;
    }
    __syncwarp();
    const uint32_t s = lane;
    const bool act = s < S;
    uint32_t da = NONE, db = NONE;
    if (act) {
        const uint32_t prev = cl.dipl[s];
        if constexpr (MC) { if (cl.misc[kUseMulti]) cl_update_multi_log_prob(cl, T, s); }
        cl_sample_diplotype<MC, true, true>(cl, T, s, ploidy[s], prng, nzm);
        if constexpr (MC) cl_update_multi_multiplicities(cl, s, prev);
        else if (cl.dipl[s] != prev) cl.stats_update[s] = 1;  // …Haplotypes.cpp:199-201
        da = cl.dipl[s] & 0xFFFFu; db = cl.dipl[s] >> 16;
        if (collect) cl.tally[(size_t)cl.slot(da == NONE ? H : da, db == NONE ? H : db) * S + s]++;
        // HaplotypeFrequencyDistribution::incrementCount (HaplotypeFrequencyDistribution.cpp:114-126): counts are sums over samples
        if (da != NONE) atomicAdd(&cl.obs[da], 1u);
        if (db != NONE) atomicAdd(&cl.obs[db], 1u);
    }
    const uint32_t n_hap = __popc(__ballot_sync(FULL, act && da != NONE)) + __popc(__ballot_sync(FULL, act && db != NONE));
    const uint32_t n_missing = __popc(__ballot_sync(FULL, act && da == NONE)) + __popc(__ballot_sync(FULL, act && db == NONE));
    if (lane == 0) { cl.misc[kNumHap] += n_hap; cl.misc[kNumMissing] += n_missing; }
    prng.t_draw++;
    __syncwarp();
    if (collect && act) cl_update_allele_stats_sample<MC>(cl, s);
    if constexpr (MC) { if (lane == 0) cl.misc[kUseMulti] = cl.misc[kNMultiSub] > 0; }
    __syncwarp();
}

// VariantClusterGenotyper::getNoiseCounts (VariantClusterGenotyper.cpp:757-779): lane = sample; only (n_obs, sum) per sample are
// ever read from the merged CountAllocation (CountDistribution.cpp:188-200)
__device__ __forceinline__ void clw_noise_counts(const Cl &cl, unsigned long long *sh_stat, uint32_t lane) {
    const uint32_t s = lane;
    if (s >= cl.S) return;
    const uint32_t n_sub = cl.misc[kNSub];
    const uint32_t da = cl.dipl[s] & 0xFFFFu, db = cl.dipl[s] >> 16, g = cl.u->sample_gender[s];
    uint32_t n0 = 0, c0 = 0;
    for (uint32_t j = 0; j < n_sub; j++) {
        const uint8_t mm = (uint8_t)(cl.tileDiplMult(j, da, db) + cl.tile_ic[j * 2 + g]), cc = cl.tile_c[j * cl.S + s];
        n0 += mm == 0; c0 += mm == 0 ? cc : 0u;
    }
    if (n0) { atomicAdd(sh_stat + 2 * s, (unsigned long long)n0); atomicAdd(sh_stat + 2 * s + 1, (unsigned long long)c0); }
}

// One fill task of a large cluster (lock-step modes, where the caches are cleared every iteration): the task takes every parts-th
// pair of live haplotypes and lane s sums the k-mer tile for sample s — the sums that calcDiplotypeLogProb
// (VariantClusterGenotyper.cpp:597-666) would take on first use, in the same order, so the cached values are the same bits.
__device__ __forceinline__ void clw_fill_cache(Cl &cl, const Tables &T, const uint8_t *ploidy, uint32_t lane, uint32_t part, uint32_t parts) {
    const uint32_t H = cl.H, s = lane, n_sub = cl.misc[kNSub];
    const uint8_t pl = s < cl.S ? ploidy[s] : 0;
    uint32_t e = 0;
    for (uint32_t a = 0; a < H; a++) {
        if (!cl.nz[a]) continue;
        for (uint32_t b = a; b < H; b++) {
            if (!cl.nz[b]) continue;
            const uint32_t mine = e++;
            if (mine % parts != part) continue;
            if (pl == 2) cl.ucache[(size_t)s * cl.Dall + cl.slot(a, b)] = tile_entry_sum(cl, T, s, a, b, n_sub);
            else if (pl == 1 && a == b) cl.ucache[(size_t)s * cl.Dall + cl.slot(a, H)] = tile_entry_sum(cl, T, s, a, NONE, n_sub);
        }
    }
}

// VariantClusterGenotyper::clearCache (…Genotyper.cpp:131-138)
template <bool MC>
__device__ __forceinline__ void clw_clear_cache(Cl &cl, uint32_t lane) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (cl.has_cache)
        for (uint32_t j = lane; j < cl.S * cl.Dall; j += 32) cl.ucache[j] = nan;
    if constexpr (MC) {
        if (cl.n_multi)
            for (uint32_t j = lane; j < cl.S * cl.Dall; j += 32) cl.mcache[j] = nan;
    }
}

// per-sample halves of cl_update_nested_info / cl_add_nested_stats and of the info a child inherits (gibbs_core.cuh)
__device__ __forceinline__ void clw_pass_nested_info(Cl &cl, const DevUnit &du, uint64_t c0, uint32_t v, uint32_t ns, uint32_t lane) {
    const uint32_t S = cl.S, s = lane;
    for (uint64_t e = du.cl_edge_off[c0 + v]; e < du.cl_edge_off[c0 + v + 1]; e++) {
        const uint32_t t = du.edge_mut[e], nt = du.nest_slot[c0 + t];
        if (s < S) {
            du.nest_pl[(size_t)nt * S + s] = du.nest_pl[(size_t)ns * S + s];
            const uint32_t nk = du.nest_k[(size_t)ns * S + s];
            du.nest_k[(size_t)nt * S + s] = (uint8_t)nk;
            for (uint32_t k = 0; k < nk; k++) {
                const size_t a = ((size_t)ns * S + s) * 2 + k, b = ((size_t)nt * S + s) * 2 + k;
                du.nest_n[b] = du.nest_n[a]; du.nest_f[2 * b] = du.nest_f[2 * a]; du.nest_f[2 * b + 1] = du.nest_f[2 * a + 1];
            }
            cl_update_nested_info(cl, nt, du.cluster_idx[c0 + t], s, s + 1);
        }
        __syncwarp();
    }
}

// VariantClusterGroup::shuffleBranchOrdering (cumulative, VariantClusterGroup.cpp:208-218) and the depth-first pre-order that
// runGibbsSample's recursion visits; one lane
__device__ __forceinline__ void group_branch_order(const DevUnit &du, const btg_gibbs_opts &o, uint32_t g, uint32_t chain) {
    const uint64_t c0 = du.group_cluster_off[g];
    const uint32_t n = (uint32_t)(du.group_cluster_off[g + 1] - c0);
    const uint64_t s0 = du.group_src_off[g], s1 = du.group_src_off[g + 1];
    Philox br;
    br.init(o.random_seed, group_index(o, g), 0, kRngBranch, chain);
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
    uint32_t top = 0, len = 0;
    for (uint64_t e = s1; e > s0; e--) du.dfs_stack[c0 + top++] = du.src_mut[e - 1];
    while (top > 0) {
        const uint32_t v = du.dfs_stack[c0 + --top];
        du.dfs_order[c0 + len++] = v;
        for (uint64_t e = du.cl_edge_off[c0 + v + 1]; e > du.cl_edge_off[c0 + v]; e--) du.dfs_stack[c0 + top++] = du.edge_mut[e - 1];
    }
    const uint8_t *ploidy = du.group_ploidy + (size_t)g * du.S;
    for (uint64_t e = s0; e < s1; e++) {  // the info a source vertex receives: the chromosome ploidy, no enclosing allele
        const uint32_t ns = du.nest_slot[c0 + du.src_mut[e]];
        for (uint32_t s = 0; s < du.S; s++) { du.nest_pl[(size_t)ns * du.S + s] = ploidy[s]; du.nest_k[(size_t)ns * du.S + s] = 0; }
    }
}

// One Gibbs iteration of one group (VariantClusterGroup::estimateGenotypes): its clusters in depth-first order; with sh_stat also
// getNoiseCounts + clearGenotyperCache of the lock-step modes (sampleGenotypesCallback, InferenceEngine.cpp:77-98).
template <bool MC>
__device__ __forceinline__ void group_iteration(const DevUnit &du, const Tables &T, const btg_gibbs_opts &o, uint32_t g, bool collect, unsigned long long *sh_stat,
                                                uint32_t lane) {
    const uint64_t c0 = du.group_cluster_off[g];
    const uint32_t n = MC ? (uint32_t)(du.group_cluster_off[g + 1] - c0) : 1u, S = du.S;
    const uint64_t gidx = group_index(o, g);
    for (uint32_t pos = 0; pos < n; pos++) {
        const uint32_t v = MC ? du.dfs_order[c0 + pos] : 0u;
        Cl cl;
        cl.bind(du, (uint32_t)(c0 + v));
        const uint32_t ns = MC ? du.nest_slot[c0 + v] : 0u;
        const uint8_t *ploidy = MC ? du.nest_pl + (size_t)ns * S : du.group_ploidy + (size_t)g * S;
        Philox prng, fr;
        prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
        clw_sample_diplotypes<MC>(cl, T, ploidy, collect, prng, lane);
        if constexpr (MC) {
            if (collect && lane < S) cl_add_nested_stats(cl, ns, lane, lane + 1);
        }
        if (lane == 0) {
            fr.load(cl.rng, kRng1, o.random_seed, gidx, du.cluster_idx[cl.c]);
            cl_sample_frequencies(cl, fr);
            fr.save(cl.rng, kRng1);
            cl.rng[kRng0 + 8] = prng.t_draw;  // the sequential part of the genotyper's stream is untouched by an iteration
        }
        __syncwarp();
        if constexpr (MC) clw_pass_nested_info(cl, du, c0, v, ns, lane);
        if (sh_stat) {
            clw_noise_counts(cl, sh_stat, lane);
            clw_clear_cache<MC>(cl, lane);
        }
        __syncwarp();
    }
}

// InferenceEngine::estimateGenotypesCallback (InferenceEngine.cpp:278-333), default mode, single-cluster groups: one warp per
// cluster (or per chain-split position of a large cluster), all chains
__global__ void __launch_bounds__(128) k_estimate_genotypes_wide(DevUnit du, Tables T, btg_gibbs_opts o, ResultView R) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const uint32_t n_virtual = du.n_split * kChainSplit;
    uint32_t cluster, pos = NONE32, chain0 = 0, chain_step = 1;
    bool split = false;
    if (w < n_virtual) {
        const uint32_t j = w / kChainSplit, v = w % kChainSplit;
        cluster = du.split_cluster[j];
        pos = du.split_pos[(size_t)j * kChainSplit + v];
        chain0 = v; chain_step = kChainSplit;
        split = true;
    } else {
        const uint32_t i = w - n_virtual;
        if (i >= du.n_regular) return;
        cluster = du.order[i];
        if (du.split_of[cluster] != NONE32) return;  // handled above
    }
    Cl cl;
    cl.bind(du, cluster, pos);
    const uint64_t gidx = group_index(o, cl.g);
    const uint8_t *ploidy = du.group_ploidy + (size_t)cl.g * du.S;
    cl_construct_warp(cl, o, gidx, 0, lane);
    const uint32_t iters = (uint32_t)o.gibbs_burn_in + o.gibbs_samples;
    for (uint32_t chain = chain0; chain < o.n_chains; chain += chain_step) {
        Philox prng, fr;
        prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, chain);
        fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, chain);
        clw_reset<false>(cl, o, prng, lane, true);
        for (uint32_t it = 0; it < iters; it++) {
            clw_sample_diplotypes<false>(cl, T, ploidy, it >= o.gibbs_burn_in, prng, lane);
            if (lane == 0) cl_sample_frequencies(cl, fr);
            __syncwarp();
        }
    }
    if (!split) cl_summarise(cl, o, ploidy, R, lane, 32);
}

// the same for a group with several clusters: one warp = one GROUP (k_estimate_genotypes_nested of gibbs.cu, lane = sample)
__global__ void __launch_bounds__(128) k_estimate_genotypes_nested_wide(DevUnit du, Tables T, btg_gibbs_opts o, ResultView R) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (w >= du.n_nested_groups) return;
    const uint32_t g = du.nested_groups[w];
    const uint64_t c0 = du.group_cluster_off[g];
    const uint32_t n = (uint32_t)(du.group_cluster_off[g + 1] - c0);
    const uint64_t gidx = group_index(o, g);
    if (lane == 0) {
        for (uint64_t e = du.group_src_off[g]; e < du.group_src_off[g + 1]; e++) du.src_mut[e] = du.group_src[e];
        for (uint64_t e = du.cl_edge_off[c0]; e < du.cl_edge_off[c0 + n]; e++) du.edge_mut[e] = du.edge_dst[e];
    }
    for (uint32_t j = 0; j < n; j++) {  // VariantClusterGroup::initGenotyper: genotypers are constructed once
        Cl cl;
        cl.bind(du, (uint32_t)(c0 + j));
        cl_construct_warp(cl, o, gidx, 0, lane);
        if (lane == 0) {
            Philox prng, fr;
            prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, 0);
            fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, 0);
            prng.save(cl.rng, kRng0);
            fr.save(cl.rng, kRng1);
        }
        __syncwarp();
    }
    const uint32_t iters = (uint32_t)o.gibbs_burn_in + o.gibbs_samples;
    for (uint32_t chain = 0; chain < o.n_chains; chain++) {
        for (uint32_t j = 0; j < n; j++) {
            Cl cl;
            cl.bind(du, (uint32_t)(c0 + j));
            Philox prng;
            prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
            clw_reset<true>(cl, o, prng, lane, false);
            if (lane == 0) prng.save(cl.rng, kRng0);
            __syncwarp();
        }
        if (lane == 0) group_branch_order(du, o, g, chain);
        __syncwarp();
        for (uint32_t it = 0; it < iters; it++) group_iteration<true>(du, T, o, g, it >= o.gibbs_burn_in, nullptr, lane);
    }
    const uint8_t *ploidy = du.group_ploidy + (size_t)g * du.S;
    for (uint32_t j = 0; j < n; j++) {  // collectGenotypes: every cluster is summarised with the chromosome ploidy
        Cl cl;
        cl.bind(du, (uint32_t)(c0 + j));
        cl_summarise(cl, o, ploidy, R, lane, 32);
    }
}

// One chain of a lock-step mode as one persistent cooperative kernel (k_noise_chain of gibbs.cu for the warp-per-group shape).
// sel[i] = first cluster of the i-th selected group.  joint = 0: estimateNoise (fresh genotypers each chain, streams of chain
// `chain`); joint = 1: estimateNoiseAndGenotypes — genotypers constructed in the first chain only (streams of chain 0), samples
// collected after the burn-in, groups may hold nested clusters.
// sel[0 .. n_big): large single-cluster groups — their caches are filled by the whole grid (fill_tasks: (index in sel, part, parts)) while the
// other groups take their step, and they sample from the filled caches after a grid barrier.
__global__ void __launch_bounds__(256, 2) k_noise_chain_wide(DevUnit du, Tables T, btg_gibbs_opts o, const uint32_t *sel, uint32_t n_sel, uint32_t n_big,
                                                            const uint32_t *fill_tasks, uint32_t n_fill_tasks, uint32_t chain, uint32_t iters,
                                                            NoiseState ns, float prior_shape, float prior_scale, unsigned long long *hist, int joint, PeerExchange px,
                                                            GridBarrier gb) {
    __shared__ unsigned long long sh_tot[kMailRow];
    __shared__ double sh_rates[BTG_MAX_SAMPLES];
    __shared__ unsigned long long sh_stat[BTG_MAX_SAMPLES * 2];
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u, n_warps = (gridDim.x * blockDim.x) >> 5;
    const bool first = !joint || chain == 1;
    const uint32_t stream_chain = joint ? 0 : chain;
    const unsigned long long t_start = ns.phase_ns && blockIdx.x == 0 && threadIdx.x == 0 ? global_timer_ns() : 0;
    for (uint32_t i = w; i < n_sel; i += n_warps) {  // initGenotypersCallback (InferenceEngine.cpp:60-75)
        const uint32_t g = du.layout[sel[i]].group;
        const uint64_t c0 = du.group_cluster_off[g];
        const uint32_t n = (uint32_t)(du.group_cluster_off[g + 1] - c0);
        const uint64_t gidx = group_index(o, g);
        if (n > 1 && first && lane == 0) {
            for (uint64_t e = du.group_src_off[g]; e < du.group_src_off[g + 1]; e++) du.src_mut[e] = du.group_src[e];
            for (uint64_t e = du.cl_edge_off[c0]; e < du.cl_edge_off[c0 + n]; e++) du.edge_mut[e] = du.edge_dst[e];
        }
        for (uint32_t j = 0; j < n; j++) {
            Cl cl;
            cl.bind(du, (uint32_t)(c0 + j));
            Philox prng, fr;
            if (first) {
                cl_construct_warp(cl, o, gidx, stream_chain, lane);
                prng.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngGenotyper, stream_chain);
                fr.init(o.random_seed, gidx, du.cluster_idx[cl.c], kRngFrequency, stream_chain);
                if (lane == 0) fr.save(cl.rng, kRng1);
            } else {
                prng.load(cl.rng, kRng0, o.random_seed, gidx, du.cluster_idx[cl.c]);
            }
            if (n > 1) clw_reset<true>(cl, o, prng, lane, false); else clw_reset<false>(cl, o, prng, lane, false);
            if (lane == 0) prng.save(cl.rng, kRng0);
            __syncwarp();
        }
        if (n > 1) {
            if (lane == 0) group_branch_order(du, o, g, chain - 1);  // InferenceEngine.cpp:71: seed of chain index chain - 1
            __syncwarp();
        }
    }
    if (blockIdx.x == 0 && ns.trace) noise_update_block(ns, du.S, prior_shape, prior_scale, o.random_seed, 3, 0, (double)chain, 0, 1, sh_rates);
    grid_barrier(gb);
    // BTG_NOISE_PHASES=1: block 0's laps through [fill + small groups, large clusters, exchange + update, release]; [4] = slowest fill
    // task, [5] = slowest large cluster of phase B, [6] = slowest other group (ns << 32 | cluster)
    const bool timing = ns.phase_ns && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_prev = timing ? global_timer_ns() : 0;
    auto lap = [&](int phase) {
        if (timing) { const unsigned long long t = global_timer_ns(); ns.phase_ns[phase] += t - t_prev; t_prev = t; }
    };
    if (timing) ns.phase_ns[7] += t_prev - t_start;
    for (uint32_t it = 1; it <= iters; it++) {
        if (threadIdx.x < 2 * du.S) sh_stat[threadIdx.x] = 0;
        __syncthreads();
        const bool collect = joint && it > o.gibbs_burn_in;
        // phase A: fill tasks of the large clusters, dealt from the LAST warp downwards (the other groups occupy the first warps) ...
        for (uint32_t t = n_warps - 1 - w; t < n_fill_tasks; t += n_warps) {
            const unsigned long long t_in = ns.phase_ns ? global_timer_ns() : 0;
            Cl cl;
            cl.bind(du, sel[fill_tasks[3 * t]]);
            clw_fill_cache(cl, T, du.group_ploidy + (size_t)cl.g * du.S, lane, fill_tasks[3 * t + 1], fill_tasks[3 * t + 2]);
            if (ns.phase_ns && lane == 0) atomicMax(ns.phase_ns + 4, ((global_timer_ns() - t_in) << 32) | cl.c);
        }
        // ... while every other group takes its whole step (sampleGenotypesCallback)
        for (uint32_t i = n_big + w; i < n_sel; i += n_warps) {
            const unsigned long long t_in = ns.phase_ns ? global_timer_ns() : 0;
            const uint32_t g = du.layout[sel[i]].group;
            if (du.group_cluster_off[g + 1] - du.group_cluster_off[g] > 1) group_iteration<true>(du, T, o, g, collect, sh_stat, lane);
            else group_iteration<false>(du, T, o, g, collect, sh_stat, lane);
            if (ns.phase_ns && lane == 0) atomicMax(ns.phase_ns + 6, ((global_timer_ns() - t_in) << 32) | sel[i]);
        }
        if (n_fill_tasks) grid_barrier(gb);
        lap(0);
        // phase B: the large clusters sample from their filled caches
        for (uint32_t i = w; i < n_big; i += n_warps) {
            const unsigned long long t_in = ns.phase_ns ? global_timer_ns() : 0;
            group_iteration<false>(du, T, o, du.layout[sel[i]].group, collect, sh_stat, lane);
            if (ns.phase_ns && lane == 0) atomicMax(ns.phase_ns + 5, ((global_timer_ns() - t_in) << 32) | sel[i]);
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
}

}  // namespace

namespace btg_gibbs {

cudaError_t wide_estimate_genotypes(const DevUnit &du, const Tables &T, const btg_gibbs_opts &o, const ResultView &R, cudaStream_t st) {
    if (du.n_nested_groups) {
        k_estimate_genotypes_nested_wide<<<(du.n_nested_groups + 3) / 4, 128, 0, st>>>(du, T, o, R);
        BTG_LAUNCHED();
    }
    const uint32_t n_warps = du.n_regular + du.n_split * kChainSplit;
    if (n_warps) {
        k_estimate_genotypes_wide<<<(n_warps + 3) / 4, 128, 0, st>>>(du, T, o, R);
        BTG_LAUNCHED();
    }
    return cudaGetLastError();
}

cudaError_t wide_noise_chain(const DevUnit &du, const Tables &T, const btg_gibbs_opts &o, const uint32_t *d_sel, uint32_t n_sel, uint32_t n_big,
                             const uint32_t *d_tasks, uint32_t n_tasks, uint32_t chain, uint32_t iters,
                             const NoiseState &ns, float prior_shape, float prior_scale, unsigned long long *hist, int joint, const PeerExchange &px,
                             const GridBarrier &gb, uint32_t share, int sm_count, cudaStream_t st) {
    const uint32_t bs = 256;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_noise_chain_wide, bs, 0);
    const uint32_t capacity = (uint32_t)std::max(1, per_sm) * (uint32_t)sm_count;
    const uint32_t max_blocks = std::max(1u, (share > 1 ? capacity - capacity / 16 : capacity) / std::max(1u, share));
    const uint32_t grid = std::max(1u, std::min((std::max(n_sel, n_tasks) + 7) / 8, max_blocks));
    DevUnit du_ = du; Tables T_ = T; btg_gibbs_opts o_ = o; NoiseState ns_ = ns; PeerExchange px_ = px; GridBarrier gb_ = gb;
    void *args[] = {&du_, &T_, &o_, &d_sel, &n_sel, &n_big, &d_tasks, &n_tasks, &chain, &iters, &ns_, &prior_shape, &prior_scale, &hist, &joint, &px_, &gb_};
    BTG_LAUNCHED();
    return cudaLaunchCooperativeKernel((void *)k_noise_chain_wide, dim3(grid), dim3(bs), args, 0, st);
}

}  // namespace btg_gibbs
