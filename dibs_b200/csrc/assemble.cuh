// The per-particle "assemble" step -- merge the Monte-Carlo / acyclicity partials, graph prior, chain rule through
// S = U V^T, Gaussian prior on Z, next step's sub-keys -- as a device function, and its fusion into the gradient
// kernels of the step loop (the particle's last gradient CTA runs it).
//
// replaces: the final reshapes of dibs/inference/dibs.py:376-385,451-457,531-549, log_graph_prior_particle and
// eltwise_grad_latent_prior (dibs.py:604-658), dibs/models/graph.py:93-108,182-196,263-276 (unnormalized_log_prob_soft),
// and the key splits of dibs/inference/svgd.py:245,251,695,699,703.
#pragma once
#include "common.cuh"
#include "kernels_peer.cuh"

namespace dibs {

// inputs of the per-particle assemble step (assemble_particle below)
struct AsmParams {
    const float* z; int z_ld;
    const float* scores;                  // [n_local][d*d] raw U V^T (k_prologue)
    int n_local, d, k;
    const StepState* st; int t_override;
    float alpha_linear, beta_linear;
    // Z-likelihood partials
    const float* zacc; const float* zstats; int z_chunks; int z_mode;   // MC_Z_SCORE / MC_Z_REPARAM; zacc null = skip
    int n_samples;
    float sf_coef;                        // score_function_baseline
    const float* baselines_in; float* baselines_out;
    // theta partials
    const float* thacc; const float* thstats; int th_chunks; int th_dim;
    // prior
    const float* acyc; int n_acyc;        // [n_local][acyc_chunks][d*d] sums over A samples; null = skip prior terms entirely
    int acyc_chunks;
    int constraint_only;                  // hook: return mean_a grad h alone (no beta, no other terms)
    int prior_kind; float er_coef;        // log p - log(1-p)
    float sigma_z2;                       // latent_prior_std ** 2
    float* grad_z; int gz_ld;
    float* grad_th; int gth_ld;
    // next step's loop state and per-pass sub-keys (step loop only; null: skip)
    uint32_t* next_keys; StepState* st_next;
    int n_step_splits, n_particles, partitionable, m_offset; uint32_t pre_split_mask;
    // peer-memory exchange fused into the kernel: every gradient value is also stored into the same row of each
    // peer's gradient buffer, the last CTA raises the flags (push.world == 0: off)
    PeerPush push;
};

// assemble fused into the gradient kernels of the step loop (fuse_arrive below)
struct FuseAsm {
    uint32_t* arrive;        // [n_local] arrival counters (null: not fused -- hooks)
    int total;               // CTAs per particle over all gradient kernels of the step
    AsmParams a;
};

__device__ __forceinline__ void fuse_arrive(const FuseAsm& f, int m, float* smem);   // defined below

// merged softmax normaliser of the chunk partials: weights w_c = exp(m_c - max) / sum_c l_c exp(m_c - max) into sW[c]
// (one warp; chunks <= 32 handled by lanes, more by a strided loop), and sum of log-probs into *sum_lp
__device__ __forceinline__ void merge_stats_warp(const float* __restrict__ stats, int chunks, float* sW, float* sum_lp,
                                                 int lane) {
    // (max, sum of exp, sum of log-probs, -) of chunk `lane` in one 128-bit read (the common case: chunks <= 32)
    const float4* st4 = reinterpret_cast<const float4*>(stats);
    const float4 mine = lane < chunks ? __ldcg(st4 + lane) : make_float4(-INFINITY, 0.0f, 0.0f, 0.0f);
    float mx = mine.x;
    for (int c = lane + 32; c < chunks; c += 32) mx = fmaxf(mx, __ldcg(stats + c * 4));
    mx = warp_max(mx);
    float l = 0.0f, sl = 0.0f;
    for (int c = lane; c < chunks; c += 32) {
        const float4 sv = c == lane ? mine : __ldcg(st4 + c);
        const float w = (sv.x == -INFINITY) ? 0.0f : expf(sv.x - mx);
        sW[c] = w;
        l += sv.y * w;
        sl += sv.z;
    }
    // fixed-order (butterfly) sums: deterministic
    l = warp_sum(l); sl = warp_sum(sl);
    __syncwarp();
    for (int c = lane; c < chunks; c += 32) sW[c] = sW[c] / l;
    if (lane == 0) *sum_lp = sl;
}

// sum_c w[c] * src[c * stride] (w == nullptr: plain sum) in chunk order, the L2 reads issued four at a time: a rolled
// loop would serialise one L2 round trip per chunk, and this step is nothing but latency
__device__ __forceinline__ float chunk_sum_cg(const float* __restrict__ src, size_t stride, int chunks, const float* w) {
    float acc = 0.0f;
    for (int c0 = 0; c0 < chunks; c0 += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (c0 + u < chunks) ? __ldcg(src + (size_t)(c0 + u) * stride) : 0.0f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c0 + u < chunks) acc = w ? fmaf(v[u], w[c0 + u], acc) : acc + v[u];
    }
    return acc;
}

// Per-particle assemble step, executed by ONE whole CTA of any size >= 32 threads (all threads must call it; `smem`
// = assemble_smem() bytes of shared memory the caller no longer needs).  Two callers:
//   * k_assemble_grad (hooks): one CTA per particle;
//   * the step loop: the LAST CTA of the particle's gradient kernels to finish (Monte-Carlo passes and acyclicity pass,
//     three kernels on sibling streams; see fuse_arrive below) -- the row's partials are in L2, no extra kernel sits
//     between the gradient phase and phi on the critical path.
// Partials written by other CTAs are read with ld.global.cg (L2): they may be younger than this SM's L1 lines.
// Latency-bound by construction (a few KB of inputs): every global read is issued as early as possible and branch-free
// so the loads of a phase overlap; the serial threefry chains of the NEXT step's sub-keys run on an otherwise idle warp.
static __device__ __noinline__ void assemble_particle(const AsmParams& p, int m, float* smem) {
    const int d = p.d, k = p.k, dd = d * d, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarp = nthr >> 5;
    const int t = p.st ? p.st->t : p.t_override;
    const float alpha = p.alpha_linear * (float)t;
    const float beta = p.beta_linear * (float)t;
    float* sZ = smem;                 // [2dk]
    float* sP = sZ + 2 * d * k;       // [d*d]
    float* sDS = sP + dd;             // [d*d]
    float* sCol = sDS + dd;           // [d]
    float* sWz = sCol + d;            // [z_chunks]
    float* sWt = sWz + p.z_chunks;    // [th_chunks]
    float* sMisc = sWt + p.th_chunks; // [2] sum of z log-probs, sum of theta log-probs

    // ---- phase A: stage Z and the edge probabilities; one warp each merges the softmax statistics of the two passes,
    // the last warp derives keys (with fewer than three warps the same warp does these one after the other)
    // This function runs ONCE per particle and step, from cold instruction cache (32 KB per SM, the gradient kernels'
    // main loops own it): its cost is instruction fetch as much as data latency.  Loops whose trip count is 2-5 are kept
    // rolled (an unrolled-by-4 body would never execute, only bloat), and the threefry / softmax-merge code exists at
    // ONE call site each.
    const float* zrow = p.z + (size_t)m * p.z_ld;
#pragma unroll 1
    for (int e = tid; e < 2 * d * k; e += nthr) sZ[e] = zrow[e];
    const float* srow = p.scores + (size_t)m * dd;
#pragma unroll 1
    for (int e = tid; e < dd; e += nthr) {
        const int i = e / d, j = e - i * d;
        sP[e] = (i == j) ? 0.0f : sigmoidf_ref(alpha * srow[e]);
    }
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        // warp 0: Z pass, warp 1: Theta pass (a one-warp CTA does both in turn)
        const float* acc = which ? p.thacc : p.zacc;
        if (warp != which % nwarp || acc == nullptr) continue;
        const int chunks = which ? p.th_chunks : p.z_chunks;
        merge_stats_warp((which ? p.thstats : p.zstats) + (size_t)m * chunks * 4, chunks, which ? sWt : sWz, sMisc + which, lane);
    }
    if (warp == nwarp - 1 && p.next_keys) {
        // loop state: key <- after this step's (M+1)-way splits, t <- t + 1 (svgd.py:245,251,272); the sub-keys of
        // the next step (svgd.py:245,251 / 695,699,703 and the pre-draw splits dibs.py:350,430) for this particle.
        // Lane L < n_step_splits derives sub-key L: n_step_splits + L splits "row 0 of M + 1" (the first n_step_splits
        // of them give the carried key), then row m + 1 of M + 1, then (pre-draw passes) row 1 of 2 -- one loop, the
        // split's row / count chosen per lane and step
        const int ns = p.n_step_splits;
        const uint32_t big = (uint32_t)p.n_particles + 1u, mine = (uint32_t)(p.m_offset + m) + 1u;
        const bool pre = (p.pre_split_mask >> lane) & 1u;
        const bool part = p.partitionable != 0;
        uint2 sk = make_uint2(p.st->key[0], p.st->key[1]), key = sk;
        const int my_steps = lane < ns ? ns + lane + 1 + (pre ? 1 : 0) : ns;
#pragma unroll 1
        for (int w = 0; w < 2 * ns + 1; ++w) {
            if (w < my_steps) {
                const bool own = w == ns + lane, last = w == ns + lane + 1;
                sk = jax_split_row(sk, own ? mine : (last ? 1u : 0u), last ? 2u : big, part);
            }
            if (w == ns - 1) key = sk;
        }
        if (lane < ns) {
            uint32_t* o = p.next_keys + ((size_t)lane * p.n_local + m) * 2;
            o[0] = sk.x; o[1] = sk.y;
        }
        if (m == 0 && lane == 0) {
            p.st_next->key[0] = key.x; p.st_next->key[1] = key.y;
            p.st_next->t = t + 1; p.st_next->pad = 0;
        }
    }
    __syncthreads();
    if (p.acyc && !p.constraint_only && p.prior_kind == 1) {
        if (tid < d) {
            float indeg = 0.0f;
            for (int i = 0; i < d; ++i) indeg += sP[i * d + tid];        // soft_g.sum(0)  (graph.py:195)
            sCol[tid] = -3.0f / (1.0f + indeg);
        }
        for (int jj = tid + nthr; jj < d; jj += nthr) {                  // CTAs smaller than d
            float indeg = 0.0f;
            for (int i = 0; i < d; ++i) indeg += sP[i * d + jj];
            sCol[jj] = -3.0f / (1.0f + indeg);
        }
        __syncthreads();
    }
    const float base_fac = (p.zacc && p.z_mode == MC_Z_SCORE && p.sf_coef > 0.0f) ? expf(-p.baselines_in[m]) : 1.0f;
    // ---- phase B: dS.  The parameter block is read through a generic reference: everything the loops use is taken
    // into registers first (a store in the loop would otherwise force its re-load, one more round trip each time)
    const float inv_acyc = 1.0f / (float)p.n_acyc;
    {
        const int z_chunks = p.z_chunks, acyc_chunks = p.acyc_chunks, z_mode = p.z_mode, prior_kind = p.prior_kind;
        const int constraint_only = p.constraint_only;
        const float er_coef = p.er_coef;
        const float* za0 = p.zacc ? p.zacc + (size_t)m * z_chunks * dd : nullptr;
        const float* ac0 = p.acyc ? p.acyc + (size_t)m * acyc_chunks * dd : nullptr;
        auto ds_of = [&](int e, float w, float acs) -> float {
            const int i = e / d, j = e - i * d;
            float ds = 0.0f;
            if (i != j) {
                const float pe = sP[e];
                // score: e^{-b} alpha (Gbar - P) (App. B-1/2; dibs.py:363-382); reparam: softmax-weighted dS
                if (za0) ds += (z_mode == MC_Z_SCORE) ? base_fac * alpha * (w - pe) : w;
                if (ac0) {
                    const float acm = acs * inv_acyc;                                  // .mean(0)  (dibs.py:601)
                    if (constraint_only) ds += acm;
                    else {
                        ds -= beta * acm;
                        const float coef = prior_kind == 0 ? er_coef : (prior_kind == 1 ? sCol[j] : 0.0f);
                        ds += coef * alpha * pe * (1.0f - pe);                          // App. B-5
                    }
                }
            }
            return ds;
        };
        // two entries per trip: their 4 x (chunks) L2 reads are in flight together (the shared-memory store between
        // trips pins the order of the reads of different trips)
        for (int e = tid; e < dd; e += 2 * nthr) {
            const int e1 = e + nthr;
            const bool two = e1 < dd;
            const float w0 = za0 ? chunk_sum_cg(za0 + e, (size_t)dd, z_chunks, sWz) : 0.0f;
            const float w1 = (za0 && two) ? chunk_sum_cg(za0 + e1, (size_t)dd, z_chunks, sWz) : 0.0f;
            const float a0 = ac0 ? chunk_sum_cg(ac0 + e, (size_t)dd, acyc_chunks, nullptr) : 0.0f;
            const float a1 = (ac0 && two) ? chunk_sum_cg(ac0 + e1, (size_t)dd, acyc_chunks, nullptr) : 0.0f;
            const float d0 = ds_of(e, w0, a0);
            const float d1 = two ? ds_of(e1, w1, a1) : 0.0f;
            sDS[e] = d0;
            if (two) sDS[e1] = d1;
        }
    }
    // theta gradient: softmax-weighted partial sums (dibs.py:531-549); independent of the barrier below
    if (p.thacc) {
        const int th_dim = p.th_dim, th_chunks = p.th_chunks;
        float* gth = p.grad_th + (size_t)m * p.gth_ld;
        const float* ta = p.thacc + (size_t)m * th_chunks * th_dim;
        for (int e = tid; e < th_dim; e += 2 * nthr) {
            const int e1 = e + nthr;
            const bool two = e1 < th_dim;
            const float g0 = chunk_sum_cg(ta + e, (size_t)th_dim, th_chunks, sWt);
            const float g1 = two ? chunk_sum_cg(ta + e1, (size_t)th_dim, th_chunks, sWt) : 0.0f;
            gth[e] = g0;
            if (two) gth[e1] = g1;
        }
    }
    __syncthreads();
    // ---- phase C: chain rule through S = U V^T: dU = dS V, dV = dS^T U; Gaussian prior -Z/sigma^2 (dibs.py:657)
    float* gz = p.grad_z + (size_t)m * p.gz_ld;
    const bool gauss = p.acyc && !p.constraint_only;
    const float sigma_z2 = p.sigma_z2;
    for (int e = tid; e < d * k; e += nthr) {
        const int i = e / k, kk = e - i * k;
        float du = 0.0f, dv = 0.0f;
        for (int j = 0; j < d; ++j) {
            const float2 zj = *reinterpret_cast<const float2*>(&sZ[(j * k + kk) * 2]);
            du = fmaf(sDS[i * d + j], zj.y, du);
            dv = fmaf(sDS[j * d + i], zj.x, dv);
        }
        if (gauss) {
            du -= sZ[2 * e] / sigma_z2;
            dv -= sZ[2 * e + 1] / sigma_z2;
        }
        gz[2 * e] = du; gz[2 * e + 1] = dv;
    }
    if (p.zacc && p.baselines_out && tid == 0) {
        const float b_in = p.baselines_in ? p.baselines_in[m] : 0.0f;
        // dibs.py:388-389 (only the score estimator touches the baseline)
        p.baselines_out[m] = (p.z_mode == MC_Z_SCORE)
            ? p.sf_coef * (sMisc[0] / (float)p.n_samples) + (1.0f - p.sf_coef) * b_in : b_in;
    }
    if (p.push.world) {
        // fused exchange: the finished gradient row [dZ | dTheta] (just written, L1/L2-hot) goes to the same row of
        // every peer's buffer as 128-bit stores (rows are 16-byte aligned, stride a multiple of 4 floats)
        __syncthreads();
        const int n4 = p.gz_ld / 4, world = p.push.world, rank = p.push.rank;
        const size_t row4 = (size_t)(p.m_offset + m) * n4;
        const float4* src = reinterpret_cast<const float4*>(gz);
        // peer-outer: the destination pointer is read from the parameter block once per peer, not once per store
#pragma unroll 1
        for (int q = 0; q < world; ++q) {
            if (q == rank) continue;
            float4* dst = reinterpret_cast<float4*>(p.push.dst[q]) + row4;
            for (int e = tid; e < n4; e += nthr) dst[e] = __ldcg(src + e);
        }
        peer_signal(p.push, (unsigned)p.n_local);       // one signal per particle; the last one raises the flags
    }
}


// Fusion of the assemble step into the gradient kernels of the step loop.  Every CTA of the particle's Monte-Carlo and
// acyclicity kernels calls fuse_arrive() as its last action (all threads): it publishes the CTA's partials, counts the
// CTA on the particle's arrival counter and -- in the CTA that completes the count -- runs assemble_particle and
// re-arms the counter for the next step.  Summation orders are those of assemble_particle (fixed by chunk index), so
// the result does not depend on which CTA happens to be last.

__device__ __forceinline__ void fuse_arrive(const FuseAsm& f, int m, float* smem) {
    if (f.arrive == nullptr) return;
    __shared__ int s_fuse_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(f.arrive + m, 1u);
        s_fuse_last = (old == (unsigned)f.total - 1u) ? 1 : 0;
        if (s_fuse_last) f.arrive[m] = 0u;
    }
    __syncthreads();
    if (!s_fuse_last) return;
    __threadfence();
    assemble_particle(f.a, m, smem);
}

inline size_t assemble_smem(int d, int k, int z_chunks, int th_chunks) {
    return ((size_t)2 * d * k + 2 * (size_t)d * d + d + z_chunks + th_chunks + 4) * sizeof(float);
}

}  // namespace dibs
