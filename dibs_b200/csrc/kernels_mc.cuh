// Monte-Carlo passes over (particle x sampled graph): graph sampling fused with likelihood scoring,
// closed-form gradients and online-softmax aggregation.  Nothing of shape [M, S, d, d] is ever written.
//
// replaces (reference, larslorch/dibs): dibs/inference/dibs.py:325-391 (score-function estimator),
// :395-459 (Gumbel-softmax reparam estimator), :488-551 (grad_theta estimator) and the likelihood
// plugins they vmap over: dibs/models/linearGaussian.py:63-170 (BGe), :278-338 (LinearGaussian),
// dibs/models/nonlinearGaussian.py:248-326 (DenseNonlinearGaussian).
//
// Work decomposition: CTA = (particle, chunk of MC samples); thread = (sample, node j[, hidden unit h]).
// A thread owns column j of graph s in registers (its parent weights), streams the N observations from
// shared memory, and produces the node's log-prob and the column of x^T R that every gradient needs.
#pragma once
#include "common.cuh"
#include "assemble.cuh"

namespace dibs {

struct McParams {
    const float* z; int z_ld;            // particle latents, row stride (floats)
    const float* scores;                 // [n_local][d*d] raw U V^T from k_prologue (null: caller supplies graphs)
    const float* theta; int th_ld;       // particle parameters (may be null)
    int n_local;                         // particles in this launch
    int m_offset;                        // global index of the first particle (key derivation)
    int n_particles;                     // M, global
    int d, k, n_obs, n_samples;
    int n_chunks, s_per_chunk, gpb;      // chunking of the MC axis; graphs per block-round
    int paired;                          // units are sample pairs (s, s + S/2)
    const float* x;                      // [N, d]
    const int32_t* mask;                 // [N, d] or null
    const StepState* st; int which_split; int partitionable;
    const uint32_t* keys_override;       // [n_local, 2] per-particle sub-keys of this pass
    int t_override;                      // used when st == null
    float alpha_linear, tau;
    // LinearGaussian / DenseNN constants (fp32, rounded like the reference)
    float s2, log2pis2;                  // (sqrt(obs_noise))^2 and log(2 pi s2)
    float mean_edge, sig2_edge, lognorm_edge;   // N(theta; mean_edge, sig_edge): sig^2, log(2 pi sig^2)
    int hidden, hp;                      // DenseNN: H and next pow2 >= H
    int activation;                      // DenseNN: 0 relu, 1 tanh, 2 sigmoid, 3 leakyrelu(0.01)  (stax, nonlinearGaussian.py:52-61)
    // BGe
    const double* bge_r;                 // [d or 1][d][d] fp64 R_j
    int bge_r_stride;                    // d*d if per-node R (interventions), 0 if shared
    const float* bge_table;              // [d][d+1]: log-gamma terms for (node j, n_parents l)
    const float* bge_coef;               // [d][2]: (N_j + alpha_lambd - d), valid flag
    float bge_alpha_mu, bge_alpha_lambd; // hyper-parameters (soft-graph BGe: log-gamma terms at real-valued parent counts)
    const float* g_ext;                  // [n_local, S, d, d] graphs supplied by the caller (hooks) or null
    float* part_acc; int acc_size;       // [n_local][n_chunks][acc_size]
    float* part_stats;                   // [n_local][n_chunks][4]: running max, sum exp, sum lp, -
    float* lp_out;                       // optional [n_local][S]
    FuseAsm fuse;                        // step loop: the particle's last gradient CTA runs the assemble step
};

__device__ __forceinline__ float norm_logpdf_pre(float x, float loc, float sig2, float lognorm) {
    // jax.scipy.stats.norm.logpdf: -(log(2 pi s^2) + (x-loc)^2 / s^2) / 2
    float dx = x - loc;
    return -(lognorm + dx * dx / sig2) * 0.5f;
}

// Per-thread graph entry G_s[i][j] for flat sample index s (global over the particle's S samples).
template <bool HARD>
__device__ __forceinline__ float graph_entry(const McParams& p, uint2 key, const float* sA, int s, int i, int j, int d,
                                             float tau) {
    if (i == j) return 0.0f;  // zero_diagonal (utils/func.py:117-125); the draw is consumed and discarded
    uint32_t e = ((uint32_t)s * d + i) * d + j;
    uint32_t bits = jax_bits(key, e, (uint32_t)p.n_samples * d * d, p.partitionable);
    if (HARD) return bits_to_unit(bits) < sA[i * d + j] ? 1.0f : 0.0f;
    return sigmoidf_ref(tau * (logistic_from_bits(bits) + sA[i * d + j]));
}

// Shared prologue: alpha * scores (soft) or edge probabilities sigmoid(alpha * scores), zero diagonal (hard), into
// sA[i*d + j]; returns alpha.  The raw scores U V^T come from k_prologue (one evaluation per particle per step).
__device__ __forceinline__ float stage_scores(const McParams& p, int m, float* sA, bool hard, int t) {
    const int d = p.d, dd = d * d;
    const float alpha = p.alpha_linear * (float)t;  // dibs.py:70: fp32 product of slope and step
    const float* srow = p.scores ? p.scores + (size_t)m * dd : nullptr;
    for (int e = threadIdx.x; e < dd; e += blockDim.x) {
        const float a = srow ? alpha * srow[e] : 0.0f;
        const int i = e / d;
        sA[e] = hard ? (e == i * (d + 1) ? 0.0f : sigmoidf_ref(a)) : a;
    }
    __syncthreads();
    return alpha;
}

// per-particle key of this pass (already pre-split by k_prologue where the reference splits first)
__device__ __forceinline__ uint2 mc_key(const McParams& p, int m_local) {
    return make_uint2(p.keys_override[2 * m_local], p.keys_override[2 * m_local + 1]);
}

// Online-softmax bookkeeping shared by all likelihood kernels. Call by all threads of the CTA.
// sNode: [gpb][d] per-node log-probs written by the owners; sLpS: [gpb].
struct SoftmaxRun {
    float m_run, l_run, sum_lp;
};

// ------------------------------------------------------------------------------------------
// LinearGaussian (linearGaussian.py:278-338)
// ------------------------------------------------------------------------------------------
template <int DMAX, int MODE>
__global__ void __launch_bounds__(256) k_mc_lingauss(const __grid_constant__ McParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr bool HARD = (MODE == MC_THETA_HARD || MODE == MC_Z_SCORE);
    const int d = p.d, N = p.n_obs, gpb = p.gpb;
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;

    float* sA = smem;                       // [d*d]
    float* sTh = sA + d * d;                // [d*d]
    float* sCnt = sTh + d * d;              // [d] unmasked observation count per node
    float* sNode = sCnt + d;                // [gpb*d]
    float* sLpS = sNode + gpb * d;          // [gpb]
    float* sBig = smem + ((2 * d * d + d + gpb * d + gpb + 3) & ~3);  // 16B aligned
    float* sX = sBig;                       // [N*DMAX] zero-padded rows
    float* sKeep = sX + N * DMAX;           // [N*d] (only if mask)

    const bool use_ext = p.g_ext != nullptr;
    const float alpha = stage_scores(p, m, sA, HARD, t);
    const float* throw_ = p.theta + (size_t)m * p.th_ld;
    for (int e = tid; e < d * d; e += blockDim.x) sTh[e] = throw_[e];
    for (int e = tid; e < N * DMAX; e += blockDim.x) {
        int n = e / DMAX, i = e % DMAX;
        sX[e] = i < d ? p.x[n * d + i] : 0.0f;
    }
    if (p.mask)
        for (int e = tid; e < N * d; e += blockDim.x) sKeep[e] = p.mask[e] ? 0.0f : 1.0f;
    __syncthreads();
    if (tid < d) {
        float cnt = (float)N;
        if (p.mask) { cnt = 0.0f; for (int n = 0; n < N; ++n) cnt += sKeep[n * d + tid]; }
        sCnt[tid] = cnt;
    }
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);
    __syncthreads();

    const bool active = tid < gpb * d;
    const int s_local = tid / d, j = tid % d;
    const int s_begin = c * p.s_per_chunk;
    const int s_end = min(p.n_samples, s_begin + p.s_per_chunk);
    const float inv_s2 = 1.0f / p.s2;

    float acc[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;

    for (int s0 = s_begin; s0 < s_end; s0 += gpb) {
        const int s = s0 + s_local;
        const bool valid = active && s < s_end;
        float w[DMAX], b[DMAX], gf[DMAX];
        float node_lp = 0.0f;
        if (valid) {
            float prior = 0.0f;
#pragma unroll
            for (int i = 0; i < DMAX; ++i) {
                float g = 0.0f, th = 0.0f;
                if (i < d) {
                    g = use_ext ? (i == j ? 0.0f : p.g_ext[(((size_t)m * p.n_samples + s) * d + i) * d + j])
                                : graph_entry<HARD>(p, key, sA, s, i, j, d, p.tau);
                    th = sTh[i * d + j];
                    // log p(theta | G): sum g * logN(theta; mean_edge, sig_edge)  (linearGaussian.py:289)
                    prior = fmaf(g, norm_logpdf_pre(th, p.mean_edge, p.sig2_edge, p.lognorm_edge), prior);
                }
                gf[i] = g; w[i] = g * th; b[i] = 0.0f;
            }
            float ssq = 0.0f;
            for (int n = 0; n < N; ++n) {
                const float4* xr = reinterpret_cast<const float4*>(sX + n * DMAX);
                float xv[DMAX];
#pragma unroll
                for (int q = 0; q < DMAX / 4; ++q) {
                    float4 v = xr[q];
                    xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
                }
                float mean = 0.0f;
#pragma unroll
                for (int i = 0; i < DMAX; ++i) mean = fmaf(xv[i], w[i], mean);
                float r = sX[n * DMAX + j] - mean;       // x - x @ (g * theta)   (linearGaussian.py:314)
                if (p.mask) r *= sKeep[n * d + j];       // jnp.where(interv_targets, 0, .)
                ssq = fmaf(r, r, ssq);
#pragma unroll
                for (int i = 0; i < DMAX; ++i) b[i] = fmaf(xv[i], r, b[i]);   // column j of x^T R
            }
            node_lp = prior - 0.5f * (sCnt[j] * p.log2pis2 + ssq * inv_s2);
            sNode[s_local * d + j] = node_lp;
        }
        __syncthreads();
        if (tid < gpb) {
            float lp = -INFINITY;
            if (s0 + tid < s_end) {
                lp = 0.0f;
                for (int jj = 0; jj < d; ++jj) lp += sNode[tid * d + jj];
                if (p.lp_out) p.lp_out[(size_t)m * p.n_samples + s0 + tid] = lp;
            }
            sLpS[tid] = lp;
        }
        __syncthreads();
        if (MODE != MC_LP_ONLY) {
            float m_new = m_run;
            for (int g = 0; g < gpb; ++g) m_new = fmaxf(m_new, sLpS[g]);
            const float scale = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
            float lsum = 0.0f, lpsum = 0.0f;
            for (int g = 0; g < gpb; ++g) {
                float lp = sLpS[g];
                if (lp != -INFINITY) { lsum += expf(lp - m_new); lpsum += lp; }
            }
            l_run = l_run * scale + lsum;
            sum_lp += lpsum;
            m_run = m_new;
            const float e = valid ? expf(sLpS[s_local] - m_new) : 0.0f;
#pragma unroll
            for (int i = 0; i < DMAX; ++i) {
                float val = 0.0f;
                if (valid && i < d) {
                    if (MODE == MC_THETA_HARD) {
                        // d/dtheta: g * (-(theta-mu)/sig^2) + g * (x^T R)/s2          
                        float th = sTh[i * d + j];
                        val = gf[i] * (-(th - p.mean_edge) / p.sig2_edge + b[i] * inv_s2);
                    } else if (MODE == MC_Z_REPARAM) {
                        // dS = d lp/dG * tau*alpha*g(1-g)                                (App. B-4, B-6)
                        float th = sTh[i * d + j];
                        float dg = norm_logpdf_pre(th, p.mean_edge, p.sig2_edge, p.lognorm_edge) + th * b[i] * inv_s2;
                        val = dg * (p.tau * alpha) * gf[i] * (1.0f - gf[i]);
                    } else {
                        val = gf[i];  // score function: weighted mean graph (App. B-2)
                    }
                }
                acc[i] = acc[i] * scale + e * val;
            }
        }
        __syncthreads();
    }
    if (MODE == MC_LP_ONLY) return;

    // deterministic reduction over the gpb sample slots: sRed[s_local][i*d+j]
    float* sRed = sBig;
    if (active) {
#pragma unroll
        for (int i = 0; i < DMAX; ++i)
            if (i < d) sRed[(size_t)s_local * d * d + i * d + j] = acc[i];
    }
    __syncthreads();
    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    for (int e = tid; e < d * d; e += blockDim.x) {
        float sum = 0.0f;
        for (int g = 0; g < gpb; ++g) sum += sRed[(size_t)g * d * d + e];
        out[e] = sum;
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t mc_lingauss_smem(int d, int k, int n_obs, int gpb, int dmax, bool has_mask) {
    size_t head = 2 * (size_t)d * d + d + (size_t)gpb * d + gpb;
    head = (head + 3) & ~(size_t)3;
    size_t big = (size_t)n_obs * dmax + (has_mask ? (size_t)n_obs * d : 0);
    size_t red = (size_t)gpb * d * d;
    size_t zz = (size_t)2 * d * k;
    if (red > big) big = red;
    if (zz > big) big = zz;
    return (head + big + 4) * sizeof(float);
}

}  // namespace dibs
