// LinearGaussian Monte-Carlo pass, observational data, n_vars <= 32: QR-factor form.
//
// replaces (same reference functions as kernels_mc.cuh::k_mc_lingauss): dibs/models/linearGaussian.py:278-338
// (log_prob_parameters, log_likelihood, interventional_log_joint_prob) under the estimators of
// dibs/inference/dibs.py:325-459 (grad_z score / reparam) and :488-551 (grad_theta).
//
// The reference evaluates, per sampled graph, the N x d residual  R = x - x (G o Theta)  and needs from it only
//   ssq_j = sum_n R_nj^2            (the Gaussian log-likelihood, linearGaussian.py:314-316)
//   B_:j  = x^T R_:j                (every gradient).
// With x = Q Rx (thin QR, done once per dibs_set_data in fp64) and u_j = e_j - (G o Theta)_:j:
//   R_:j = x u_j = Q (Rx u_j)   =>   ssq_j = |Rx u_j|^2 ,   B_:j = Rx^T (Rx u_j).
// Two d x d triangular mat-vecs per (graph, node) -- d^2 FMAs instead of 2 N d -- with the SAME conditioning
// as the reference's direct form (a sum of squares of one fp32 mat-vec; no Gram-matrix cancellation).
//
// Work decomposition: CTA = (particle, chunk of sample PAIRS); thread = (pair q, node j) owns the columns j of the
// two graphs s = q and s = q + S/2 from the draw to the accumulation: in JAX's legacy threefry layout those two
// draws are the two lanes of ONE threefry block (counter pair (e, e + n/2)), so no random bits are thrown away, and
// a thread draws exactly the entries it consumes -- into registers, nothing is staged for other threads.  The only
// cross-thread step of a round is the sum of the d node log-probs of a sample: ONE barrier per round, after which
// every warp redoes the (tiny) per-sample reduction and softmax bookkeeping for itself.  (Host side: odd S or the
// partitionable PRNG layout fall back to k_mc_lingauss, which draws one entry per block.)
// Rx lives in the kernel-parameter constant bank: with the mat-vec loops fully unrolled every FMA takes its
// Rx operand from a uniform register (LDCU), no shared-memory traffic at all in the inner loops.
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"

namespace dibs {

// packed upper triangle, row i / column k >= i at i*DMAX - i(i-1)/2 + (k - i); zero-padded to DMAX
template <int DMAX>
struct RTri { float v[DMAX * (DMAX + 1) / 2]; };

template <int DMAX>
__host__ __device__ constexpr int rtri_off(int i, int k) { return i * DMAX - i * (i - 1) / 2 + (k - i); }

// uniform in [eps, 1) that random.logistic feeds to log(u) - log1p(-u)
__device__ __forceinline__ float logistic_u_from_bits(uint32_t bits) {
    const float eps = 1.1920928955078125e-07f;
    return fmaxf(eps, __fadd_rn(__fmul_rn(bits_to_unit(bits), 1.0f - eps), eps));
}

// value of one graph entry from its random bits.  sa = P_ij (hard) | exp(-alpha s_ij) (soft, tau == 1) | alpha s_ij
template <bool HARD>
__device__ __forceinline__ float entry_from_bits(uint32_t bits, float sa, bool fast_soft, float tau) {
    if (HARD) return bits_to_unit(bits) < sa ? 1.0f : 0.0f;
    const float u = logistic_u_from_bits(bits);
    // sigmoid(log(u/(1-u)) + a) = u / (u + (1-u) e^{-a}): the logistic noise and the sigmoid cancel analytically
    // u / (u + (1-u) e^{-a}) with the approximate reciprocal (1 ulp): one MUFU + one multiply; an infinite
    // denominator (e^{-a} overflowed) gives 0, as the division would
    if (fast_soft) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(1.0f - u, sa, u)));
        return u * r;
    }
    return sigmoidf_ref(tau * ((logf(u) - log1pf(-u)) + sa));
}

// shared-memory strides: thread (slot, j) owns PRIVATE columns of the per-slot tables (stride-DMAX walks down a
// column); a slot stride congruent to d (mod 32) makes the bank of such an access equal to the flat thread index
__host__ __device__ inline int qr_pad_stride(int base, int d) { return base + ((d - base) % 32 + 32) % 32; }

// FULL: n_vars == DMAX, the variable count is a compile-time constant (no padding rows: every `i < d` folds away)
template <int DMAX, int MODE, bool FULL>
__global__ void __launch_bounds__((DMAX <= 20 && MODE == MC_THETA_HARD ? 256 : 192), (DMAX <= 20 ? 2 : 1))
k_mc_lin_qr(const __grid_constant__ McParams p, const __grid_constant__ RTri<DMAX> R) {
    extern __shared__ __align__(16) float smem[];
    constexpr bool HARD = (MODE == MC_THETA_HARD || MODE == MC_Z_SCORE);
    constexpr int MAT = DMAX * DMAX;                   // all per-(i,j) tables use the compile-time stride DMAX
    constexpr int NS = DMAX + 1;                       // odd row stride of the node log-prob table
    constexpr int IL = 4;                              // independent threefry chains per thread
    static_assert(DMAX % IL == 0 && DMAX <= 32, "row batches of 4; one bit per row in the hard-graph masks");
    const int d = FULL ? DMAX : p.d, dd = d * d, gpb = p.gpb;        // gpb = sample pairs ("slots") per round, <= 16
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int t = p.st ? p.st->t : p.t_override;
    const int S = p.n_samples, Qh = (S + 1) >> 1;      // slot q holds samples q and q + Qh
    const int acc_stride = qr_pad_stride(MAT, d), g_stride = qr_pad_stride(2 * MAT, d);

    float* sA = smem;                       // [MAT] hard: P_ij; soft tau==1: exp(-alpha s_ij); soft: alpha s_ij
    float* sTh = sA + MAT;                  // [MAT] theta_ij
    float* sLpTh = sTh + MAT;               // [MAT] logN(theta_ij; mean_edge, sig_edge)
    float* sAccAll = sLpTh + MAT;           // [gpb][acc_stride] softmax-weighted running sums, one private column per thread
    float* sGall = sAccAll + gpb * acc_stride;      // [gpb][g_stride] soft graph entries of the round (private columns)
    float* sNode = sGall + gpb * g_stride;  // [2][2*gpb][NS] per-node log-probs, double-buffered by round parity

    // caller-supplied graphs exist for the log-prob hook only (dibs_log_joint_prob): a compile-time false elsewhere
    const bool use_ext = (MODE == MC_LP_ONLY) && p.g_ext != nullptr;
    const bool fast_soft = !HARD && !use_ext && p.tau == 1.0f;
    const float alpha = p.alpha_linear * (float)t;     // dibs.py:70: fp32 product of slope and step
    {
        const float* srow = p.scores ? p.scores + (size_t)m * dd : nullptr;
        const float* throw_ = p.theta + (size_t)m * p.th_ld;
        const float inv_d_f = 1.0f / (float)d;
        for (int e = tid; e < dd; e += blockDim.x) {
            const int i = __float2int_rz(((float)e + 0.5f) * inv_d_f), j = e - i * d;
            const float a = srow ? alpha * srow[e] : 0.0f;
            float sa;
            if (HARD) sa = (i == j) ? 0.0f : sigmoidf_ref(a);             // edge_probs (dibs.py:168-184)
            else sa = fast_soft ? expf(-a) : a;
            const float th = throw_[e];
            sA[i * DMAX + j] = sa;
            sTh[i * DMAX + j] = th;
            sLpTh[i * DMAX + j] = norm_logpdf_pre(th, p.mean_edge, p.sig2_edge, p.lognorm_edge);
        }
    }
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);

    const int slot = tid / d, j = tid - slot * d;
    const bool active = slot < gpb;
    const int q_begin = c * p.s_per_chunk;
    const int q_end = min(Qh, q_begin + p.s_per_chunk);
    const float inv_s2 = 1.0f / p.s2, inv_se2 = 1.0f / p.sig2_edge;
    const uint32_t half = ((uint32_t)S * dd) >> 1;
    const float cst = (float)p.n_obs * p.log2pis2;

    float* sAcc = sAccAll + (active ? slot : 0) * acc_stride + j;     // sAcc[i*DMAX]
    float* sG = sGall + (active ? slot : 0) * g_stride + j;           // sG[g*MAT + i*DMAX]
    if (active && MODE != MC_LP_ONLY) {
#pragma unroll
        for (int i = 0; i < DMAX; ++i) sAcc[i * DMAX] = 0.0f;
    }
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;
    __syncthreads();

    int buf = 0;
    for (int q0 = q_begin; q0 < q_end; q0 += gpb, buf ^= 1) {
        // ---- draw: thread (slot, j) draws column j of the two graphs of its slot (and nothing else).  In JAX's
        // legacy threefry layout flat elements e and e + n/2 -- samples q and q + S/2 -- are the two lanes of ONE block.
        const int q = q0 + slot;
        const bool v0 = active && q < q_end;
        const bool v1 = v0 && q + Qh < S;
        float prior0 = 0.0f, prior1 = 0.0f;
        uint32_t gm0 = 0u, gm1 = 0u;                   // hard graphs: bit i = G[i][j]
        uint32_t e_row = (uint32_t)q * dd + j;         // flat index of entry (i, j) of sample q, row by row
        // a ROLLED loop over batches of IL rows (IL independent threefry chains in flight): the entries go to the
        // hard-graph masks or to the thread's private shared-memory column, so nothing but the two priors stays live
#pragma unroll 1
        for (int i0 = 0; i0 < DMAX; i0 += IL) {
            uint32_t x0[IL], x1[IL];
            if (use_ext) {
#pragma unroll
                for (int w = 0; w < IL; ++w) {
                    const int i = i0 + w;
                    float ga = 0.0f, gb = 0.0f;
                    if (v0 && i < d && i != j) {
                        ga = p.g_ext[((size_t)m * S + q) * dd + i * d + j];
                        if (v1) gb = p.g_ext[((size_t)m * S + q + Qh) * dd + i * d + j];
                    }
                    x0[w] = __float_as_uint(ga); x1[w] = __float_as_uint(gb);
                }
            } else {
#pragma unroll
                for (int w = 0; w < IL; ++w) { x0[w] = e_row; x1[w] = e_row + half; e_row += (uint32_t)d; }
                threefry2x32_n<IL>(key.x, key.y, x0, x1);
            }
#pragma unroll
            for (int w = 0; w < IL; ++w) {
                const int i = i0 + w;
                // zero_diagonal (utils/func.py:117-125): diagonal draws are consumed and discarded.  Branch-free: the
                // entries of padding rows, of the diagonal and of slots past the chunk are computed and then zeroed
                const bool real = v0 && (FULL || i < d) && i != j;
                const bool in_tab = FULL || i < d;
                const float sa = in_tab ? sA[i * DMAX + j] : 0.0f;
                // log p(theta | G): sum g * logN(theta; mean_edge, sig_edge)   (linearGaussian.py:289)
                const float lpth = in_tab ? sLpTh[i * DMAX + j] : 0.0f;
                float ga, gb;
                if (use_ext) { ga = __uint_as_float(x0[w]); gb = __uint_as_float(x1[w]); }
                else {
                    ga = entry_from_bits<HARD>(x0[w], sa, fast_soft, p.tau);
                    gb = entry_from_bits<HARD>(x1[w], sa, fast_soft, p.tau);
                }
                ga = real ? ga : 0.0f;
                gb = (real && v1) ? gb : 0.0f;
                prior0 = fmaf(ga, lpth, prior0);
                prior1 = fmaf(gb, lpth, prior1);
                if (HARD && !use_ext) { gm0 |= (ga != 0.0f ? 1u : 0u) << i; gm1 |= (gb != 0.0f ? 1u : 0u) << i; }
                else if (active && (FULL || i < d)) { sG[i * DMAX] = ga; sG[MAT + i * DMAX] = gb; }     // own column only
            }
        }
        float* sNd = sNode + buf * (2 * gpb * NS);
        const float ta = p.tau * alpha;
        // ---- the two graphs one after the other (same code, half the registers of doing them side by side)
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            const uint32_t gm = g ? gm1 : gm0;
            float* sGg = sG + g * MAT;
            float u[DMAX];                             // u = e_j - (G o Theta)_:j
#pragma unroll
            for (int i = 0; i < DMAX; ++i) {
                float gv = 0.0f, th = 0.0f;
                if (FULL || i < d) {
                    gv = (HARD && !use_ext) ? (((gm >> i) & 1u) ? 1.0f : 0.0f) : sGg[i * DMAX];
                    th = sTh[i * DMAX + j];
                }
                u[i] = (i == j) ? 1.0f : -gv * th;
            }
            asm volatile("" ::: "memory");
            // y = Rx u (in place, ascending rows), ssq = |y|^2; Rx comes from the constant bank
            float ssq = 0.0f;
#pragma unroll
            for (int i = 0; i < DMAX; ++i) {
                float a = 0.0f;
#pragma unroll
                for (int k = i; k < DMAX; ++k) a = fmaf(R.v[rtri_off<DMAX>(i, k)], u[k], a);
                u[i] = a;
                ssq = fmaf(a, a, ssq);
            }
            // b = Rx^T y (in place, descending columns) = column j of x^T (x - x (G o Theta))
            if (MODE == MC_THETA_HARD || MODE == MC_Z_REPARAM) {
#pragma unroll
                for (int k = DMAX - 1; k >= 0; --k) {
                    float a = 0.0f;
#pragma unroll
                    for (int i = 0; i <= k; ++i) a = fmaf(R.v[rtri_off<DMAX>(i, k)], u[i], a);
                    u[k] = a;
                }
            }
            if (v0) sNd[(2 * slot + g) * NS + j] = (g ? prior1 : prior0) - 0.5f * (cst + ssq * inv_s2);
            // compiler fence: the table entries the epilogue needs are RE-LOADED below instead of being carried in
            // registers across the two mat-vecs (measured: 168 -> ~100 registers)
            asm volatile("" ::: "memory");
            // the sample's UNWEIGHTED contribution to the column, parked in the thread's private column until the
            // softmax weight of the sample is known (after the barrier)
            if (MODE != MC_LP_ONLY) {
#pragma unroll
                for (int i = 0; i < DMAX; ++i) {
                    if (FULL || i < d) {
                        float val;
                        if (MODE == MC_THETA_HARD) {
                            // d/dtheta: g * (-(theta-mu)/sig^2) + g * (x^T R)/s2          
                            const float pr = -(sTh[i * DMAX + j] - p.mean_edge) * inv_se2;
                            const float gv = (HARD && !use_ext) ? (((gm >> i) & 1u) ? 1.0f : 0.0f) : sGg[i * DMAX];
                            val = gv * fmaf(u[i], inv_s2, pr);
                        } else if (MODE == MC_Z_REPARAM) {
                            // dS = d lp/dG * tau*alpha*g(1-g)                                (App. B-4, B-6)
                            const float gv = sGg[i * DMAX];
                            const float th = sTh[i * DMAX + j] * inv_s2, lpth = sLpTh[i * DMAX + j];
                            val = fmaf(th, u[i], lpth) * ta * gv * (1.0f - gv);
                        } else {
                            // score function: weighted mean graph (App. B-2)
                            val = (HARD && !use_ext) ? (((gm >> i) & 1u) ? 1.0f : 0.0f) : sGg[i * DMAX];
                        }
                        if (active) sGg[i * DMAX] = val;
                    }
                }
            }
        }
        __syncthreads();                    // the round's only barrier: node log-probs are visible
        // ---- per-sample log-probs and softmax statistics, by EVERY warp (lane = slot*2 + which; identical results)
        float lp = -INFINITY;
        if (lane < 2 * gpb) {
            const int sl = lane >> 1, wh = lane & 1;
            const int s = q0 + sl + wh * Qh;
            if (q0 + sl < q_end && s < S) {
                lp = 0.0f;
                for (int jj = 0; jj < d; ++jj) lp += sNd[lane * NS + jj];
                if (p.lp_out && tid < 32) p.lp_out[(size_t)m * S + s] = lp;
            }
        }
        if (MODE != MC_LP_ONLY) {
            const float mx = warp_max(lp);
            const float m_new = fmaxf(m_run, mx);
            const float ex = (lp == -INFINITY) ? 0.0f : expf(lp - m_new);
            const float se = warp_sum(ex);
            const float sl_ = warp_sum(lp == -INFINITY ? 0.0f : lp);
            const float scale = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
            l_run = l_run * scale + se;
            sum_lp += sl_;
            m_run = m_new;
            const int src = active ? 2 * slot : 0;
            const float e0 = __shfl_sync(0xffffffffu, ex, src), e1 = __shfl_sync(0xffffffffu, ex, src + 1);
            if (active) {
#pragma unroll
                for (int i = 0; i < DMAX; ++i)
                    if (FULL || i < d) sAcc[i * DMAX] = fmaf(sAcc[i * DMAX], scale, fmaf(e0, sG[i * DMAX], e1 * sG[MAT + i * DMAX]));
            }
        }
        // no second barrier: the next round writes the OTHER node buffer and only private columns
    }
    if (MODE == MC_LP_ONLY) return;

    // deterministic reduction over the slots
    __syncthreads();
    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    const float inv_d2 = 1.0f / (float)d;
    for (int e = tid; e < dd; e += blockDim.x) {
        const int i = __float2int_rz(((float)e + 0.5f) * inv_d2), jj = e - i * d;
        float sum = 0.0f;
        for (int g = 0; g < gpb; ++g) sum += sAccAll[g * acc_stride + i * DMAX + jj];
        out[e] = sum;
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t mc_lin_qr_smem(int dmax, int gpb) {
    const size_t mat = (size_t)dmax * dmax;
    // tables + accumulators + private graph columns (strides padded by at most 31) + node table
    return (3 * mat + (size_t)gpb * (mat + 31) + (size_t)gpb * (2 * mat + 31) + (size_t)2 * 2 * gpb * (dmax + 1) + 8) * sizeof(float);
}

}  // namespace dibs
