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
// Work decomposition: CTA = (particle, chunk of sample PAIRS).  Per round the CTA first draws the entries of
// gpb graph pairs into shared memory with a flat thread mapping, then thread = (pair q, node j) owns the
// columns j of the two graphs s = q and s = q + S/2: in JAX's legacy threefry layout those two draws are the
// two lanes of ONE threefry block (counter pair (e, e + n/2)), so no random bits are thrown away.  (Host side:
// odd S or the partitionable PRNG layout fall back to k_mc_lingauss, which draws one entry per block.)
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
    if (fast_soft) return __fdividef(u, fmaf(1.0f - u, sa, u));
    return sigmoidf_ref(tau * ((logf(u) - log1pf(-u)) + sa));
}

template <int DMAX, int MODE>
__global__ void __launch_bounds__(192, (DMAX <= 20 ? 2 : 1)) k_mc_lin_qr(const __grid_constant__ McParams p, const __grid_constant__ RTri<DMAX> R) {
    extern __shared__ __align__(16) float smem[];
    constexpr bool HARD = (MODE == MC_THETA_HARD || MODE == MC_Z_SCORE);
    constexpr int MAT = DMAX * DMAX;                   // all per-(i,j) tables use the compile-time stride DMAX
    const int d = p.d, dd = d * d, gpb = p.gpb;        // gpb = sample pairs ("slots") per round
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const int S = p.n_samples, Qh = (S + 1) >> 1;      // slot q holds samples q and q + Qh

    float* sA = smem;                       // [MAT] hard: P_ij; soft tau==1: exp(-alpha s_ij); soft: alpha s_ij
    float* sTh = sA + MAT;                  // [MAT] theta_ij
    float* sLpTh = sTh + MAT;               // [MAT] logN(theta_ij; mean_edge, sig_edge)
    float* sAccAll = sLpTh + MAT;           // [gpb][MAT] softmax-weighted running sums, one private column per thread
    float* sGall = sAccAll + gpb * MAT;     // [gpb][2][MAT] graph entries of the current round
    float* sNode = sGall + 2 * gpb * MAT;   // [2*gpb][DMAX] per-node log-probs
    float* sLpS = sNode + 2 * gpb * DMAX;   // [2*gpb] per-sample log-probs
    float* sStat = sLpS + 2 * gpb;          // [4] round max, sum exp, sum lp

    const bool use_ext = p.g_ext != nullptr;
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
    const float inv_dd = 1.0f / (float)dd, inv_d = 1.0f / (float)d;

    float* sAcc = sAccAll + slot * MAT + j;            // sAcc[i*DMAX]
    if (active && MODE != MC_LP_ONLY) {
#pragma unroll
        for (int i = 0; i < DMAX; ++i) sAcc[i * DMAX] = 0.0f;
    }
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;
    __syncthreads();

    for (int q0 = q_begin; q0 < q_end; q0 += gpb) {
        // ---- phase 1: all threads draw the round's graph entries into shared memory (flat, coalesced)
        const int n_draw = gpb * dd;
        constexpr int IL = 4;                  // independent threefry chains per thread
        for (int idx0 = tid; idx0 < n_draw; idx0 += IL * blockDim.x) {
            uint32_t x0[IL], x1[IL];
            int off[IL];                       // destination in sGall, or -1
            float sa[IL];
            bool draw[IL];
#pragma unroll
            for (int u = 0; u < IL; ++u) {
                const int idx = idx0 + u * blockDim.x;
                // idx -> (slot, i, j) without integer division (idx < 2^14: the float quotients are exact after +0.5)
                const int sl = __float2int_rz(((float)idx + 0.5f) * inv_dd), ij = idx - sl * dd;
                const int i = __float2int_rz(((float)ij + 0.5f) * inv_d), jj = ij - i * d;
                const int qq = q0 + sl;
                const bool in = idx < n_draw;
                off[u] = in ? (2 * sl) * MAT + i * DMAX + jj : -1;
                // zero_diagonal (utils/func.py:117-125): diagonal draws are discarded
                draw[u] = in && qq < q_end && i != jj;
                // legacy threefry layout, S even: flat elements e0 and e0 + n/2 are the two lanes of one block
                const uint32_t e0 = (uint32_t)qq * dd + ij;
                x0[u] = e0; x1[u] = e0 + half;
                sa[u] = draw[u] ? sA[i * DMAX + jj] : 0.0f;
                if (use_ext) {
                    float ga = 0.0f, gb = 0.0f;
                    if (draw[u]) {
                        ga = p.g_ext[((size_t)m * S + qq) * dd + ij];
                        if (qq + Qh < S) gb = p.g_ext[((size_t)m * S + qq + Qh) * dd + ij];
                    }
                    x0[u] = __float_as_uint(ga); x1[u] = __float_as_uint(gb);
                }
            }
            if (!use_ext) threefry2x32_n<IL>(key.x, key.y, x0, x1);
#pragma unroll
            for (int u = 0; u < IL; ++u) {
                if (off[u] < 0) continue;
                float ga, gb;
                if (use_ext) { ga = __uint_as_float(x0[u]); gb = __uint_as_float(x1[u]); }
                else {
                    ga = draw[u] ? entry_from_bits<HARD>(x0[u], sa[u], fast_soft, p.tau) : 0.0f;
                    gb = draw[u] ? entry_from_bits<HARD>(x1[u], sa[u], fast_soft, p.tau) : 0.0f;
                }
                sGall[off[u]] = ga; sGall[off[u] + MAT] = gb;
            }
        }
        __syncthreads();
        // ---- phase 2: thread (slot, j) owns column j of the two graphs of its slot
        const int q = q0 + slot;
        const bool v0 = active && q < q_end;
        const bool v1 = v0 && q + Qh < S;
        const float* sG0 = sGall + (2 * slot) * MAT + j;     // sG0[i*DMAX] = G_s0[i][j]
        const float* sG1 = sG0 + MAT;
        float u0[DMAX], u1[DMAX];
        float prior0 = 0.0f, prior1 = 0.0f;
#pragma unroll
        for (int i = 0; i < DMAX; ++i) {
            float ga = 0.0f, gb = 0.0f, th = 0.0f;
            if (v0 && i < d) {
                ga = sG0[i * DMAX]; gb = sG1[i * DMAX];
                th = sTh[i * DMAX + j];
                // log p(theta | G): sum g * logN(theta; mean_edge, sig_edge)   (linearGaussian.py:289)
                const float lpth = sLpTh[i * DMAX + j];
                prior0 = fmaf(ga, lpth, prior0);
                prior1 = fmaf(gb, lpth, prior1);
            }
            // u = e_j - (G o Theta)_:j
            u0[i] = (i == j) ? 1.0f : -ga * th;
            u1[i] = (i == j) ? 1.0f : -gb * th;
        }
#ifdef DIBS_QR_FFMA2
        // experiment (compile with -DDIBS_QR_FFMA2): the two graphs of the slot as one packed pair per row, so each
        // mat-vec step is ONE FFMA2 with the Rx operand broadcast -- bit-identical results, half the FMA issue slots
        f32x2 uu[DMAX];
#pragma unroll
        for (int i = 0; i < DMAX; ++i) uu[i] = pack2(u0[i], u1[i]);
        f32x2 ssq2 = 0ull;
#pragma unroll
        for (int i = 0; i < DMAX; ++i) {
            f32x2 a = 0ull;
#pragma unroll
            for (int k = i; k < DMAX; ++k) {
                const float r = R.v[rtri_off<DMAX>(i, k)];
                a = fma2(pack2(r, r), uu[k], a);
            }
            uu[i] = a;
            ssq2 = fma2(a, a, ssq2);
        }
        float ssq0 = lo2(ssq2), ssq1 = hi2(ssq2);
        if (MODE == MC_THETA_HARD || MODE == MC_Z_REPARAM) {
#pragma unroll
            for (int k = DMAX - 1; k >= 0; --k) {
                f32x2 a = 0ull;
#pragma unroll
                for (int i = 0; i <= k; ++i) {
                    const float r = R.v[rtri_off<DMAX>(i, k)];
                    a = fma2(pack2(r, r), uu[i], a);
                }
                uu[k] = a;
            }
        }
#pragma unroll
        for (int i = 0; i < DMAX; ++i) { u0[i] = lo2(uu[i]); u1[i] = hi2(uu[i]); }
#else
        // y = Rx u (in place, ascending rows), ssq = |y|^2
        float ssq0 = 0.0f, ssq1 = 0.0f;
#pragma unroll
        for (int i = 0; i < DMAX; ++i) {
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int k = i; k < DMAX; ++k) {
                const float r = R.v[rtri_off<DMAX>(i, k)];
                a0 = fmaf(r, u0[k], a0);
                a1 = fmaf(r, u1[k], a1);
            }
            u0[i] = a0; u1[i] = a1;
            ssq0 = fmaf(a0, a0, ssq0);
            ssq1 = fmaf(a1, a1, ssq1);
        }
        // b = Rx^T y (in place, descending columns) = column j of x^T (x - x (G o Theta))
        if (MODE == MC_THETA_HARD || MODE == MC_Z_REPARAM) {
#pragma unroll
            for (int k = DMAX - 1; k >= 0; --k) {
                float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                for (int i = 0; i <= k; ++i) {
                    const float r = R.v[rtri_off<DMAX>(i, k)];
                    a0 = fmaf(r, u0[i], a0);
                    a1 = fmaf(r, u1[i], a1);
                }
                u0[k] = a0; u1[k] = a1;
            }
        }
#endif
        if (v0) {
            const float cst = (float)p.n_obs * p.log2pis2;
            sNode[(2 * slot) * DMAX + j] = prior0 - 0.5f * (cst + ssq0 * inv_s2);
            sNode[(2 * slot + 1) * DMAX + j] = prior1 - 0.5f * (cst + ssq1 * inv_s2);
        }
        __syncthreads();
        // per-sample log-probs and the round's softmax statistics, by warp 0 (2*gpb <= 32 entries)
        if (tid < 32) {
            float lp = -INFINITY;
            if (tid < 2 * gpb) {
                const int sl = tid >> 1, wh = tid & 1;
                const int s = q0 + sl + wh * Qh;
                if (q0 + sl < q_end && s < S) {
                    lp = 0.0f;
                    for (int jj = 0; jj < d; ++jj) lp += sNode[tid * DMAX + jj];
                    if (p.lp_out) p.lp_out[(size_t)m * S + s] = lp;
                }
                sLpS[tid] = lp;
            }
            if (MODE != MC_LP_ONLY) {
                const float mx = warp_max(lp);
                const float ex = (lp == -INFINITY) ? 0.0f : expf(lp - fmaxf(mx, m_run));
                const float se = warp_sum(ex);
                const float sl_ = warp_sum(lp == -INFINITY ? 0.0f : lp);
                if (tid == 0) { sStat[0] = mx; sStat[1] = se; sStat[2] = sl_; }
            }
        }
        __syncthreads();
        if (MODE != MC_LP_ONLY) {
            const float m_new = fmaxf(m_run, sStat[0]);
            const float scale = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
            l_run = l_run * scale + sStat[1];
            sum_lp += sStat[2];
            m_run = m_new;
            const float e0 = v0 ? expf(sLpS[2 * slot] - m_new) : 0.0f;
            const float e1 = v1 ? expf(sLpS[2 * slot + 1] - m_new) : 0.0f;
            if (v0) {
#pragma unroll
                for (int i = 0; i < DMAX; ++i) {
                    if (i < d) {
                        const float ga = sG0[i * DMAX], gb = sG1[i * DMAX];
                        float val0, val1;
                        if (MODE == MC_THETA_HARD) {
                            // d/dtheta: g * (-(theta-mu)/sig^2) + g * (x^T R)/s2          
                            const float pr = -(sTh[i * DMAX + j] - p.mean_edge) * inv_se2;
                            val0 = ga * fmaf(u0[i], inv_s2, pr);
                            val1 = gb * fmaf(u1[i], inv_s2, pr);
                        } else if (MODE == MC_Z_REPARAM) {
                            // dS = d lp/dG * tau*alpha*g(1-g)                                (App. B-4, B-6)
                            const float th = sTh[i * DMAX + j] * inv_s2, lpth = sLpTh[i * DMAX + j];
                            val0 = fmaf(th, u0[i], lpth) * (p.tau * alpha) * ga * (1.0f - ga);
                            val1 = fmaf(th, u1[i], lpth) * (p.tau * alpha) * gb * (1.0f - gb);
                        } else {
                            val0 = ga; val1 = gb;   // score function: weighted mean graph (App. B-2)
                        }
                        sAcc[i * DMAX] = fmaf(sAcc[i * DMAX], scale, fmaf(e0, val0, e1 * val1));
                    }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == MC_LP_ONLY) return;

    // deterministic reduction over the slots
    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    const float inv_d2 = 1.0f / (float)d;
    for (int e = tid; e < dd; e += blockDim.x) {
        const int i = __float2int_rz(((float)e + 0.5f) * inv_d2), jj = e - i * d;
        float sum = 0.0f;
        for (int g = 0; g < gpb; ++g) sum += sAccAll[g * MAT + i * DMAX + jj];
        out[e] = sum;
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t mc_lin_qr_smem(int dmax, int gpb) {
    size_t mat = (size_t)dmax * dmax;
    return (3 * mat + (size_t)gpb * mat + 2 * (size_t)gpb * mat + 2 * (size_t)gpb * dmax + 2 * gpb + 8) * sizeof(float);
}

}  // namespace dibs
