// DenseNonlinearGaussian (one hidden ReLU layer) scoring + closed-form backward, fused with graph sampling.
//
// replaces: dibs/models/nonlinearGaussian.py:248-326 (log_prob_parameters, log_likelihood,
// interventional_log_joint_prob; stax Dense -> Relu -> Dense(1) per node, nonlinearGaussian.py:35-81,116-135)
// under the estimators of dibs/inference/dibs.py:395-459,488-551.  Closed forms: SURVEY App. B-7.
// theta layout per particle: W1[j,i,h] | b1[j,h] | W2[j,h] | b2[j].
//
// Work decomposition: CTA = (particle, chunk of sample slots); a slot is the sample pair (s, s + S/2) -- the two
// lanes of one threefry block in JAX's legacy layout -- or a single sample.  Per slot the CTA draws the graph
// entries with a flat thread mapping (both lanes of every block used), then thread = (node j, hidden unit h):
// JPW = 32 / H nodes per warp, the H lanes of a node adjacent, so the node mean is H warp shuffles.  The thread
// keeps its masked first-layer weights w[i] = G_ij W1[j,i,h] and their gradient in registers as PACKED PAIRS and
// streams the N observations from shared memory as 128-bit broadcasts, two observations per iteration:
// 5 LDS.128 + 20 FFMA2 per observation at d = 20.
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"
#include "kernels_mc_lin_qr.cuh"   // entry_from_bits
#include <type_traits>

namespace dibs {

static inline int nn_hp(int h) { int p = 1; while (p < h) p <<= 1; return p; }
static inline int nn_jpw(int h) { return 32 / h; }                                 // nodes per warp
static inline int nn_threads(int d, int h) { return ((d + nn_jpw(h) - 1) / nn_jpw(h)) * 32; }

// CTA size = 32 * ceil(d / (32 / H)); the bound below is what the launcher checks against
template <int DMAX> constexpr int nn_max_threads() { return DMAX <= 32 ? 256 : (DMAX <= 64 ? 384 : 768); }

// HC: compile-time hidden-layer width (the node-mean shuffles unroll), 0 = runtime width
template <int DMAX, int MODE, int HC>
__global__ void __launch_bounds__(nn_max_threads<DMAX>()) k_mc_nn(const __grid_constant__ McParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr bool HARD = (MODE == MC_THETA_HARD || MODE == MC_Z_SCORE);
    constexpr int NP = DMAX / 2;                  // packed pairs per row
    const int d = p.d, dd = d * d, N = p.n_obs, H = p.hidden;
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = p.st ? p.st->t : p.t_override;
    const int S = p.n_samples;
    const bool paired = p.paired != 0;
    const int per = paired ? 2 : 1;
    const int Qh = paired ? (S >> 1) : S;
    const int dth = d * (d * H + 2 * H + 1);
    const int oB1 = dd * H, oW2 = oB1 + d * H, oB2 = oW2 + d * H;
    const int JPW = 32 / H;

    float* sA = smem;                                    // [dd] hard: P_ij; soft tau==1: exp(-alpha s_ij); soft: alpha s_ij
    float* sG = sA + ((dd + 3) & ~3);                    // [2][dd] graph entries of the slot, [i*d + j]
    float* sNode = sG + ((2 * dd + 3) & ~3);             // [d] per-node log-probs of the current sample
    float* sCnt = sNode + ((d + 3) & ~3);                // [d] unmasked observation count per node
    float* sTh = sCnt + ((d + 3) & ~3);                  // [dth]
    float* sX = sTh + ((dth + 3) & ~3);                  // [N][DMAX] zero-padded rows
    float* sKeep = sX + (size_t)N * DMAX;                // [N*d] (only if mask)

    const bool use_ext = p.g_ext != nullptr;
    const bool fast_soft = !HARD && !use_ext && p.tau == 1.0f;
    const float alpha = p.alpha_linear * (float)t;       // dibs.py:70
    {
        const float* srow = p.scores ? p.scores + (size_t)m * dd : nullptr;
        for (int e = tid; e < dd; e += blockDim.x) {
            const int i = e / d, j = e - i * d;
            const float a = srow ? alpha * srow[e] : 0.0f;
            float sa;
            if (HARD) sa = (i == j) ? 0.0f : sigmoidf_ref(a);             // edge_probs (dibs.py:168-184)
            else sa = fast_soft ? expf(-a) : a;
            sA[e] = sa;
        }
        const float* throw_ = p.theta + (size_t)m * p.th_ld;
        for (int e = tid; e < dth; e += blockDim.x) sTh[e] = throw_[e];
        for (int e = tid; e < N * DMAX; e += blockDim.x) {
            const int n = e / DMAX, i = e - n * DMAX;
            sX[e] = i < d ? p.x[n * d + i] : 0.0f;
        }
        if (p.mask)
            for (int e = tid; e < N * d; e += blockDim.x) sKeep[e] = p.mask[e] ? 0.0f : 1.0f;
    }
    __syncthreads();
    if (tid < d) {
        float cnt = (float)N;
        if (p.mask) { cnt = 0.0f; for (int n = 0; n < N; ++n) cnt += sKeep[n * d + tid]; }
        sCnt[tid] = cnt;
    }
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);

    // thread -> (node j, hidden unit h)
    const int jl = lane / H, h = lane - jl * H;
    const int j = warp * JPW + jl;
    const bool on = jl < JPW && j < d;                   // lanes beyond JPW*H and nodes beyond d idle (carry zeros)
    const int jj = on ? j : 0;
    const int base_lane = jl * H;                        // first lane of this node's group
    const float inv_s2 = 1.0f / p.s2;
    const float inv_sp2 = 1.0f / p.sig2_edge;            // sig_param^2 (fill_mc maps the NN prior onto these fields)
    const float* xcol = sX + jj;                         // x[n][j] = xcol[n * DMAX]
    const float* keepcol = p.mask ? sKeep + jj : nullptr;
    const float w2 = on ? sTh[oW2 + jj * H + h] : 0.0f;
    const float b1 = on ? sTh[oB1 + jj * H + h] : 0.0f;
    const float b2 = on ? sTh[oB2 + jj] : 0.0f;
    // parameter prior terms that do not depend on the graph (nonlinearGaussian.py:257-276)
    float prior_fix = 0.0f;
    if (on) {
        prior_fix = norm_logpdf_pre(b1, 0.0f, p.sig2_edge, p.lognorm_edge) + norm_logpdf_pre(w2, 0.0f, p.sig2_edge, p.lognorm_edge);
        if (h == 0) prior_fix += norm_logpdf_pre(b2, 0.0f, p.sig2_edge, p.lognorm_edge);
    }

    float acc[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) acc[i] = 0.0f;
    float acc_b1 = 0.0f, acc_w2 = 0.0f, acc_b2 = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;

    const int q_begin = c * p.s_per_chunk;
    const int q_end = min(Qh, q_begin + p.s_per_chunk);
    const uint32_t n_total = (uint32_t)S * dd;
    const uint32_t half = n_total >> 1;
    __syncthreads();

    for (int q = q_begin; q < q_end; ++q) {
        // ---- draw the slot's graph entries (flat mapping, coalesced shared-memory writes)
        for (int e = tid; e < dd; e += blockDim.x) {
            const int i = e / d, je = e - i * d;
            float g0 = 0.0f, g1 = 0.0f;
            if (i != je) {                                // zero_diagonal: diagonal draws are discarded
                if (use_ext) {
                    g0 = p.g_ext[((size_t)m * S + q) * dd + e];
                } else if (paired) {
                    const uint32_t e0 = (uint32_t)q * dd + e;
                    const uint2 bits = threefry2x32(key.x, key.y, e0, e0 + half);
                    const float sa = sA[e];
                    g0 = entry_from_bits<HARD>(bits.x, sa, fast_soft, p.tau);
                    g1 = entry_from_bits<HARD>(bits.y, sa, fast_soft, p.tau);
                } else {
                    const uint32_t bits = jax_bits(key, (uint32_t)q * dd + e, n_total, p.partitionable);
                    g0 = entry_from_bits<HARD>(bits, sA[e], fast_soft, p.tau);
                }
            }
            sG[e] = g0; sG[dd + e] = g1;
        }
        __syncthreads();
#pragma unroll 1
        for (int hs = 0; hs < per; ++hs) {
            const int s = q + hs * Qh;
            const float* G = sG + hs * dd;
            // masked first-layer weights of (j, h) and the graph-dependent prior term
            f32x2 w[NP], gw[NP];
            float prior = prior_fix;
#pragma unroll
            for (int ip = 0; ip < NP; ++ip) {
                float wv[2] = {0.0f, 0.0f};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = 2 * ip + u;
                    if (on && i < d) {
                        const float g = G[i * d + jj];
                        const float w1 = sTh[(jj * d + i) * H + h];
                        wv[u] = g * w1;
                        // first-layer weights masked by G[i,j] in the prior (nonlinearGaussian.py:264-266)
                        prior = fmaf(g, norm_logpdf_pre(w1, 0.0f, p.sig2_edge, p.lognorm_edge), prior);
                    }
                }
                w[ip] = pack2(wv[0], wv[1]);
                gw[ip] = 0ull;
            }
            float gb1 = 0.0f, gw2 = 0.0f, gb2 = 0.0f, ssq = 0.0f;
            // ---- forward + backward over the observations, two per iteration (independent chains)
            const float onf = on ? 1.0f : 0.0f;
            // the activation is a compile-time constant of the observation loop (the loop is instantiated once per
            // activation and selected outside it: a run-time switch inside cost 35 % on the ReLU path)
            auto body = [&](int n0, auto two_c, auto act_c) {
                constexpr bool TWO = decltype(two_c)::value;
                constexpr int actk = decltype(act_c)::value;
                const ulonglong2* xa = reinterpret_cast<const ulonglong2*>(sX + (size_t)n0 * DMAX);
                const ulonglong2* xb = reinterpret_cast<const ulonglong2*>(sX + (size_t)(TWO ? n0 + 1 : n0) * DMAX);
                f32x2 xva[NP], xvb[NP];
#pragma unroll
                for (int qd = 0; qd < DMAX / 4; ++qd) {
                    const ulonglong2 va = xa[qd];
                    xva[2 * qd] = va.x; xva[2 * qd + 1] = va.y;
                    if (TWO) { const ulonglong2 vb = xb[qd]; xvb[2 * qd] = vb.x; xvb[2 * qd + 1] = vb.y; }
                }
                // four independent accumulation chains (even / odd pairs of each observation)
                f32x2 pa0 = pack2(b1, 0.0f), pa1 = 0ull, pb0 = pa0, pb1 = 0ull;
#pragma unroll
                for (int ip = 0; ip < NP; ip += 2) {
                    pa0 = fma2(xva[ip], w[ip], pa0);
                    if (ip + 1 < NP) pa1 = fma2(xva[ip + 1], w[ip + 1], pa1);
                    if (TWO) {
                        pb0 = fma2(xvb[ip], w[ip], pb0);
                        if (ip + 1 < NP) pb1 = fma2(xvb[ip + 1], w[ip + 1], pb1);
                    }
                }
                const f32x2 pa = add2(pa0, pa1), pb = add2(pb0, pb1);
                const float pre_a = lo2(pa) + hi2(pa), pre_b = lo2(pb) + hi2(pb);
                float act_a, act_b, dact_a, dact_b;          // activation and its derivative (stax Relu / Tanh / Sigmoid / LeakyRelu)
                if (actk == 0) {
                    act_a = fmaxf(pre_a, 0.0f); act_b = fmaxf(pre_b, 0.0f);
                    dact_a = pre_a > 0.0f ? 1.0f : 0.0f; dact_b = pre_b > 0.0f ? 1.0f : 0.0f;
                } else if (actk == 1) {
                    act_a = tanhf(pre_a); act_b = tanhf(pre_b);
                    dact_a = 1.0f - act_a * act_a; dact_b = 1.0f - act_b * act_b;
                } else if (actk == 2) {
                    act_a = sigmoidf_ref(pre_a); act_b = sigmoidf_ref(pre_b);
                    dact_a = act_a * (1.0f - act_a); dact_b = act_b * (1.0f - act_b);
                } else {
                    act_a = pre_a >= 0.0f ? pre_a : 0.01f * pre_a; act_b = pre_b >= 0.0f ? pre_b : 0.01f * pre_b;
                    dact_a = pre_a >= 0.0f ? 1.0f : 0.01f; dact_b = pre_b >= 0.0f ? 1.0f : 0.01f;
                }
                const float part_a = act_a * w2, part_b = act_b * w2;
                float mean_a = b2, mean_b = b2;
                if (HC > 0) {
#pragma unroll
                    for (int hh = 0; hh < (HC > 0 ? HC : 1); ++hh) {
                        mean_a += __shfl_sync(0xffffffffu, part_a, base_lane + hh);
                        if (TWO) mean_b += __shfl_sync(0xffffffffu, part_b, base_lane + hh);
                    }
                } else {
                    for (int hh = 0; hh < H; ++hh) {
                        mean_a += __shfl_sync(0xffffffffu, part_a, base_lane + hh);
                        if (TWO) mean_b += __shfl_sync(0xffffffffu, part_b, base_lane + hh);
                    }
                }
                float ra = onf * (xcol[(size_t)n0 * DMAX] - mean_a);
                float rb = TWO ? onf * (xcol[(size_t)(n0 + 1) * DMAX] - mean_b) : 0.0f;
                if (keepcol) { ra *= keepcol[n0 * d]; if (TWO) rb *= keepcol[(n0 + 1) * d]; }
                ssq = fmaf(ra, ra, ssq);
                if (TWO) ssq = fmaf(rb, rb, ssq);
                const float da = ra * inv_s2, db = rb * inv_s2;
                const float dpa = (actk == 0) ? (pre_a > 0.0f ? da * w2 : 0.0f) : da * w2 * dact_a;
                const float dpb = !TWO ? 0.0f : ((actk == 0) ? (pre_b > 0.0f ? db * w2 : 0.0f) : db * w2 * dact_b);
                const f32x2 dpa2 = pack2(dpa, dpa), dpb2 = pack2(dpb, dpb);
#pragma unroll
                for (int ip = 0; ip < NP; ++ip) {
                    gw[ip] = fma2(xva[ip], dpa2, gw[ip]);
                    if (TWO) gw[ip] = fma2(xvb[ip], dpb2, gw[ip]);
                }
                gb1 += dpa; gw2 = fmaf(da, act_a, gw2); gb2 += da;
                if (TWO) { gb1 += dpb; gw2 = fmaf(db, act_b, gw2); gb2 += db; }
            };
            auto run = [&](auto act_c) {
                for (int n0 = 0; n0 + 1 < N; n0 += 2) body(n0, std::true_type{}, act_c);
                if (N & 1) body(N - 1, std::false_type{}, act_c);
            };
            switch (p.activation) {
                case 1 /* DIBS_ACT_TANH */: run(std::integral_constant<int, 1>{}); break;
                case 2 /* DIBS_ACT_SIGMOID */: run(std::integral_constant<int, 2>{}); break;
                case 3 /* DIBS_ACT_LEAKYRELU */: run(std::integral_constant<int, 3>{}); break;
                default: run(std::integral_constant<int, 0>{}); break;
            }
            // node log-prob: prior terms of the group's lanes + Gaussian likelihood
            float psum = 0.0f;
            for (int hh = 0; hh < H; ++hh) psum += __shfl_sync(0xffffffffu, prior, base_lane + hh);
            if (on && h == 0) sNode[j] = psum - 0.5f * (sCnt[j] * p.log2pis2 + ssq * inv_s2);
            __syncthreads();
            float lp = 0.0f;
            for (int q2 = 0; q2 < d; ++q2) lp += sNode[q2];                  // every thread: same fixed order
            if (p.lp_out && tid == 0) p.lp_out[(size_t)m * S + s] = lp;
            if (MODE != MC_LP_ONLY) {
                const float m_new = fmaxf(m_run, lp);
                const float scale = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
                const float e = expf(lp - m_new);
                l_run = l_run * scale + e;
                sum_lp += lp;
                m_run = m_new;
                const float ew = on ? e : 0.0f;
#pragma unroll
                for (int i = 0; i < DMAX; ++i) {
                    float val = 0.0f;
                    if (on && i < d) {
                        const float g = G[i * d + jj];
                        const float w1 = sTh[(jj * d + i) * H + h];
                        const float gwi = (i & 1) ? hi2(gw[i >> 1]) : lo2(gw[i >> 1]);
                        if (MODE == MC_THETA_HARD) {
                            val = g * (gwi - w1 * inv_sp2);                      // dW1[j,i,h]
                        } else if (MODE == MC_Z_REPARAM) {
                            // this h's share of d lp/dG[i,j], through the sigmoid (App. B-4, B-7)
                            const float dg = gwi * w1 + norm_logpdf_pre(w1, 0.0f, p.sig2_edge, p.lognorm_edge);
                            val = dg * (p.tau * alpha) * g * (1.0f - g);
                        } else {
                            val = (h == 0) ? g : 0.0f;
                        }
                    }
                    acc[i] = fmaf(acc[i], scale, ew * val);
                }
                if (MODE == MC_THETA_HARD) {
                    acc_b1 = fmaf(acc_b1, scale, ew * (gb1 - b1 * inv_sp2));
                    acc_w2 = fmaf(acc_w2, scale, ew * (gw2 - w2 * inv_sp2));
                    acc_b2 = fmaf(acc_b2, scale, (h == 0) ? ew * (gb2 - b2 * inv_sp2) : 0.0f);
                }
            }
            __syncthreads();                                                    // sNode is rewritten by the next sample
        }
    }
    if (MODE == MC_LP_ONLY) return;

    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    if (MODE == MC_THETA_HARD) {
        // one thread per (j, h): the chunk's sums go straight to the partial buffer
        if (on) {
#pragma unroll
            for (int i = 0; i < DMAX; ++i)
                if (i < d) out[(j * d + i) * H + h] = acc[i];
            out[oB1 + j * H + h] = acc_b1;
            out[oW2 + j * H + h] = acc_w2;
            if (h == 0) out[oB2 + j] = acc_b2;
        }
    } else {
        // dS[i][j] = sum over the hidden units of the node (fixed order)
#pragma unroll
        for (int i = 0; i < DMAX; ++i) {
            float sum = 0.0f;
            for (int hh = 0; hh < H; ++hh) sum += __shfl_sync(0xffffffffu, acc[i], base_lane + hh);
            if (on && h == 0 && i < d) out[i * d + j] = sum;
        }
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t mc_nn_smem(int d, int n_obs, int dmax, int H, bool has_mask) {
    const size_t dd = (size_t)d * d;
    const size_t dth = (size_t)d * (d * H + 2 * H + 1);
    size_t f = ((dd + 3) & ~(size_t)3) + ((2 * dd + 3) & ~(size_t)3) + 2 * (((size_t)d + 3) & ~(size_t)3) + ((dth + 3) & ~(size_t)3);
    f += (size_t)n_obs * dmax + (has_mask ? (size_t)n_obs * d : 0);
    return (f + 4) * sizeof(float);
}

}  // namespace dibs
