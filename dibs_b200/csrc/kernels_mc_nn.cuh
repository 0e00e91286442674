// DenseNonlinearGaussian (one hidden ReLU layer) scoring + closed-form backward, fused with graph sampling.
//
// replaces: dibs/models/nonlinearGaussian.py:248-326 (log_prob_parameters, log_likelihood,
// interventional_log_joint_prob; stax Dense -> Relu -> Dense(1) per node, nonlinearGaussian.py:35-81,116-135)
// under the estimators of dibs/inference/dibs.py:395-459,488-551.
//
// thread = (sample s, node j, hidden unit h); the HP = 2^ceil(log2 H) lanes of one (s, j) share the node's
// mean through warp shuffles.  Closed forms: SURVEY App. B-7.
// theta layout per particle: W1[j,i,h] | b1[j,h] | W2[j,h] | b2[j].
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"

namespace dibs {

static inline int nn_hp(int h) { int p = 1; while (p < h) p <<= 1; return p; }

template <int DMAX, int MODE>
__global__ void __launch_bounds__(256) k_mc_nn(McParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr bool HARD = (MODE == MC_THETA_HARD || MODE == MC_Z_SCORE);
    const int d = p.d, N = p.n_obs, gpb = p.gpb, H = p.hidden, HP = p.hp;
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const int dth = d * (d * H + 2 * H + 1);
    const int oB1 = d * d * H, oW2 = oB1 + d * H, oB2 = oW2 + d * H;

    float* sA = smem;                         // [d*d]
    float* sCnt = sA + d * d;                 // [d]
    float* sNode = sCnt + d;                  // [gpb*d]
    float* sLpS = sNode + gpb * d;            // [gpb]
    float* sG = sLpS + gpb;                   // [gpb*d*d]: sG[(s_local*d + j)*d + i]
    float* sTh = smem + ((d * d + d + gpb * d + gpb + gpb * d * d + 3) & ~3);   // [dth]
    float* sBig = sTh + ((dth + 3) & ~3);
    float* sX = sBig;                         // [N*DMAX]
    float* sKeep = sX + N * DMAX;             // [N*d]

    const bool use_ext = p.g_ext != nullptr;
    const float alpha = stage_scores(p, m, sA, HARD, t);
    const float* throw_ = p.theta + (size_t)m * p.th_ld;
    for (int e = tid; e < dth; e += blockDim.x) sTh[e] = throw_[e];
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

    const int per_graph = d * HP;
    const bool active = tid < gpb * per_graph;
    const int s_local = tid / per_graph, j = (tid % per_graph) / HP, h = tid % HP;
    const bool hact = h < H;
    const int s_begin = c * p.s_per_chunk;
    const int s_end = min(p.n_samples, s_begin + p.s_per_chunk);
    const float inv_s2 = 1.0f / p.s2;
    const float inv_sp2 = 1.0f / p.sig2_edge;   // sig_param^2 (fill_mc maps the NN prior onto these fields)
    // lanes of one (s, j) group are contiguous and HP | 32, so a group never straddles a warp
    const unsigned gmask = HP == 32 ? 0xffffffffu : (((1u << HP) - 1u) << ((tid & 31) / HP * HP));

    float acc[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) acc[i] = 0.0f;
    float acc_b1 = 0.0f, acc_w2 = 0.0f, acc_b2 = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;

    for (int s0 = s_begin; s0 < s_end; s0 += gpb) {
        const int s = s0 + s_local;
        const bool valid = active && s < s_end;
        // the HP lanes of a group split the column's d draws, then share it through shared memory
        if (valid) {
            for (int i = h; i < d; i += HP) {
                float g = use_ext ? (i == j ? 0.0f : p.g_ext[(((size_t)m * p.n_samples + s) * d + i) * d + j])
                                  : graph_entry<HARD>(p, key, sA, s, i, j, d, p.tau);
                sG[(s_local * d + j) * d + i] = g;
            }
        }
        __syncwarp();
        float w[DMAX], gw[DMAX];
        float gb1 = 0.0f, gw2 = 0.0f, gb2 = 0.0f, ssq = 0.0f, prior = 0.0f;
        float w2 = 0.0f, b1 = 0.0f, b2 = 0.0f;
        if (valid) {
            w2 = hact ? sTh[oW2 + j * H + h] : 0.0f;
            b1 = hact ? sTh[oB1 + j * H + h] : 0.0f;
            b2 = sTh[oB2 + j];
#pragma unroll
            for (int i = 0; i < DMAX; ++i) {
                float wv = 0.0f;
                if (i < d && hact) {
                    float g = sG[(s_local * d + j) * d + i];
                    float w1 = sTh[(j * d + i) * H + h];
                    wv = g * w1;
                    // first-layer weights masked by G[i,j] in the prior (nonlinearGaussian.py:264-266)
                    prior = fmaf(g, norm_logpdf_pre(w1, 0.0f, p.sig2_edge, p.lognorm_edge), prior);
                }
                w[i] = wv; gw[i] = 0.0f;
            }
            if (hact) prior += norm_logpdf_pre(b1, 0.0f, p.sig2_edge, p.lognorm_edge) +
                               norm_logpdf_pre(w2, 0.0f, p.sig2_edge, p.lognorm_edge);
            if (h == 0) prior += norm_logpdf_pre(b2, 0.0f, p.sig2_edge, p.lognorm_edge);
        } else {
#pragma unroll
            for (int i = 0; i < DMAX; ++i) { w[i] = 0.0f; gw[i] = 0.0f; }
        }
        // all lanes of the warp run the loop (shuffles); invalid lanes carry zeros
        for (int n = 0; n < N; ++n) {
            const float4* xr = reinterpret_cast<const float4*>(sX + n * DMAX);
            float xv[DMAX];
#pragma unroll
            for (int q = 0; q < DMAX / 4; ++q) {
                float4 v = xr[q];
                xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
            }
            float pre = b1;
#pragma unroll
            for (int i = 0; i < DMAX; ++i) pre = fmaf(xv[i], w[i], pre);
            const float act = fmaxf(pre, 0.0f);
            float part = act * w2;
            for (int o = HP >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(gmask, part, o);
            const float mean = part + b2;
            float r = valid ? sX[n * DMAX + j] - mean : 0.0f;
            if (p.mask && valid) r *= sKeep[n * d + j];
            ssq = fmaf(r, r, ssq);
            const float delta = r * inv_s2;
            const float dpre = (pre > 0.0f) ? delta * w2 : 0.0f;
#pragma unroll
            for (int i = 0; i < DMAX; ++i) gw[i] = fmaf(xv[i], dpre, gw[i]);
            gb1 += dpre; gw2 = fmaf(delta, act, gw2); gb2 += delta;
        }
        for (int o = HP >> 1; o > 0; o >>= 1) prior += __shfl_xor_sync(gmask, prior, o);
        if (valid && h == 0) sNode[s_local * d + j] = prior - 0.5f * (sCnt[j] * p.log2pis2 + ssq * inv_s2);
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
            const bool on = valid && hact;
            const float e = on ? expf(sLpS[s_local] - m_new) : 0.0f;
#pragma unroll
            for (int i = 0; i < DMAX; ++i) {
                float val = 0.0f;
                if (on && i < d) {
                    float g = sG[(s_local * d + j) * d + i];
                    float w1 = sTh[(j * d + i) * H + h];
                    if (MODE == MC_THETA_HARD) {
                        val = g * (gw[i] - w1 * inv_sp2);                       // dW1[j,i,h]
                    } else if (MODE == MC_Z_REPARAM) {
                        float dg = gw[i] * w1 + norm_logpdf_pre(w1, 0.0f, p.sig2_edge, p.lognorm_edge);   // this h's share of d lp/dG[i,j]
                        val = dg * (p.tau * alpha) * g * (1.0f - g);
                    } else {
                        val = (h == 0) ? g : 0.0f;
                    }
                }
                acc[i] = acc[i] * scale + e * val;
            }
            if (MODE == MC_THETA_HARD) {
                acc_b1 = acc_b1 * scale + e * (gb1 - b1 * inv_sp2);
                acc_w2 = acc_w2 * scale + e * (gw2 - w2 * inv_sp2);
                acc_b2 = acc_b2 * scale + ((on && h == 0) ? e * (gb2 - b2 * inv_sp2) : 0.0f);
            }
        }
        __syncthreads();
    }
    if (MODE == MC_LP_ONLY) return;

    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    float* sRed = sBig;
    if (MODE == MC_THETA_HARD) {
        // sRed[s_local][theta index]
        if (active && hact) {
            float* r = sRed + (size_t)s_local * dth;
#pragma unroll
            for (int i = 0; i < DMAX; ++i)
                if (i < d) r[(j * d + i) * H + h] = acc[i];
            r[oB1 + j * H + h] = acc_b1;
            r[oW2 + j * H + h] = acc_w2;
            if (h == 0) r[oB2 + j] = acc_b2;
        }
        __syncthreads();
        for (int e = tid; e < dth; e += blockDim.x) {
            float sum = 0.0f;
            for (int g = 0; g < gpb; ++g) sum += sRed[(size_t)g * dth + e];
            out[e] = sum;
        }
    } else {
        // fold the hidden-unit lanes first (fixed shuffle tree), then the sample slots
#pragma unroll
        for (int i = 0; i < DMAX; ++i)
            for (int o = HP >> 1; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(gmask, acc[i], o);
        if (active && h == 0) {
#pragma unroll
            for (int i = 0; i < DMAX; ++i)
                if (i < d) sRed[(size_t)s_local * d * d + i * d + j] = acc[i];
        }
        __syncthreads();
        for (int e = tid; e < d * d; e += blockDim.x) {
            float sum = 0.0f;
            for (int g = 0; g < gpb; ++g) sum += sRed[(size_t)g * d * d + e];
            out[e] = sum;
        }
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
}

inline size_t mc_nn_smem(int d, int k, int n_obs, int gpb, int dmax, int H, int HP, bool has_mask) {
    size_t dth = (size_t)d * (d * H + 2 * H + 1);
    size_t head = ((size_t)d * d + d + (size_t)gpb * d + gpb + (size_t)gpb * d * d + 3) & ~(size_t)3;
    head += (dth + 3) & ~(size_t)3;
    size_t big = (size_t)n_obs * dmax + (has_mask ? (size_t)n_obs * d : 0);
    size_t red = (size_t)gpb * (dth > (size_t)d * d ? dth : (size_t)d * d);
    size_t zz = (size_t)2 * d * k;
    if (red > big) big = red;
    if (zz > big) big = zz;
    return (head + big + 4) * sizeof(float);
}

}  // namespace dibs
