// BGe marginal likelihood scoring of sampled hard graphs, fused with Bernoulli sampling and the
// score-function aggregation (MarginalDiBS hot path).
//
// replaces: dibs/models/linearGaussian.py:63-170 (BGe._log_marginal_likelihood_single, log_marginal_likelihood,
// interventional_log_marginal_prob) and dibs/utils/func.py:128-145 (_slogdet_jax).
//
// The reference pads the parent sub-matrix of R with an identity and runs two d x d LU factorisations per
// node per graph, recomputing R (which depends only on the data) every time.  Here R_j is precomputed once
// per dibs_set_data in fp64, and each (graph, node) task does ONE Cholesky of the gathered
// (n_parents+1) x (n_parents+1) block ordered parents-first, which yields both determinants
// (SURVEY App. B-10):  logdet R[P,P] = sum log pivots,  logdet R[P+j,P+j] = logdet R[P,P] + log(last pivot).
// Factorisations run in fp64: conditioning of R (~N var(x) / small_t) would otherwise leak ~1e-3 absolute
// error into log-probs that feed a softmax.
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"

namespace dibs {

// thread = (sample s, node j): Bernoulli parents of j, one fp64 Cholesky in local memory, node score
template <int DMAX, int MODE>
__global__ void __launch_bounds__(256) k_mc_bge(McParams p) {
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, gpb = p.gpb;
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;

    float* sA = smem;                  // [d*d] edge probabilities
    float* sNode = sA + d * d;         // [gpb*d]
    float* sLpS = sNode + gpb * d;     // [gpb]
    float* sBig = smem + ((d * d + gpb * d + gpb + 3) & ~3);

    const bool use_ext = p.g_ext != nullptr;
    if (!use_ext) stage_scores(p, m, sA, true, t);
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);
    __syncthreads();

    const bool active = tid < gpb * d;
    const int s_local = tid / d, j = tid % d;
    const int s_begin = c * p.s_per_chunk;
    const int s_end = min(p.n_samples, s_begin + p.s_per_chunk);

    float acc[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;

    for (int s0 = s_begin; s0 < s_end; s0 += gpb) {
        const int s = s0 + s_local;
        const bool valid = active && s < s_end;
        unsigned long long par = 0ull;
        if (valid) {
            unsigned char idx[DMAX + 1];
            int l = 0;
            for (int i = 0; i < d; ++i) {
                float g = use_ext ? (i == j ? 0.0f : p.g_ext[(((size_t)m * p.n_samples + s) * d + i) * d + j])
                                  : graph_entry<true>(p, key, sA, s, i, j, d, 0.0f);
                if (g > 0.5f) { par |= 1ull << i; idx[l++] = (unsigned char)i; }
            }
            idx[l] = (unsigned char)j;
            float node_lp = 0.0f;
            if (p.bge_coef[2 * j + 1] != 0.0f) {
                const double* R = p.bge_r + (size_t)j * p.bge_r_stride;
                double L[(DMAX + 1) * (DMAX + 2) / 2];
                double logdet = 0.0, schur = 1.0;
                for (int r = 0; r <= l; ++r) {
                    const double* Rrow = R + (size_t)idx[r] * d;
                    const int ro = r * (r + 1) / 2;
                    for (int cc = 0; cc <= r; ++cc) {
                        const int co = cc * (cc + 1) / 2;
                        double sum = Rrow[idx[cc]];
                        for (int kk = 0; kk < cc; ++kk) sum -= L[ro + kk] * L[co + kk];
                        if (cc < r) L[ro + cc] = sum * L[co + cc];         // diagonal slots hold 1/L_cc
                        else if (r < l) { logdet += log(sum); L[ro + r] = rsqrt(sum); }
                        else schur = sum;
                    }
                }
                // 0.5 (A) logdet R_PP - 0.5 (A+1) logdet R_(P+j)(P+j),  A = N_j + alpha_lambd - d + l  (linearGaussian.py:109-115)
                const double a1 = (double)p.bge_coef[2 * j] + (double)l + 1.0;
                node_lp = (float)((double)p.bge_table[j * (d + 1) + l] - 0.5 * logdet - 0.5 * a1 * log(schur));
            }
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
            for (int i = 0; i < DMAX; ++i) acc[i] = acc[i] * scale + (((par >> i) & 1ull) ? e : 0.0f);
        }
        __syncthreads();
    }
    if (MODE == MC_LP_ONLY) return;
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
}

inline size_t mc_bge_smem(int d, int k, int gpb) {
    size_t head = ((size_t)d * d + (size_t)gpb * d + gpb + 3) & ~(size_t)3;
    size_t big = (size_t)gpb * d * d;
    if ((size_t)2 * d * k > big) big = (size_t)2 * d * k;
    return (head + big + 4) * sizeof(float);
}

}  // namespace dibs
