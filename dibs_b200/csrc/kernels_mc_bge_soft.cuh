// BGe marginal likelihood on SOFT graphs: MarginalDiBS with grad_estimator_z = 'reparam' (n_vars <= 32).
//
// replaces: dibs/models/linearGaussian.py:63-144 evaluated on Gumbel-soft adjacency matrices (README.md:88-90 lists
// BGe + reparam as supported) under dibs/inference/dibs.py:395-459 (grad_z_likelihood_gumbel), including
// _slogdet_jax (dibs/utils/func.py:128-145): the reference masks R with the OUTER PRODUCT of the soft parent
// indicator, adds the identity elsewhere, takes an LU log-determinant, and lets `n_parents = g.sum(0)` be real-valued
// inside gammaln -- then autodiffs all of it.  Closed form used here, per node j with p = g[:, j] (p_j = 0), l = sum p,
// D = diag(p):
//   A = D R D + I - D^2                      (= mask * R + (1 - mask) * I, symmetric positive definite)
//   logdet B = logdet A + log s,  s = R_jj - u^T A^-1 u,  u_a = p_a R_aj      (B: the same with p_j = 1)
//   d logdet A / dp_i = 2 sum_b (A^-1)_ib p_b (R_ib - delta_ib)
//   d s / dp_i        = -2 R_ij w_i + 2 w_i sum_b p_b (R_ib - delta_ib) w_b,   w = A^-1 u
//   the log-gamma terms differentiate through l with digamma
// (checked against the reference's autodiff through tests/golden/step_marginal_bge_reparam*.npz).
//
// Work decomposition: CTA = (particle, chunk of samples), one sample at a time; a warp takes a node: lane i holds row
// i of A in fp64 registers and the warp inverts it in place by Gauss-Jordan elimination (SPD: no pivoting; the pivot
// row is published through a shared-memory line) -- logdet A falls out as the sum of the log pivots, A^-1 serves
// every gradient term.  O(d^3) per (sample, node): this path is about supporting the combination, not about speed.
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"
#include "kernels_mc_lin_qr.cuh"   // entry_from_bits

namespace dibs {

__device__ __forceinline__ double shfl_xor_d(double v, int o) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o); hi = __shfl_xor_sync(0xffffffffu, hi, o);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}
// digamma for x > 0: recurrence up to x >= 6, then the asymptotic series
__device__ __forceinline__ double digamma_pos(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f * (1.0 / 132.0)))));
}

template <int DMAX, int MODE>
__global__ void __launch_bounds__(256) k_mc_bge_soft(const __grid_constant__ McParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(DMAX <= 32, "one lane per matrix row");
    const int d = p.d, dd = d * d;
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = p.st ? p.st->t : p.t_override;
    const int S = p.n_samples;
    const bool r_shared = p.bge_r_stride == 0;

    double* sR = reinterpret_cast<double*>(smem_raw);                         // [dd] (only when R is shared by all nodes)
    double* sLine = sR + (r_shared ? ((dd + 1) & ~1) : 0);                    // [8][4][DMAX]: pivot row, p, u, w per warp
    float* sS = reinterpret_cast<float*>(sLine + 8 * 4 * DMAX);               // [dd] alpha*scores (or exp(-alpha*scores))
    float* sG = sS + dd;                                                      // [dd] soft graph of the current sample
    float* sDS = sG + dd;                                                     // [dd] d lp / d G of the current sample
    float* sAcc = sDS + dd;                                                   // [dd] softmax-weighted running sum of dS
    float* sNode = sAcc + dd;                                                 // [DMAX] per-node log-probs

    const bool use_ext = p.g_ext != nullptr;
    const bool fast_soft = !use_ext && p.tau == 1.0f;
    const float alpha = p.alpha_linear * (float)t;
    for (int e = tid; e < dd; e += blockDim.x) {
        const float a = p.scores ? alpha * p.scores[(size_t)m * dd + e] : 0.0f;
        sS[e] = fast_soft ? expf(-a) : a;
        sAcc[e] = 0.0f;
        if (r_shared) sR[e] = p.bge_r[e];
    }
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);
    const uint32_t n_total = (uint32_t)S * dd;
    const double a_mu = (double)p.bge_alpha_mu, a_l = (double)p.bge_alpha_lambd;
    const double log_t = log((a_mu * (a_l - d - 1)) / (a_mu + 1));
    const float ta = p.tau * alpha;
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;
    double* myLine = sLine + (size_t)warp * 4 * DMAX;
    double* sPiv = myLine; double* sP = myLine + DMAX; double* sU = myLine + 2 * DMAX; double* sW = myLine + 3 * DMAX;
    __syncthreads();

    const int s_begin = c * p.s_per_chunk, s_end = min(S, s_begin + p.s_per_chunk);
    for (int s = s_begin; s < s_end; ++s) {
        // ---- the sample's soft graph (dibs.py:121-140, 430-431), zero diagonal
        for (int e = tid; e < dd; e += blockDim.x) {
            const int i = e / d, j = e - i * d;
            float g = 0.0f;
            if (i != j) {
                if (use_ext) g = p.g_ext[((size_t)m * S + s) * dd + e];
                else g = entry_from_bits<false>(jax_bits(key, (uint32_t)s * dd + e, n_total, p.partitionable), sS[e], fast_soft, p.tau);
            }
            sG[e] = g;
        }
        __syncthreads();
        // ---- one node per warp
        for (int j = warp; j < d; j += 8) {
            const bool row_ok = lane < d;
            const double pi = (row_ok && lane != j) ? (double)sG[lane * d + j] : 0.0;
            const double* Rj = r_shared ? sR : p.bge_r + (size_t)j * p.bge_r_stride;
            if (row_ok) sP[lane] = pi;
            __syncwarp();
            const double l = warp_sum_d(pi);
            // row `lane` of A = D R D + I - D^2
            double a[DMAX];
#pragma unroll
            for (int b = 0; b < DMAX; ++b) {
                double v = 0.0;
                if (row_ok && b < d) {
                    const double pb = sP[b];
                    v = pi * pb * Rj[lane * d + b];
                    if (b == lane) v += 1.0 - pi * pi;
                }
                a[b] = (b == lane && !(row_ok && b < d)) ? 1.0 : v;      // identity on the padding
            }
            // in-place Gauss-Jordan inversion; logdet A = sum of the log pivots
            double logdet = 0.0;
#pragma unroll
            for (int k = 0; k < DMAX; ++k) {
                if (k < d) {
                    if (lane == k) {
#pragma unroll
                        for (int b = 0; b < DMAX; ++b) sPiv[b] = a[b];
                    }
                    __syncwarp();
                    const double piv = sPiv[k];
                    const double inv = 1.0 / piv;
                    logdet += log(piv);
                    const double f = a[k];
                    if (lane == k) {
#pragma unroll
                        for (int b = 0; b < DMAX; ++b) a[b] = (b == k) ? inv : a[b] * inv;
                    } else {
#pragma unroll
                        for (int b = 0; b < DMAX; ++b) a[b] = (b == k) ? -f * inv : fma(-f * inv, sPiv[b], a[b]);
                    }
                    __syncwarp();
                }
            }
            // u_a = p_a R_aj,  w = A^-1 u,  s = R_jj - u^T w
            const double ui = row_ok ? pi * Rj[lane * d + j] : 0.0;
            if (row_ok) sU[lane] = ui;
            __syncwarp();
            double wi = 0.0;
#pragma unroll
            for (int b = 0; b < DMAX; ++b) if (b < d) wi = fma(a[b], sU[b], wi);
            if (row_ok) sW[lane] = wi;
            const double sch = Rj[j * d + j] - warp_sum_d(ui * wi);
            __syncwarp();
            // v_i = sum_b (A^-1)_ib p_b (R_ib - delta_ib),  t_i = sum_b p_b (R_ib - delta_ib) w_b
            double vi = 0.0, ti = 0.0;
#pragma unroll
            for (int b = 0; b < DMAX; ++b) {
                if (row_ok && b < d) {
                    const double rm = Rj[lane * d + b] - (b == lane ? 1.0 : 0.0);
                    const double pb = sP[b];
                    vi = fma(a[b] * rm, pb, vi);
                    ti = fma(pb * rm, sW[b], ti);
                }
            }
            const double cn = (double)p.bge_coef[2 * j];                  // N_j + alpha_lambd - d
            const bool valid = p.bge_coef[2 * j + 1] != 0.0f;             // N_j > 0   (linearGaussian.py:118)
            const double n_j = cn - (a_l - d);
            const double la = logdet, lb = la + log(sch);
            const double c1 = 0.5 * (cn + l), c2 = 0.5 * (cn + l + 1.0);
            const double lgam = 0.5 * (log(a_mu) - log(n_j + a_mu)) + lgamma(0.5 * (cn + l + 1.0)) - lgamma(0.5 * (a_l - d + l + 1.0)) -
                                0.5 * n_j * 1.1447298858494002 /* log(pi) */ + 0.5 * (a_l - d + 2.0 * l + 1.0) * log_t;
            const double dlg = 0.5 * digamma_pos(0.5 * (cn + l + 1.0)) - 0.5 * digamma_pos(0.5 * (a_l - d + l + 1.0)) + log_t;
            const double dla = 2.0 * vi;
            const double dsch = -2.0 * Rj[lane * d + j] * wi + 2.0 * wi * ti;
            const double dlb = dla + dsch / sch;
            const double col = dlg + 0.5 * la - 0.5 * lb + c1 * dla - c2 * dlb;
            if (row_ok) sDS[lane * d + j] = (valid && lane != j) ? (float)col : 0.0f;
            if (lane == 0) sNode[j] = valid ? (float)(lgam + c1 * la - c2 * lb) : 0.0f;
            __syncwarp();
        }
        __syncthreads();
        float lp = 0.0f;
        for (int j = 0; j < d; ++j) lp += sNode[j];                  // every thread, same order
        if (p.lp_out && tid == 0) p.lp_out[(size_t)m * S + s] = lp;
        if (MODE != MC_LP_ONLY) {
            // online softmax over the samples (dibs.py:451-457): running max, sum of exponentials, weighted dS
            const float m_new = fmaxf(m_run, lp);
            const float scale = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
            const float ex = expf(lp - m_new);
            l_run = l_run * scale + ex;
            sum_lp += lp;
            m_run = m_new;
            for (int e = tid; e < dd; e += blockDim.x) {
                const float g = sG[e];
                const float val = sDS[e] * ta * g * (1.0f - g);      // dS = d lp/dG * tau*alpha*g(1-g)  (App. B-4)
                sAcc[e] = fmaf(sAcc[e], scale, ex * val);
            }
        }
        __syncthreads();
    }
    if (MODE == MC_LP_ONLY) return;
    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    for (int e = tid; e < dd; e += blockDim.x) out[e] = sAcc[e];
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, reinterpret_cast<float*>(smem_raw));
}

inline size_t mc_bge_soft_smem(int d, int dmax, bool r_shared) {
    const size_t dd = (size_t)d * d;
    size_t bytes = (r_shared ? ((dd + 1) & ~(size_t)1) : 0) * sizeof(double);
    bytes += (size_t)8 * 4 * dmax * sizeof(double);
    bytes += (4 * dd + dmax + 4) * sizeof(float);
    return bytes + 16;
}

}  // namespace dibs
