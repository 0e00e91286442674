// Acyclicity-constraint gradient, n_vars <= 32: one warp per Gumbel-soft graph, one lane per matrix row.
//
// replaces: dibs/inference/dibs.py:557-601 (constraint_gumbel, grad_constraint_gumbel: mean over A logistic-noise
// samples of grad_Z h(soft_G(Z, eps))) and dibs/graph_utils.py:8-28 (acyclic_constr_nograd,
// h(G) = tr((I + G/d)^d) - d via jnp.linalg.matrix_power).  Closed form:
//   dh/dG = ((I + G/d)^(d-1))^T,   dS = dh/dG o tau alpha G (1 - G)  off the diagonal.
//
// Decomposition: CTA = (particle, 4 sample PAIRS (a, a + A/2) -- in JAX's legacy threefry layout the two lanes of
// one block); all threads draw the 8 soft graphs into shared memory, then every thread runs the binary
// exponentiation (LSB first, like jnp.linalg.matrix_power) for ONE row of ONE sample, holding the row of the running
// square and of the running result in registers.  A product row_i(X) * Z streams Z's rows from shared memory as
// 128-bit broadcasts: d^2/4 LDS.128 for d^2 FMAs per thread, DMAX independent accumulators.
#pragma once
#include "common.cuh"
#include "kernels_prior.cuh"
#include "kernels_mc_lin_qr.cuh"   // logistic_u_from_bits

namespace dibs {

// out[j] = sum_k x[k] * Zs[k][j]   (Zs row-major, leading dimension DMAX, zero-padded); packed FFMA2: the row of Z
// arrives as 128-bit broadcasts = two register pairs, x[k] is duplicated into a pair once per k
template <int DMAX>
__device__ __forceinline__ void row_times_smem(const float (&x)[DMAX], const float* __restrict__ Zs, int d,
                                               float (&out)[DMAX]) {
    f32x2 o2[DMAX / 2];
#pragma unroll
    for (int j = 0; j < DMAX / 2; ++j) o2[j] = 0ull;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        if (k < d) {
            const f32x2 xk = pack2(x[k], x[k]);
            const ulonglong2* zr = reinterpret_cast<const ulonglong2*>(Zs + k * DMAX);
#pragma unroll
            for (int q = 0; q < DMAX / 4; ++q) {
                const ulonglong2 z = zr[q];
                o2[2 * q] = fma2(xk, z.x, o2[2 * q]);
                o2[2 * q + 1] = fma2(xk, z.y, o2[2 * q + 1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < DMAX / 2; ++j) { out[2 * j] = lo2(o2[j]); out[2 * j + 1] = hi2(o2[j]); }
}

// Thread = (sample slot, matrix row): ACYC_SLOTS = 8 soft graphs (4 sample pairs) per CTA with their rows packed
// densely over the CTA's lanes (thread t -> slot t / d, row t % d), so at n_vars = 20 all 160 lanes of the 5 warps
// carry a row -- one warp per sample left 12 of 32 lanes idle in every product (37 % of the FMA issue slots).  A
// warp then straddles two samples: its Z-row reads are two broadcasts per instruction, which land in different bank
// groups (slot stride MAT = 400 floats = 16 banks), and the row write-back of a squaring is fenced by CTA barriers.
constexpr int ACYC_SLOTS = 8;

template <int DMAX>
__global__ void __launch_bounds__(8 * DMAX, (DMAX <= 20 ? 5 : 2)) k_acyclic_rows(const __grid_constant__ AcycParams p) {
    extern __shared__ __align__(16) float smem[];
    static_assert(DMAX % 4 == 0 && DMAX <= 32, "128-bit row loads");
    const int d = p.d, dd = d * d;
    const int m = blockIdx.x, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const float alpha = p.alpha_linear * (float)t;
    constexpr int MAT = DMAX * DMAX;

    float* sS = smem;                                   // [d*d] alpha*scores, or exp(-alpha*scores) when tau == 1
    float* sM = sS + ((dd + 3) & ~3);                   // [SLOTS][DMAX][DMAX]  I + G/d of each sample (then its running square)
    float* sF = sM + ACYC_SLOTS * MAT;                  // [SLOTS][DMAX][DMAX]  tau alpha G (1 - G)
    float* sR = sF + ACYC_SLOTS * MAT + tid * DMAX;     // [threads][DMAX]      this thread's row of the running result

    const bool fast_soft = p.tau == 1.0f;
    for (int e = tid; e < dd; e += blockDim.x) {
        const float a = alpha * p.scores[(size_t)m * dd + e];
        sS[e] = fast_soft ? expf(-a) : a;
    }
    const uint2 key = make_uint2(p.keys_override[2 * m], p.keys_override[2 * m + 1]);
    // zero the padding once (rows / columns >= d are never written again); none at n_vars == DMAX
    if (d < DMAX)
        for (int e = tid; e < 2 * ACYC_SLOTS * MAT / 4; e += blockDim.x)
            reinterpret_cast<float4*>(sM)[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncthreads();

    const float inv_d = 1.0f / (float)d, inv_dd = 1.0f / (float)dd;
    const float ta = p.tau * alpha;
    const int half_a = p.n_samples >> 1;
    const uint32_t half = ((uint32_t)p.n_samples * dd) >> 1;
    // blockIdx.y = chunk of ACYC_SLOTS / 2 consecutive sample pairs (a fixed decomposition of the A axis: the summation
    // order never depends on how many particles a rank owns)
    const int a_begin = blockIdx.y * (ACYC_SLOTS / 2);
    const int n_pairs = min(half_a, a_begin + ACYC_SLOTS / 2) - a_begin;

    // ---- draw the soft graphs of the chunk with all threads: pair q -> slots 2q (sample a) and 2q + 1 (sample a + A/2),
    // the two lanes of one threefry block in JAX's legacy layout
    constexpr int IL = 4;                 // independent threefry chains per thread
    for (int idx0 = tid; idx0 < n_pairs * dd; idx0 += IL * blockDim.x) {
        uint32_t x0[IL], x1[IL];
        int pr[IL], ee[IL];
#pragma unroll
        for (int u = 0; u < IL; ++u) {
            const int idx = idx0 + u * blockDim.x;
            pr[u] = __float2int_rz(((float)idx + 0.5f) * inv_dd);           // exact for idx < 2^13
            ee[u] = idx - pr[u] * dd;
            const uint32_t e0 = (uint32_t)(a_begin + pr[u]) * dd + ee[u];
            x0[u] = e0; x1[u] = e0 + half;
        }
        threefry2x32_n<IL>(key.x, key.y, x0, x1);
#pragma unroll
        for (int u = 0; u < IL; ++u) {
            if (idx0 + u * blockDim.x >= n_pairs * dd) continue;
            const int e = ee[u];
            const int i = __float2int_rz(((float)e + 0.5f) * inv_d), j = e - i * d;
            float g0 = 0.0f, g1 = 0.0f;
            if (i != j) {
                const float sa = sS[e];
                g0 = entry_from_bits<false>(x0[u], sa, fast_soft, p.tau);
                g1 = entry_from_bits<false>(x1[u], sa, fast_soft, p.tau);
            }
            const float eye = (i == j) ? 1.0f : 0.0f;
            float* m0 = sM + (2 * pr[u]) * MAT + i * DMAX + j;
            float* f0 = sF + (2 * pr[u]) * MAT + i * DMAX + j;
            m0[0] = eye + inv_d * g0;                              // graph_utils.py:22-25
            m0[MAT] = eye + inv_d * g1;
            f0[0] = ta * g0 * (1.0f - g0);
            f0[MAT] = ta * g1 * (1.0f - g1);
        }
    }
    __syncthreads();

    const int slot_raw = tid / d;
    const bool valid = slot_raw < 2 * n_pairs;          // this thread carries a row of a live sample
    const int slot = valid ? slot_raw : 0;
    const int row = valid ? tid - slot_raw * d : 0;     // idle threads shadow (slot 0, row 0) and never write
    float* Zs = sM + slot * MAT;
    const float* Fs = sF + slot * MAT;
    // the result row is parked in shared memory between its (few) products: 20 registers less, which is what lets a
    // fourth CTA fit on the SM at n_vars = 20 (strip stride DMAX floats: conflict-free 128-bit accesses)
    float zr[DMAX], out[DMAX];
    float4* resv = reinterpret_cast<float4*>(sR);
    {
        const float4* src = reinterpret_cast<const float4*>(Zs + row * DMAX);
#pragma unroll
        for (int q = 0; q < DMAX / 4; ++q) {
            const float4 v = src[q];
            zr[4 * q] = v.x; zr[4 * q + 1] = v.y; zr[4 * q + 2] = v.z; zr[4 * q + 3] = v.w;
        }
    }
    // E = M^(d-1): binary exponentiation, least-significant bit first (jnp.linalg.matrix_power)
    bool have_res = false;
    int n = d - 1;
#pragma unroll 1
    while (n > 0) {
        if (n & 1) {
            if (!have_res) {
#pragma unroll
                for (int q = 0; q < DMAX / 4; ++q) resv[q] = make_float4(zr[4 * q], zr[4 * q + 1], zr[4 * q + 2], zr[4 * q + 3]);
                have_res = true;
            } else {
                float res[DMAX];
#pragma unroll
                for (int q = 0; q < DMAX / 4; ++q) {
                    const float4 v = resv[q];
                    res[4 * q] = v.x; res[4 * q + 1] = v.y; res[4 * q + 2] = v.z; res[4 * q + 3] = v.w;
                }
                row_times_smem<DMAX>(res, Zs, d, out);
#pragma unroll
                for (int q = 0; q < DMAX / 4; ++q) resv[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            }
        }
        n >>= 1;
        if (n > 0) {
            row_times_smem<DMAX>(zr, Zs, d, out);
            __syncthreads();                               // everyone is done reading the old squares
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(Zs + row * DMAX);
#pragma unroll
                for (int q = 0; q < DMAX / 4; ++q)
                    dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            }
#pragma unroll
            for (int j = 0; j < DMAX; ++j) zr[j] = out[j];
            __syncthreads();
        }
    }
    // dS[j][i] = E[i][j] * tau alpha g_ji (1 - g_ji): thread (slot, i) owns column i of its sample's dS
    __syncthreads();                                       // the squares are dead: sM becomes the reduction buffer
    float* sRed = sM;                                      // [SLOTS][d*d]
    if (valid) {
#pragma unroll
        for (int q = 0; q < DMAX / 4; ++q) {
            const float4 v = resv[q];
            out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < DMAX; ++j)
            if (j < d) sRed[(size_t)slot * dd + j * d + row] = (d == 1) ? Fs[j * DMAX + row] : out[j] * Fs[j * DMAX + row];
    }
    __syncthreads();
    // deterministic reduction over the chunk's samples, in slot order
    float* outp = p.ds_out + ((size_t)m * gridDim.y + blockIdx.y) * dd;
    for (int e = tid; e < dd; e += blockDim.x) {
        float sum = 0.0f;
        for (int w = 0; w < 2 * n_pairs; ++w) sum += sRed[(size_t)w * dd + e];
        outp[e] = sum;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t acyclic_rows_smem(int d, int dmax) {
    return ((((size_t)d * d + 3) & ~(size_t)3) + (size_t)2 * ACYC_SLOTS * dmax * dmax + (size_t)ACYC_SLOTS * dmax * dmax + 4) * sizeof(float);
}
inline int acyclic_rows_threads(int d) { return ((ACYC_SLOTS * d + 31) / 32) * 32; }

}  // namespace dibs
