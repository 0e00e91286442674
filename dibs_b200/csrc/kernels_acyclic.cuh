// Acyclicity-constraint gradient, n_vars <= 32: one warp per Gumbel-soft graph, one lane per matrix row.
//
// replaces: dibs/inference/dibs.py:557-601 (constraint_gumbel, grad_constraint_gumbel: mean over A logistic-noise
// samples of grad_Z h(soft_G(Z, eps))) and dibs/graph_utils.py:8-28 (acyclic_constr_nograd,
// h(G) = tr((I + G/d)^d) - d via jnp.linalg.matrix_power).  Closed form:
//   dh/dG = ((I + G/d)^(d-1))^T,   dS = dh/dG o tau alpha G (1 - G)  off the diagonal.
//
// Decomposition: CTA = particle; a warp takes sample PAIRS (a, a + A/2) -- in JAX's legacy threefry layout the
// two lanes of one block -- draws both soft graphs into shared memory with all 32 lanes, then runs the binary
// exponentiation (LSB first, like jnp.linalg.matrix_power) with lane i holding row i of the running square
// and of the running result in registers.  A product row_i(X) * Z streams Z's rows from shared memory as
// warp-uniform 128-bit broadcasts: d^2/4 LDS.128 for d^2 FMAs per lane, DMAX independent accumulators.
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

// 4 CTAs of 128 threads per SM -> 128 registers per thread (96 / 80 registers measured slower: spills in the products)
template <int DMAX>
__global__ void __launch_bounds__(128, 4) k_acyclic_rows(const __grid_constant__ AcycParams p) {
    extern __shared__ __align__(16) float smem[];
    static_assert(DMAX % 4 == 0 && DMAX <= 32, "one lane per row, 128-bit row loads");
    const int d = p.d, k = p.k, dd = d * d;
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int t = p.st ? p.st->t : p.t_override;
    const float alpha = p.alpha_linear * (float)t;
    constexpr int MAT = DMAX * DMAX;

    float* sS = smem;                                   // [d*d] alpha*scores, or exp(-alpha*scores) when tau == 1
    float* sZ = sS + ((dd + 3) & ~3);                   // [2*d*k] staging, later the cross-warp reduction [n_warps][d*d]
    const int stage = max(2 * d * k, n_warps * dd);
    float* wbase = sZ + ((stage + 3) & ~3) + (size_t)warp * 4 * MAT;
    float* sM = wbase;                                  // [2][DMAX][DMAX]  I + G/d of the pair (running square of the active one)
    float* sF = wbase + 2 * MAT;                        // [2][DMAX][DMAX]  tau alpha G (1 - G)

    const bool fast_soft = p.tau == 1.0f;
    for (int e = tid; e < dd; e += blockDim.x) {
        const float a = alpha * p.scores[(size_t)m * dd + e];
        sS[e] = fast_soft ? expf(-a) : a;
    }
    const uint2 key = make_uint2(p.keys_override[2 * m], p.keys_override[2 * m + 1]);
    // zero the padding of this warp's matrices once (rows / columns >= d are never written again)
    for (int e = lane; e < 4 * MAT; e += 32) wbase[e] = 0.0f;
    __syncthreads();

    const float inv_d = 1.0f / (float)d;
    const float ta = p.tau * alpha;
    const int half_a = p.n_samples >> 1;
    const uint32_t half = ((uint32_t)p.n_samples * dd) >> 1;
    const int row = lane < d ? lane : d - 1;           // idle lanes shadow the last row and never write

    float accT[DMAX];                                   // accT[j] = sum over this warp's samples of dS[j][lane]
#pragma unroll
    for (int j = 0; j < DMAX; ++j) accT[j] = 0.0f;

    // blockIdx.y = chunk of n_warps consecutive sample pairs (a fixed decomposition of the A axis: the summation
    // order never depends on how many particles a rank owns)
    const int a_begin = blockIdx.y * n_warps;
    const int a_end = min(half_a, a_begin + n_warps);
    for (int a = a_begin + warp; a < a_end; a += n_warps) {
        // ---- draw the two soft graphs (samples a and a + A/2) with all lanes
        constexpr int IL = 4;                 // independent threefry chains per lane
        for (int eb = lane; eb < dd; eb += IL * 32) {
            uint32_t x0[IL], x1[IL];
#pragma unroll
            for (int u = 0; u < IL; ++u) {
                const uint32_t e0 = (uint32_t)a * dd + (eb + u * 32);
                x0[u] = e0; x1[u] = e0 + half;
            }
            threefry2x32_n<IL>(key.x, key.y, x0, x1);
#pragma unroll
            for (int u = 0; u < IL; ++u) {
                const int e = eb + u * 32;
                if (e >= dd) continue;
                const int i = __float2int_rz(((float)e + 0.5f) * inv_d), j = e - i * d;    // exact for e < 2^13
                float g0 = 0.0f, g1 = 0.0f;
                if (i != j) {
                    const float sa = sS[e];
                    g0 = entry_from_bits<false>(x0[u], sa, fast_soft, p.tau);
                    g1 = entry_from_bits<false>(x1[u], sa, fast_soft, p.tau);
                }
                const float eye = (i == j) ? 1.0f : 0.0f;
                sM[i * DMAX + j] = eye + inv_d * g0;                     // graph_utils.py:22-25
                sM[MAT + i * DMAX + j] = eye + inv_d * g1;
                sF[i * DMAX + j] = ta * g0 * (1.0f - g0);
                sF[MAT + i * DMAX + j] = ta * g1 * (1.0f - g1);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            float* Zs = sM + s * MAT;
            const float* Fs = sF + s * MAT;
            float zr[DMAX], res[DMAX], out[DMAX];
            {
                const float4* src = reinterpret_cast<const float4*>(Zs + row * DMAX);
#pragma unroll
                for (int q = 0; q < DMAX / 4; ++q) {
                    const float4 v = src[q];
                    zr[4 * q] = v.x; zr[4 * q + 1] = v.y; zr[4 * q + 2] = v.z; zr[4 * q + 3] = v.w;
                }
            }
            // E = M^(d-1): binary exponentiation, least-significant bit first
            bool have_res = false;
            int n = d - 1;
#pragma unroll 1
            while (n > 0) {
                if (n & 1) {
                    if (!have_res) {
#pragma unroll
                        for (int j = 0; j < DMAX; ++j) res[j] = zr[j];
                        have_res = true;
                    } else {
                        row_times_smem<DMAX>(res, Zs, d, out);
#pragma unroll
                        for (int j = 0; j < DMAX; ++j) res[j] = out[j];
                    }
                }
                n >>= 1;
                if (n > 0) {
                    row_times_smem<DMAX>(zr, Zs, d, out);
                    __syncwarp();                                  // everyone is done reading the old square
                    if (lane < d) {
                        float4* dst = reinterpret_cast<float4*>(Zs + lane * DMAX);
#pragma unroll
                        for (int q = 0; q < DMAX / 4; ++q)
                            dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
                    }
#pragma unroll
                    for (int j = 0; j < DMAX; ++j) zr[j] = out[j];
                    __syncwarp();
                }
            }
            // dS[j][i] = E[i][j] * tau alpha g_ji (1 - g_ji): lane i owns column i of dS
#pragma unroll
            for (int j = 0; j < DMAX; ++j)
                if (j < d) accT[j] = fmaf(res[j], Fs[j * DMAX + row], accT[j]);
        }
        __syncwarp();
    }
    // ---- deterministic cross-warp reduction: sRed[warp][j*d + i]
    __syncthreads();
    float* sRed = sZ;
    if (lane < d) {
#pragma unroll
        for (int j = 0; j < DMAX; ++j)
            if (j < d) sRed[(size_t)warp * dd + j * d + lane] = accT[j];
    }
    __syncthreads();
    float* outp = p.ds_out + ((size_t)m * gridDim.y + blockIdx.y) * dd;
    for (int e = tid; e < dd; e += blockDim.x) {
        float sum = 0.0f;
        for (int w = 0; w < n_warps; ++w) sum += sRed[(size_t)w * dd + e];
        outp[e] = sum;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t acyclic_rows_smem(int d, int k, int dmax, int n_warps) {
    size_t stage = (size_t)2 * d * k;
    if ((size_t)n_warps * d * d > stage) stage = (size_t)n_warps * d * d;
    return ((((size_t)d * d + 3) & ~(size_t)3) + ((stage + 3) & ~(size_t)3) + (size_t)n_warps * 4 * dmax * dmax + 4) * sizeof(float);
}

}  // namespace dibs
