// n_vars > 32: the per-graph d x d x d contractions as CTA-cooperative register-tiled matrix products on
// shared-memory operands (fp32 SIMT, packed FFMA2).  Two users:
//
//   k_acyclic_dense  -- acyclicity-constraint gradient: E = (I + G/d)^(d-1) by binary exponentiation
//                       (dibs/inference/dibs.py:557-601, dibs/graph_utils.py:8-28; closed form SURVEY App. B-3/4)
//   k_mc_lin_dense   -- LinearGaussian Monte-Carlo pass in QR-factor form for ALL nodes of a graph at once:
//                       Y = Rx (I - G o Theta), ssq_j = |Y_:j|^2, B = Rx^T Y   (dibs/models/linearGaussian.py:278-338
//                       under dibs/inference/dibs.py:325-459,488-551; see kernels_mc_lin_qr.cuh for the algebra)
//
// Tile scheme: matrices are row-major with leading dimension LD (multiple of 8, zero padded), HALF = LD / 2.
// Thread (ty, tx), ty, tx < HALF / 4, owns the 8 x 8 outputs  rows {4ty..4ty+3} U {HALF+4ty..}  x  columns
// {4tx..4tx+3} U {HALF+4tx..}: the split halves make both operand fetches contiguous across a warp (conflict-free
// 128-bit loads; same-ty / same-tx lanes broadcast).  C += A B takes the A operand TRANSPOSED (At[k][i]) so that
// both fragments of a k step are row reads: 4 LDS.128 for 32 FFMA2.
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"
#include "kernels_prior.cuh"
#include "kernels_mc_lin_qr.cuh"   // entry_from_bits

namespace dibs {

struct Tile8 { f32x2 c[8][4]; };   // c[a][q]: row a (0..3 lower half, 4..7 upper half), column pair q (0,1 lower; 2,3 upper)

__device__ __forceinline__ void tile_zero(Tile8& t) {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) t.c[a][q] = 0ull;
}

// t += sum_{k in [k0, k1)} At[k][rows] * B[k][cols];  LO / HI: which row halves take part
template <bool LO, bool HI>
__device__ __forceinline__ void tile_mm(const float* __restrict__ At, const float* __restrict__ B, int LD, int HALF,
                                        int ty, int tx, int k0, int k1, Tile8& t) {
    const float* ap = At + (size_t)k0 * LD + 4 * ty;
    const float* bp = B + (size_t)k0 * LD + 4 * tx;
#pragma unroll 2
    for (int k = k0; k < k1; ++k, ap += LD, bp += LD) {
        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(bp);
        const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(bp + HALF);
        if (LO) {
            const float4 a = *reinterpret_cast<const float4*>(ap);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const f32x2 aa = pack2(av[r], av[r]);
                t.c[r][0] = fma2(aa, b0.x, t.c[r][0]); t.c[r][1] = fma2(aa, b0.y, t.c[r][1]);
                t.c[r][2] = fma2(aa, b1.x, t.c[r][2]); t.c[r][3] = fma2(aa, b1.y, t.c[r][3]);
            }
        }
        if (HI) {
            const float4 a = *reinterpret_cast<const float4*>(ap + HALF);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const f32x2 aa = pack2(av[r], av[r]);
                t.c[4 + r][0] = fma2(aa, b0.x, t.c[4 + r][0]); t.c[4 + r][1] = fma2(aa, b0.y, t.c[4 + r][1]);
                t.c[4 + r][2] = fma2(aa, b1.x, t.c[4 + r][2]); t.c[4 + r][3] = fma2(aa, b1.y, t.c[4 + r][3]);
            }
        }
    }
}

__device__ __forceinline__ int tile_row(int ty, int HALF, int a) { return (a < 4 ? 0 : HALF) + 4 * ty + (a & 3); }
__device__ __forceinline__ int tile_col(int tx, int HALF, int b) { return (b < 4 ? 0 : HALF) + 4 * tx + (b & 3); }
__device__ __forceinline__ float tile_get(const Tile8& t, int a, int b) { return (b & 1) ? hi2(t.c[a][b >> 1]) : lo2(t.c[a][b >> 1]); }

// M[row][col] = tile (row-major store, 128-bit)
__device__ __forceinline__ void tile_store(const Tile8& t, float* M, int LD, int HALF, int ty, int tx) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        float* r = M + (size_t)tile_row(ty, HALF, a) * LD + 4 * tx;
        *reinterpret_cast<ulonglong2*>(r) = make_ulonglong2(t.c[a][0], t.c[a][1]);
        *reinterpret_cast<ulonglong2*>(r + HALF) = make_ulonglong2(t.c[a][2], t.c[a][3]);
    }
}
// Mt[col][row] = tile (transposed store)
__device__ __forceinline__ void tile_store_t(const Tile8& t, float* Mt, int LD, int HALF, int ty, int tx) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        float* r = Mt + (size_t)tile_col(tx, HALF, b) * LD + 4 * ty;
        *reinterpret_cast<float4*>(r) = make_float4(tile_get(t, 0, b), tile_get(t, 1, b), tile_get(t, 2, b), tile_get(t, 3, b));
        *reinterpret_cast<float4*>(r + HALF) = make_float4(tile_get(t, 4, b), tile_get(t, 5, b), tile_get(t, 6, b), tile_get(t, 7, b));
    }
}

static inline int dense_ld(int d) { return (d + 7) & ~7; }
static inline int dense_nt(int d) { const int q = dense_ld(d) / 8; return q * q; }       // active threads per group

// ------------------------------------------------------------------------------------------
// acyclicity gradient, any n_vars: CTA = (particle, chunk of samples); NG thread groups, one sample each per round
// ------------------------------------------------------------------------------------------
struct AcycDenseShape { int ld, nt, ng, threads, rounds, chunks; size_t smem; };

static inline AcycDenseShape acyc_dense_shape(int d, int n_samples) {
    AcycDenseShape s;
    s.ld = dense_ld(d); s.nt = dense_nt(d);
    const size_t mat = (size_t)s.ld * s.ld * sizeof(float);
    int ng = 256 / s.nt; if (ng < 1) ng = 1;
    while (ng > 1 && (4 * mat * ng + (size_t)d * d * sizeof(float) + 64 > 200 * 1024)) --ng;
    if (ng > n_samples) ng = n_samples;
    s.ng = ng;
    s.threads = ((ng * s.nt + 31) / 32) * 32;
    // a fixed decomposition of the sample axis (never a function of the particle count): <= 8 chunks per particle
    s.rounds = (n_samples + ng * 8 - 1) / (ng * 8);
    s.chunks = (n_samples + ng * s.rounds - 1) / (ng * s.rounds);
    s.smem = 4 * mat * ng + (size_t)d * d * sizeof(float) + 64;
    return s;
}

__global__ void __launch_bounds__(256, 1) k_acyclic_dense(const __grid_constant__ AcycParams p, int LD, int NT, int NG, int rounds) {
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, dd = d * d, HALF = LD / 2, TQ = HALF / 4;
    const int m = blockIdx.x, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const float alpha = p.alpha_linear * (float)t;
    const int MAT = LD * LD;
    const int grp = tid / NT, gt = tid - grp * NT;
    const bool active = grp < NG;
    const int ty = gt / TQ, tx = gt - ty * TQ;

    float* sS = smem;                                   // [dd] alpha*scores, or exp(-alpha*scores) when tau == 1
    float* gbase = smem + ((dd + 3) & ~3) + (size_t)(active ? grp : 0) * 4 * MAT;
    float* sG = gbase;                                  // soft graph
    float* sZ = gbase + MAT;                            // running square (row-major)
    float* sZt = gbase + 2 * MAT;                       // its transpose
    float* sRt = gbase + 3 * MAT;                       // running result, transposed

    const bool fast_soft = p.tau == 1.0f;
    for (int e = tid; e < dd; e += blockDim.x) {
        const float a = alpha * p.scores[(size_t)m * dd + e];
        sS[e] = fast_soft ? expf(-a) : a;
    }
    if (active) for (int e = gt; e < 4 * MAT; e += NT) gbase[e] = 0.0f;       // padding stays zero
    const uint2 key = make_uint2(p.keys_override[2 * m], p.keys_override[2 * m + 1]);
    __syncthreads();

    const float inv_d = 1.0f / (float)d;
    const float ta = p.tau * alpha;
    const uint32_t n_total = (uint32_t)p.n_samples * dd;
    Tile8 acc;                                          // sum over this group's samples of dS on the thread's tile
    tile_zero(acc);

    for (int r = 0; r < rounds; ++r) {
        const int a = (blockIdx.y * rounds + r) * NG + grp;
        const bool valid = active && a < p.n_samples;   // group-uniform; barriers below are CTA-wide
        if (valid) {
            for (int e = gt; e < dd; e += NT) {
                const int i = e / d, j = e - i * d;
                float g = 0.0f;
                if (i != j) {
                    const uint32_t bits = jax_bits(key, (uint32_t)a * dd + e, n_total, p.partitionable);
                    g = entry_from_bits<false>(bits, sS[e], fast_soft, p.tau);
                }
                const float mz = (i == j ? 1.0f : 0.0f) + inv_d * g;            // graph_utils.py:22-25
                sG[i * LD + j] = g; sZ[i * LD + j] = mz; sZt[j * LD + i] = mz;
            }
        }
        __syncthreads();
        // E = M^(d-1): binary exponentiation, least-significant bit first (jnp.linalg.matrix_power)
        bool have_res = false;
        int n = d - 1;
        while (n > 0) {
            if (n & 1) {
                if (!have_res) {
                    if (valid) for (int e = gt; e < MAT; e += NT) sRt[e] = sZt[e];
                    have_res = true;
                    __syncthreads();
                } else {
                    Tile8 c; tile_zero(c);
                    if (valid) tile_mm<true, true>(sRt, sZ, LD, HALF, ty, tx, 0, d, c);      // res * Z
                    __syncthreads();
                    if (valid) tile_store_t(c, sRt, LD, HALF, ty, tx);
                    __syncthreads();
                }
            }
            n >>= 1;
            if (n > 0) {
                Tile8 c; tile_zero(c);
                if (valid) tile_mm<true, true>(sZt, sZ, LD, HALF, ty, tx, 0, d, c);          // Z * Z
                __syncthreads();
                if (valid) { tile_store(c, sZ, LD, HALF, ty, tx); tile_store_t(c, sZt, LD, HALF, ty, tx); }
                __syncthreads();
            }
        }
        // dS[a][b] = E[b][a] * tau alpha g_ab (1 - g_ab) = Rt[a][b] * F[a][b]   (d = 1: E = I)
        if (valid) {
#pragma unroll
            for (int a8 = 0; a8 < 8; ++a8) {
                const int row = tile_row(ty, HALF, a8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = (q < 2 ? 0 : HALF) + 4 * tx + 2 * (q & 1);
                    const float2 g = *reinterpret_cast<const float2*>(&sG[row * LD + col]);
                    const float2 e = have_res ? *reinterpret_cast<const float2*>(&sRt[row * LD + col])
                                              : make_float2(row == col ? 1.0f : 0.0f, row == col + 1 ? 1.0f : 0.0f);
                    const f32x2 f = pack2(ta * g.x * (1.0f - g.x), ta * g.y * (1.0f - g.y));
                    acc.c[a8][q] = fma2(pack2(e.x, e.y), f, acc.c[a8][q]);
                }
            }
        }
        __syncthreads();
    }
    // cross-group reduction in fixed order through shared memory (group matrices are free now)
    float* sRed = smem + ((dd + 3) & ~3);               // [NG][MAT]
    if (active) tile_store(acc, sRed + (size_t)grp * MAT, LD, HALF, ty, tx);
    __syncthreads();
    float* outp = p.ds_out + ((size_t)m * gridDim.y + blockIdx.y) * dd;
    for (int e = tid; e < dd; e += blockDim.x) {
        const int i = e / d, j = e - i * d;
        float sum = 0.0f;
        for (int g = 0; g < NG; ++g) sum += sRed[(size_t)g * MAT + i * LD + j];
        outp[e] = sum;
    }
    fuse_arrive(p.fuse, m, smem);
}

// ------------------------------------------------------------------------------------------
// same pass with 4 x 4 register tiles (rows 4ty.., columns 4tx.., LD multiple of 4): four times the threads per
// matrix and a quarter of the shared memory per thread -- the better trade for 32 < n_vars <= 64, where an 8 x 8
// tiling leaves a CTA with < 64 threads per sample
// ------------------------------------------------------------------------------------------
struct Tile4 { f32x2 c[4][2]; };

__device__ __forceinline__ void tile4_mm(const float* __restrict__ At, const float* __restrict__ B, int LD, int ty, int tx,
                                         int k0, int k1, Tile4& t) {
    const float* ap = At + (size_t)k0 * LD + 4 * ty;
    const float* bp = B + (size_t)k0 * LD + 4 * tx;
#pragma unroll 4
    for (int k = k0; k < k1; ++k, ap += LD, bp += LD) {
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(bp);
        const float4 a = *reinterpret_cast<const float4*>(ap);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const f32x2 aa = pack2(av[r], av[r]);
            t.c[r][0] = fma2(aa, b.x, t.c[r][0]);
            t.c[r][1] = fma2(aa, b.y, t.c[r][1]);
        }
    }
}
__device__ __forceinline__ float tile4_get(const Tile4& t, int a, int b) { return (b & 1) ? hi2(t.c[a][b >> 1]) : lo2(t.c[a][b >> 1]); }
__device__ __forceinline__ void tile4_store(const Tile4& t, float* M, int LD, int ty, int tx) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
        *reinterpret_cast<ulonglong2*>(M + (size_t)(4 * ty + a) * LD + 4 * tx) = make_ulonglong2(t.c[a][0], t.c[a][1]);
}
__device__ __forceinline__ void tile4_store_t(const Tile4& t, float* Mt, int LD, int ty, int tx) {
#pragma unroll
    for (int b = 0; b < 4; ++b)
        *reinterpret_cast<float4*>(Mt + (size_t)(4 * tx + b) * LD + 4 * ty) =
            make_float4(tile4_get(t, 0, b), tile4_get(t, 1, b), tile4_get(t, 2, b), tile4_get(t, 3, b));
}

static inline AcycDenseShape acyc_dense4_shape(int d, int n_samples) {
    AcycDenseShape s;
    s.ld = (d + 3) & ~3;
    const int q = s.ld / 4;
    s.nt = q * q;
    const size_t mat = (size_t)s.ld * s.ld * sizeof(float);
    s.ng = 1;
    s.threads = ((s.nt + 31) / 32) * 32;
    s.rounds = (n_samples + 7) / 8;                      // <= 8 chunks per particle, fixed decomposition
    s.chunks = (n_samples + s.rounds - 1) / s.rounds;
    s.smem = 4 * mat + (size_t)d * d * sizeof(float) + 64;
    return s;
}

__global__ void __launch_bounds__(256, 2) k_acyclic_dense4(const __grid_constant__ AcycParams p, int LD, int NT, int rounds) {
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, dd = d * d, TQ = LD / 4;
    const int m = blockIdx.x, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const float alpha = p.alpha_linear * (float)t;
    const int MAT = LD * LD;
    const bool active = tid < NT;
    const int ty = active ? tid / TQ : 0, tx = active ? tid - ty * TQ : 0;

    float* sS = smem;                                   // [dd] alpha*scores, or exp(-alpha*scores) when tau == 1
    float* sG = smem + ((dd + 3) & ~3);                 // soft graph
    float* sZ = sG + MAT;                               // running square (row-major)
    float* sZt = sZ + MAT;                              // its transpose
    float* sRt = sZt + MAT;                             // running result, transposed

    const bool fast_soft = p.tau == 1.0f;
    for (int e = tid; e < dd; e += blockDim.x) {
        const float a = alpha * p.scores[(size_t)m * dd + e];
        sS[e] = fast_soft ? expf(-a) : a;
    }
    for (int e = tid; e < 4 * MAT; e += blockDim.x) sG[e] = 0.0f;          // padding stays zero
    const uint2 key = make_uint2(p.keys_override[2 * m], p.keys_override[2 * m + 1]);
    __syncthreads();

    const float inv_d = 1.0f / (float)d;
    const float ta = p.tau * alpha;
    const uint32_t n_total = (uint32_t)p.n_samples * dd;
    Tile4 acc;
#pragma unroll
    for (int a = 0; a < 4; ++a) { acc.c[a][0] = 0ull; acc.c[a][1] = 0ull; }

    for (int r = 0; r < rounds; ++r) {
        const int a = blockIdx.y * rounds + r;
        if (a >= p.n_samples) break;                     // CTA-uniform
        for (int i = tid / d, j = tid - (tid / d) * d, e = tid; e < dd; e += blockDim.x) {
            float g = 0.0f;
            if (i != j) {
                const uint32_t bits = jax_bits(key, (uint32_t)a * dd + e, n_total, p.partitionable);
                g = entry_from_bits<false>(bits, sS[e], fast_soft, p.tau);
            }
            const float mz = (i == j ? 1.0f : 0.0f) + inv_d * g;            // graph_utils.py:22-25
            sG[i * LD + j] = g; sZ[i * LD + j] = mz; sZt[j * LD + i] = mz;
            j += blockDim.x;
            while (j >= d) { j -= d; ++i; }
        }
        __syncthreads();
        // E = M^(d-1): binary exponentiation, least-significant bit first (jnp.linalg.matrix_power)
        bool have_res = false;
        int n = d - 1;
        while (n > 0) {
            if (n & 1) {
                if (!have_res) {
                    for (int e = tid; e < MAT; e += blockDim.x) sRt[e] = sZt[e];
                    have_res = true;
                    __syncthreads();
                } else {
                    Tile4 c;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { c.c[q][0] = 0ull; c.c[q][1] = 0ull; }
                    if (active) tile4_mm(sRt, sZ, LD, ty, tx, 0, d, c);      // res * Z
                    __syncthreads();
                    if (active) tile4_store_t(c, sRt, LD, ty, tx);
                    __syncthreads();
                }
            }
            n >>= 1;
            if (n > 0) {
                Tile4 c;
#pragma unroll
                for (int q = 0; q < 4; ++q) { c.c[q][0] = 0ull; c.c[q][1] = 0ull; }
                if (active) tile4_mm(sZt, sZ, LD, ty, tx, 0, d, c);          // Z * Z
                __syncthreads();
                if (active) { tile4_store(c, sZ, LD, ty, tx); tile4_store_t(c, sZt, LD, ty, tx); }
                __syncthreads();
            }
        }
        // dS[a][b] = E[b][a] * tau alpha g_ab (1 - g_ab) = Rt[a][b] * F[a][b]
        if (active) {
#pragma unroll
            for (int a4 = 0; a4 < 4; ++a4) {
                const int row = 4 * ty + a4;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col = 4 * tx + 2 * q;
                    const float2 g = *reinterpret_cast<const float2*>(&sG[row * LD + col]);
                    const float2 e = *reinterpret_cast<const float2*>(&sRt[row * LD + col]);
                    const f32x2 f = pack2(ta * g.x * (1.0f - g.x), ta * g.y * (1.0f - g.y));
                    acc.c[a4][q] = fma2(pack2(e.x, e.y), f, acc.c[a4][q]);
                }
            }
        }
        __syncthreads();
    }
    float* outp = p.ds_out + ((size_t)m * gridDim.y + blockIdx.y) * dd;
    if (active) {
#pragma unroll
        for (int a4 = 0; a4 < 4; ++a4)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int row = 4 * ty + a4, col = 4 * tx + b;
                if (row < d && col < d) outp[row * d + col] = tile4_get(acc, a4, b);
            }
    }
    fuse_arrive(p.fuse, m, smem);
}

// ------------------------------------------------------------------------------------------
// LinearGaussian MC pass, QR-factor form, any n_vars with 5 LD^2 floats of shared memory (n_vars <= 104)
// ------------------------------------------------------------------------------------------
struct LinDenseShape { int ld, nt, threads, chunks, spc; size_t smem; };

static inline size_t lin_dense_smem(int d) {
    const int ld = dense_ld(d);
    return ((size_t)5 * ld * ld + (size_t)2 * (ld / 8) * ld + 2 * ld + 16) * sizeof(float);
}

// rx: [2][LD][LD] dense Rx (upper triangular) and its transpose, zero padded
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_mc_lin_dense(const __grid_constant__ McParams p, const float* __restrict__ rx, int LD, int NT) {
    extern __shared__ __align__(16) float smem[];
    constexpr bool HARD = (MODE == MC_THETA_HARD || MODE == MC_Z_SCORE);
    const int d = p.d, dd = d * d, HALF = LD / 2, TQ = HALF / 4, MAT = LD * LD;
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const int S = p.n_samples;
    const bool active = tid < NT;
    const int ty = active ? tid / TQ : 0, tx = active ? tid - ty * TQ : 0;

    float* sRx = smem;                      // Rx[i][k]
    float* sRxT = sRx + MAT;                // Rx^T[k][i]
    float* sTh = sRxT + MAT;                // Theta[i][j]
    float* sGm = sTh + MAT;                 // G[i][j] of the current sample
    float* sW = sGm + MAT;                  // U = I - G o Theta, then Y = Rx U
    float* sPart = sW + MAT;                // [2*TQ][LD] per-row-block column partials (prior, then ssq)
    float* sNode = sPart + (size_t)2 * TQ * LD;   // [LD]
    float* sLp = sNode + LD;                // [4]

    const bool use_ext = p.g_ext != nullptr;
    const bool fast_soft = !HARD && !use_ext && p.tau == 1.0f;
    const float alpha = p.alpha_linear * (float)t;
    for (int e = tid; e < 2 * MAT; e += blockDim.x) sRx[e] = rx[e];
    for (int e = tid; e < 3 * MAT; e += blockDim.x) sTh[e] = 0.0f;
    __syncthreads();
    {
        const float* throw_ = p.theta + (size_t)m * p.th_ld;
        for (int e = tid; e < dd; e += blockDim.x) { const int i = e / d, j = e - i * d; sTh[i * LD + j] = throw_[e]; }
    }
    const float* srow = p.scores ? p.scores + (size_t)m * dd : nullptr;
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);
    const float inv_s2 = 1.0f / p.s2, inv_se2 = 1.0f / p.sig2_edge;
    const uint32_t n_total = (uint32_t)S * dd;
    const float cst = (float)p.n_obs * p.log2pis2;

    Tile8 acc; tile_zero(acc);
    float m_run = -INFINITY, l_run = 0.0f, sum_lp = 0.0f;
    const int s_begin = c * p.s_per_chunk, s_end = min(S, s_begin + p.s_per_chunk);
    __syncthreads();

    for (int s = s_begin; s < s_end; ++s) {
        // ---- draw the graph; U = I - G o Theta
        for (int e = tid; e < dd; e += blockDim.x) {
            const int i = e / d, j = e - i * d;
            float g = 0.0f;
            if (i != j) {
                if (use_ext) g = p.g_ext[((size_t)m * S + s) * dd + e];
                else {
                    const float a = srow ? alpha * srow[e] : 0.0f;
                    const float sa = HARD ? sigmoidf_ref(a) : (fast_soft ? expf(-a) : a);
                    const uint32_t bits = jax_bits(key, (uint32_t)s * dd + e, n_total, p.partitionable);
                    g = entry_from_bits<HARD>(bits, sa, fast_soft, p.tau);
                }
            }
            sGm[i * LD + j] = g;
            sW[i * LD + j] = (i == j ? 1.0f : 0.0f) - g * sTh[i * LD + j];
        }
        __syncthreads();
        // ---- prior column partials of the tile rows: sum_i g logN(theta; mean_edge, sig_edge)  (linearGaussian.py:289)
        Tile8 y; tile_zero(y);
        if (active) {
            float pc[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) pc[b] = 0.0f;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const int row = tile_row(ty, HALF, a);
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int col = tile_col(tx, HALF, b);
                    pc[b] = fmaf(sGm[row * LD + col], norm_logpdf_pre(sTh[row * LD + col], p.mean_edge, p.sig2_edge, p.lognorm_edge), pc[b]);
                }
            }
            // Y = Rx U: rows i need k >= i (Rx upper triangular)
            tile_mm<true, false>(sRxT, sW, LD, HALF, ty, tx, min(d, 4 * ty), min(d, HALF + 4 * ty), y);
            tile_mm<true, true>(sRxT, sW, LD, HALF, ty, tx, min(d, HALF + 4 * ty), d, y);
            // stash the prior partials: two row blocks per thread row (lower / upper half share pc) -> one slot per ty
#pragma unroll
            for (int b = 0; b < 8; ++b) sPart[(size_t)ty * LD + tile_col(tx, HALF, b)] = pc[b];
        }
        __syncthreads();
        for (int j = tid; j < d; j += blockDim.x) {
            float prior_j = 0.0f;
            for (int r = 0; r < TQ; ++r) prior_j += sPart[(size_t)r * LD + j];
            sNode[j] = prior_j;
        }
        __syncthreads();
        if (active) {
            tile_store(y, sW, LD, HALF, ty, tx);                     // W <- Y
            float sq[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                float v = 0.0f;
#pragma unroll
                for (int a = 0; a < 8; ++a) { const float yv = tile_get(y, a, b); v = fmaf(yv, yv, v); }
                sq[b] = v;
            }
#pragma unroll
            for (int b = 0; b < 8; ++b) sPart[(size_t)ty * LD + tile_col(tx, HALF, b)] = sq[b];
        }
        __syncthreads();
        for (int j = tid; j < d; j += blockDim.x) {
            float ssq = 0.0f;
            for (int r = 0; r < TQ; ++r) ssq += sPart[(size_t)r * LD + j];
            sNode[j] = sNode[j] - 0.5f * (cst + ssq * inv_s2);
        }
        // B = Rx^T Y: rows i need k <= i
        Tile8 bt; tile_zero(bt);
        if (active && MODE != MC_Z_SCORE && MODE != MC_LP_ONLY) {
            tile_mm<true, true>(sRx, sW, LD, HALF, ty, tx, 0, min(d, 4 * ty + 4), bt);
            tile_mm<false, true>(sRx, sW, LD, HALF, ty, tx, 4 * ty + 4, min(d, HALF + 4 * ty + 4), bt);
        }
        __syncthreads();
        if (tid == 0) {
            float lp = 0.0f;
            for (int j = 0; j < d; ++j) lp += sNode[j];
            if (p.lp_out) p.lp_out[(size_t)m * S + s] = lp;
            sLp[0] = lp;
        }
        __syncthreads();
        if (MODE != MC_LP_ONLY) {
            const float lp = sLp[0];
            const float m_new = fmaxf(m_run, lp);
            const float scale = (m_run == -INFINITY) ? 0.0f : expf(m_run - m_new);
            const float e = expf(lp - m_new);
            l_run = l_run * scale + e;
            sum_lp += lp;
            m_run = m_new;
            if (active) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int row = tile_row(ty, HALF, a);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int col = tile_col(tx, HALF, 2 * q + u);
                            const float g = sGm[row * LD + col], th = sTh[row * LD + col];
                            const float bv = tile_get(bt, a, 2 * q + u);
                            if (MODE == MC_THETA_HARD) {
                                // d/dtheta: g * (-(theta-mu)/sig^2) + g * (x^T R)/s2          
                                v[u] = g * fmaf(bv, inv_s2, -(th - p.mean_edge) * inv_se2);
                            } else if (MODE == MC_Z_REPARAM) {
                                // dS = d lp/dG * tau*alpha*g(1-g)                                (App. B-4, B-6)
                                const float lpth = norm_logpdf_pre(th, p.mean_edge, p.sig2_edge, p.lognorm_edge);
                                v[u] = fmaf(th * inv_s2, bv, lpth) * (p.tau * alpha) * g * (1.0f - g);
                            } else {
                                v[u] = g;                                                       // score function (App. B-2)
                            }
                        }
                        const f32x2 sc = pack2(scale, scale);
                        acc.c[a][q] = fma2(acc.c[a][q], sc, pack2(e * v[0], e * v[1]));
                    }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == MC_LP_ONLY) return;
    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
    if (active) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int row = tile_row(ty, HALF, a);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int col = tile_col(tx, HALF, b);
                if (row < d && col < d) out[row * d + col] = tile_get(acc, a, b);
            }
        }
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = m_run; stv[1] = l_run; stv[2] = sum_lp; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, smem);
}

}  // namespace dibs
