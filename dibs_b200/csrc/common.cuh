// Shared device helpers: JAX-compatible threefry PRNG, math with the reference's rounding, reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace dibs {

// modes of a Monte-Carlo pass (kernels_mc*.cuh) -- also the estimator tag of the assemble step
enum { MC_THETA_HARD = 0, MC_Z_SCORE = 1, MC_Z_REPARAM = 2, MC_LP_ONLY = 3 };

// ------------------------------------------------------------------------------------------
// per-step device state: the loop carry of lax.fori_loop that kernels read (svgd.py:226,272)
// ------------------------------------------------------------------------------------------
struct StepState {
    uint32_t key[2];  // loop key at the start of the current step
    int32_t t;        // loop index (alpha(t), beta(t): dibs.py:70-71)
    int32_t pad;
};

// ------------------------------------------------------------------------------------------
// Threefry-2x32-20 (Salmon et al. SC'11) == jax._src.prng.threefry2x32
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ uint2 threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += k0; x1 += k1;
#define DIBS_TF_ROUND(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
    DIBS_TF_ROUND(13) DIBS_TF_ROUND(15) DIBS_TF_ROUND(26) DIBS_TF_ROUND(6)
    x0 += k1; x1 += k2 + 1u;
    DIBS_TF_ROUND(17) DIBS_TF_ROUND(29) DIBS_TF_ROUND(16) DIBS_TF_ROUND(24)
    x0 += k2; x1 += k0 + 2u;
    DIBS_TF_ROUND(13) DIBS_TF_ROUND(15) DIBS_TF_ROUND(26) DIBS_TF_ROUND(6)
    x0 += k0; x1 += k1 + 3u;
    DIBS_TF_ROUND(17) DIBS_TF_ROUND(29) DIBS_TF_ROUND(16) DIBS_TF_ROUND(24)
    x0 += k1; x1 += k2 + 4u;
    DIBS_TF_ROUND(13) DIBS_TF_ROUND(15) DIBS_TF_ROUND(26) DIBS_TF_ROUND(6)
    x0 += k2; x1 += k0 + 5u;
#undef DIBS_TF_ROUND
    return make_uint2(x0, x1);
}

// N independent blocks in lock step: the 20-round chain of one block is strictly serial (3 dependent ALU ops per
// round), so kernels that draw many entries per thread interleave N chains to fill the 4-cycle ALU latency.
template <int N>
__device__ __forceinline__ void threefry2x32_n(uint32_t k0, uint32_t k1, uint32_t (&x0)[N], uint32_t (&x1)[N]) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define DIBS_TFN_ROUND(r)                                             \
    _Pragma("unroll") for (int q = 0; q < N; ++q) {                   \
        x0[q] += x1[q]; x1[q] = rotl32(x1[q], r); x1[q] ^= x0[q];    \
    }
#define DIBS_TFN_INJECT(a, b, c)                                      \
    _Pragma("unroll") for (int q = 0; q < N; ++q) { x0[q] += (a); x1[q] += (b) + (c); }
    DIBS_TFN_INJECT(k0, k1, 0u)
    DIBS_TFN_ROUND(13) DIBS_TFN_ROUND(15) DIBS_TFN_ROUND(26) DIBS_TFN_ROUND(6)
    DIBS_TFN_INJECT(k1, k2, 1u)
    DIBS_TFN_ROUND(17) DIBS_TFN_ROUND(29) DIBS_TFN_ROUND(16) DIBS_TFN_ROUND(24)
    DIBS_TFN_INJECT(k2, k0, 2u)
    DIBS_TFN_ROUND(13) DIBS_TFN_ROUND(15) DIBS_TFN_ROUND(26) DIBS_TFN_ROUND(6)
    DIBS_TFN_INJECT(k0, k1, 3u)
    DIBS_TFN_ROUND(17) DIBS_TFN_ROUND(29) DIBS_TFN_ROUND(16) DIBS_TFN_ROUND(24)
    DIBS_TFN_INJECT(k1, k2, 4u)
    DIBS_TFN_ROUND(13) DIBS_TFN_ROUND(15) DIBS_TFN_ROUND(26) DIBS_TFN_ROUND(6)
    DIBS_TFN_INJECT(k2, k0, 5u)
#undef DIBS_TFN_ROUND
#undef DIBS_TFN_INJECT
}

// random_bits(key, shape)[e] for a flat array of n 32-bit draws.
// legacy layout (jax_threefry_partitionable=False): counters arange(n) padded to even length,
// first half -> lane 0, second half -> lane 1 (jax._src.prng.threefry_2x32).
__host__ __device__ __forceinline__ uint32_t jax_bits(uint2 key, uint32_t e, uint32_t n, bool partitionable) {
    if (partitionable) {
        uint2 r = threefry2x32(key.x, key.y, 0u, e);
        return r.x ^ r.y;
    }
    const uint32_t h = (n + 1u) >> 1;
    if (e < h) {
        uint32_t c1 = e + h;
        if (c1 >= n) c1 = 0u;  // the zero pad of an odd-length counter array
        return threefry2x32(key.x, key.y, e, c1).x;
    }
    return threefry2x32(key.x, key.y, e - h, e).y;
}

// Both lanes of the block that holds flat element e (< h): elements e and e+h. Legacy layout, n even.
__host__ __device__ __forceinline__ uint2 jax_bits_pair(uint2 key, uint32_t e, uint32_t h) {
    return threefry2x32(key.x, key.y, e, e + h);
}

// row r of random.split(key, num)
__host__ __device__ __forceinline__ uint2 jax_split_row(uint2 key, uint32_t r, uint32_t num, bool partitionable) {
    if (partitionable) return threefry2x32(key.x, key.y, 0u, r);
    return make_uint2(jax_bits(key, 2u * r, 2u * num, false), jax_bits(key, 2u * r + 1u, 2u * num, false));
}

// bitcast((bits >> 9) | 0x3F800000) - 1.0  in [0, 1)
__host__ __device__ __forceinline__ float bits_to_unit(uint32_t bits) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
#else
    return (float)(bits >> 9) * (1.0f / 8388608.0f);
#endif
}

#ifdef __CUDACC__
// random.logistic: u = max(eps, f*(1-eps)+eps) (separate fp32 multiply and add like the CPU reference),
// then log(u) - log1p(-u).
__device__ __forceinline__ float logistic_from_bits(uint32_t bits) {
    const float eps = 1.1920928955078125e-07f;
    float f = bits_to_unit(bits);
    float u = fmaxf(eps, __fadd_rn(__fmul_rn(f, 1.0f - eps), eps));
    return logf(u) - log1pf(-u);
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// XLA's fp32 erf_inv (Giles, "Approximating the erfinv function") -- random.normal = sqrt(2) * erf_inv(u)
__device__ __forceinline__ float erfinv_xla(float x) {
    float w = -log1pf(-x * x);
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = fmaf(p, w, 3.43273939e-07f);
        p = fmaf(p, w, -3.5233877e-06f);
        p = fmaf(p, w, -4.39150654e-06f);
        p = fmaf(p, w, 0.00021858087f);
        p = fmaf(p, w, -0.00125372503f);
        p = fmaf(p, w, -0.00417768164f);
        p = fmaf(p, w, 0.246640727f);
        p = fmaf(p, w, 1.50140941f);
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = fmaf(p, w, 0.000100950558f);
        p = fmaf(p, w, 0.00134934322f);
        p = fmaf(p, w, -0.00367342844f);
        p = fmaf(p, w, 0.00573950773f);
        p = fmaf(p, w, -0.0076224613f);
        p = fmaf(p, w, 0.00943887047f);
        p = fmaf(p, w, 1.00167406f);
        p = fmaf(p, w, 2.83297682f);
    }
    return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : p * x;
}

__device__ __forceinline__ float normal_from_bits(uint32_t bits) {
    const float lo = -0.99999994f;  // nextafter(-1, 0)
    float f = bits_to_unit(bits);
    float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));  // (maxval - minval) rounds to 2.0f in fp32
    return 1.41421356237309504880f * erfinv_xla(u);
}

// ------------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): two IEEE-rounded fp32 operations per lane per instruction.
// Same FLOP peak as scalar FFMA but half the issue slots -- these kernels are issue-bound, not pipe-bound.
// Each element is bit-identical to the scalar fmaf / + / -.
// ------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;   // {lo, hi} = two floats in an aligned register pair
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(f32x2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// The key a pass of the current step uses: `which` = index of the (M+1)-way split inside _svgd_step
// (joint: 0 theta, 1 z-likelihood, 2 prior -- svgd.py:695,699,703; marginal: 0 z-likelihood, 1 prior -- :245,251).
// Returns the sub-key of global particle m; `*next` (optional) receives the carried key after that split.
__device__ __forceinline__ uint2 step_particle_key(const StepState* st, int which, uint32_t m, uint32_t n_particles,
                                                   bool partitionable, uint2* next = nullptr) {
    uint2 key = make_uint2(st->key[0], st->key[1]);
    for (int w = 0; w < which; ++w) key = jax_split_row(key, 0u, n_particles + 1u, partitionable);
    if (next) *next = jax_split_row(key, 0u, n_particles + 1u, partitionable);
    return jax_split_row(key, m + 1u, n_particles + 1u, partitionable);
}

#endif

}  // namespace dibs
