// Particle initialisation: Z ~ N(0, sigma_z^2), Theta ~ likelihood_model.sample_parameters.
//
// replaces: dibs/inference/svgd.py:125-148, 489-515 (_sample_initial_random_particles),
// dibs/models/linearGaussian.py:212-227 (LinearGaussian.sample_parameters),
// dibs/models/nonlinearGaussian.py:155-186 + stax.serial / Dense init (jax.example_libraries.stax).
// `key` is the key handed to _sample_initial_random_particles:
//   key, subk = split(key);  z = normal(subk, (M, d, k, 2)) * std
//   key, subk = split(key);  theta = sample_parameters(key=subk, ...)
#pragma once
#include "common.cuh"

namespace dibs {

struct InitParams {
    const uint32_t* key;
    int M, d, k, lik, hidden, dth, partitionable;
    float std_z, mean_edge, sig_edge, min_edge, sig_param;
    float* z; float* theta;
};

__global__ void __launch_bounds__(256) k_init_z(InitParams p) {
    const uint2 key = make_uint2(p.key[0], p.key[1]);
    const uint2 subk = jax_split_row(key, 1u, 2u, p.partitionable);
    const uint32_t n = (uint32_t)p.M * p.d * p.k * 2u;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
        p.z[e] = normal_from_bits(jax_bits(subk, e, n, p.partitionable)) * p.std_z;
}

__device__ __forceinline__ uint2 theta_key(const InitParams& p) {
    const uint2 key = make_uint2(p.key[0], p.key[1]);
    const uint2 k1 = jax_split_row(key, 0u, 2u, p.partitionable);
    return jax_split_row(k1, 1u, 2u, p.partitionable);
}

__global__ void __launch_bounds__(256) k_init_theta_lin(InitParams p) {
    const uint2 subk = theta_key(p);
    const uint32_t n = (uint32_t)p.M * p.d * p.d;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        // theta = mean_edge + sig_edge * normal; theta += sign(theta) * min_edge   (linearGaussian.py:225-226)
        float th = __fadd_rn(p.mean_edge, __fmul_rn(p.sig_edge, normal_from_bits(jax_bits(subk, e, n, p.partitionable))));
        float sg = th > 0.0f ? 1.0f : (th < 0.0f ? -1.0f : 0.0f);
        p.theta[e] = __fadd_rn(th, __fmul_rn(sg, p.min_edge));
    }
}

// one thread per (particle, node): subkeys = split(key, M*d); stax.serial(Dense(H), Relu, Dense(1)) init
__global__ void __launch_bounds__(128) k_init_theta_nn(InitParams p) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= p.M * p.d) return;
    const int m = id / p.d, j = id % p.d, d = p.d, H = p.hidden;
    const bool pt = p.partitionable;
    uint2 rng = jax_split_row(theta_key(p), (uint32_t)id, (uint32_t)(p.M * p.d), pt);
    float* th = p.theta + (size_t)m * p.dth;
    const int oB1 = d * d * H, oW2 = oB1 + d * H, oB2 = oW2 + d * H;
    // layer 0: Dense(H)
    uint2 layer = jax_split_row(rng, 1u, 2u, pt); rng = jax_split_row(rng, 0u, 2u, pt);
    uint2 k1 = jax_split_row(layer, 0u, 2u, pt), k2 = jax_split_row(layer, 1u, 2u, pt);
    for (int e = 0; e < d * H; ++e)   // W[(i, h)] -> W1[j, i, h]
        th[(size_t)j * d * H + e] = normal_from_bits(jax_bits(k1, (uint32_t)e, (uint32_t)(d * H), pt)) * p.sig_param;
    for (int h = 0; h < H; ++h) th[oB1 + j * H + h] = normal_from_bits(jax_bits(k2, (uint32_t)h, (uint32_t)H, pt)) * p.sig_param;
    // layer 1: activation (consumes a split, no parameters)
    rng = jax_split_row(rng, 0u, 2u, pt);
    // layer 2: Dense(1)
    layer = jax_split_row(rng, 1u, 2u, pt);
    k1 = jax_split_row(layer, 0u, 2u, pt); k2 = jax_split_row(layer, 1u, 2u, pt);
    for (int h = 0; h < H; ++h) th[oW2 + j * H + h] = normal_from_bits(jax_bits(k1, (uint32_t)h, (uint32_t)H, pt)) * p.sig_param;
    th[oB2 + j] = normal_from_bits(jax_bits(k2, 0u, 1u, pt)) * p.sig_param;
}

}  // namespace dibs
