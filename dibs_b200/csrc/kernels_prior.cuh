// Latent prior score: acyclicity-constraint power series over Gumbel-soft graphs, graph prior through the
// edge probabilities, Gaussian prior on Z; plus the per-particle "assemble" step that merges the MC-pass
// partials and applies the chain rule through S = U V^T once.
//
// replaces: dibs/inference/dibs.py:557-658 (constraint_gumbel, grad_constraint_gumbel,
// log_graph_prior_particle, eltwise_grad_latent_prior), dibs/graph_utils.py:8-28 (acyclic_constr_nograd),
// dibs/models/graph.py:93-108,182-196,263-276 (unnormalized_log_prob_soft), and the final reshapes of
// dibs.py:376-385,451-457,531-549.
#pragma once
#include "common.cuh"
#include "kernels_peer.cuh"
#include "assemble.cuh"

namespace dibs {


// ------------------------------------------------------------------------------------------
// step prologue: per particle, the raw scores U V^T (edge-probability pass, dibs.py:179-181) and the sub-keys of
// every pass of the step (svgd.py:245,251 / 695,699,703 and the pre-draw splits of dibs.py:350,430), so that
// the Monte-Carlo / acyclicity / assemble kernels neither re-derive keys nor redo the d x k x d contraction.
// ------------------------------------------------------------------------------------------
struct PrologueParams {
    const float* z; int z_ld;
    int n_local, m_offset, n_particles, d, k;
    const StepState* st;                 // step loop: keys derive from the loop key ...
    const uint32_t* keys_in;             // ... hooks: [n_local][2] sub-keys handed in by the caller (n_splits = 1)
    int n_splits; uint32_t pre_split_mask; int partitionable;
    float* scores;                       // [n_local][d*d]
    uint32_t* keys_out;                  // [n_splits][n_local][2]
};

__global__ void __launch_bounds__(128) k_prologue(PrologueParams p) {
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, k = p.k, m = blockIdx.x, tid = threadIdx.x;
    if (tid < p.n_splits && p.keys_out) {
        uint2 key;
        if (p.keys_in) key = make_uint2(p.keys_in[2 * m], p.keys_in[2 * m + 1]);
        else key = step_particle_key(p.st, tid, (uint32_t)(p.m_offset + m), (uint32_t)p.n_particles, p.partitionable != 0);
        if ((p.pre_split_mask >> tid) & 1u) key = jax_split_row(key, 1u, 2u, p.partitionable != 0);
        uint32_t* o = p.keys_out + ((size_t)tid * p.n_local + m) * 2;
        o[0] = key.x; o[1] = key.y;
    }
    if (!p.scores) return;
    // U -> sU[kk][i], V -> sV[kk][j]: de-interleaved and transposed so the contraction reads conflict-free
    float* sU = smem; float* sV = smem + k * d;
    const float* zrow = p.z + (size_t)m * p.z_ld;
    for (int e = tid; e < 2 * d * k; e += blockDim.x) {
        const int i = e / (2 * k), r = e - i * 2 * k, kk = r >> 1;
        ((r & 1) ? sV : sU)[kk * d + i] = zrow[e];
    }
    __syncthreads();
    float* out = p.scores + (size_t)m * d * d;
    for (int e = tid; e < d * d; e += blockDim.x) {
        const int i = e / d, j = e - i * d;
        float acc = 0.0f;
        for (int kk = 0; kk < k; ++kk) acc = fmaf(sU[kk * d + i], sV[kk * d + j], acc);
        out[e] = acc;
    }
}

struct AcycParams {
    const float* z; int z_ld;
    const float* scores;                 // [n_local][d*d] raw U V^T (k_prologue)
    int n_local, m_offset, n_particles;
    int d, k, n_samples;                 // A
    const StepState* st; int which_split; int partitionable;
    const uint32_t* keys_override; int t_override;
    float alpha_linear, tau;
    float* ds_out;                       // [n_local][d*d]: sum over samples of dh/dS (before the 1/A mean)
    FuseAsm fuse;                        // step loop: the particle's last gradient CTA runs the assemble step
};

// C = A * B for d x d row-major matrices with leading dimension ld, cooperatively by a group of threads.
__device__ __forceinline__ void group_matmul(const float* __restrict__ A, const float* __restrict__ B,
                                             float* __restrict__ C, int d, int ld, int lane, int gsize) {
    for (int e = lane; e < d * d; e += gsize) {
        int i = e / d, j = e % d;
        float acc = 0.0f;
        for (int kk = 0; kk < d; ++kk) acc = fmaf(A[i * ld + kk], B[kk * ld + j], acc);
        C[i * ld + j] = acc;
    }
}

// One group (a warp when WARP_GROUP, else the whole CTA) handles one soft-graph sample at a time:
// G = sigmoid(tau (eps + alpha S)); E = (I + G/d)^(d-1); dS += E^T o tau alpha G (1-G)   
template <bool WARP_GROUP>
__global__ void __launch_bounds__(256) k_acyclic_grad(const __grid_constant__ AcycParams p) {
    extern __shared__ __align__(16) float smem[];
    const int d = p.d, k = p.k, ld = d | 1;   // odd leading dimension: conflict-free column walks
    const int m = blockIdx.x, tid = threadIdx.x;
    const int t = p.st ? p.st->t : p.t_override;
    const float alpha = p.alpha_linear * (float)t;
    const int n_groups = WARP_GROUP ? (blockDim.x >> 5) : 1;
    const int grp = WARP_GROUP ? (tid >> 5) : 0;
    const int lane = WARP_GROUP ? (tid & 31) : tid;
    const int gsize = WARP_GROUP ? 32 : blockDim.x;
    const int mat = d * ld;

    float* sS = smem;                               // [d*d] alpha * scores
    float* sZ = sS + d * d;                         // [2*d*k]
    float* gbase = sZ + 2 * d * k + (size_t)grp * 6 * mat;
    float* bG = gbase;                              // soft graph
    float* bZ0 = gbase + mat;                       // running square
    float* bZ1 = gbase + 2 * mat;
    float* bR0 = gbase + 3 * mat;                   // running result
    float* bR1 = gbase + 4 * mat;
    float* bAcc = gbase + 5 * mat;                  // per-group accumulator of dS

    for (int e = tid; e < d * d; e += blockDim.x) sS[e] = alpha * p.scores[(size_t)m * d * d + e];
    for (int e = lane; e < mat; e += gsize) bAcc[e] = 0.0f;
    const uint2 key = make_uint2(p.keys_override[2 * m], p.keys_override[2 * m + 1]);
    __syncthreads();

    const float inv_d = 1.0f / (float)d;
    const uint32_t n_total = (uint32_t)p.n_samples * d * d;
    for (int a = grp; a < p.n_samples; a += n_groups) {
        // soft graph and M = I + G/d  (graph_utils.py:22-25)
        for (int e = lane; e < d * d; e += gsize) {
            int i = e / d, j = e % d;
            float g = 0.0f;
            if (i != j) {
                uint32_t bits = jax_bits(key, (uint32_t)a * d * d + e, n_total, p.partitionable);
                g = sigmoidf_ref(p.tau * (logistic_from_bits(bits) + sS[e]));
            }
            bG[i * ld + j] = g;
            bZ0[i * ld + j] = (i == j ? 1.0f : 0.0f) + inv_d * g;
        }
        if (WARP_GROUP) __syncwarp(); else __syncthreads();
        // E = M^(d-1) by binary exponentiation (LSB first, like jnp.linalg.matrix_power)
        float* zc = bZ0; float* zn = bZ1; float* rc = nullptr; float* rn = bR0;
        int n = d - 1;
        while (n > 0) {
            if (n & 1) {
                if (rc == nullptr) {
                    for (int e = lane; e < mat; e += gsize) rn[e] = zc[e];
                } else {
                    group_matmul(rc, zc, rn, d, ld, lane, gsize);
                }
                float* tmp = rc; rc = rn; rn = (tmp == nullptr) ? bR1 : tmp;
                if (WARP_GROUP) __syncwarp(); else __syncthreads();
            }
            n >>= 1;
            if (n > 0) {
                group_matmul(zc, zc, zn, d, ld, lane, gsize);
                float* tmp = zc; zc = zn; zn = tmp;
                if (WARP_GROUP) __syncwarp(); else __syncthreads();
            }
        }
        // dh/dG = E^T (d * (1/d) = 1); through the sigmoid: tau * alpha * g (1 - g); diagonal is masked
        for (int e = lane; e < d * d; e += gsize) {
            int i = e / d, j = e % d;
            float g = bG[i * ld + j];
            float e_t = (d == 1) ? 1.0f : rc[j * ld + i];
            if (i != j) bAcc[i * ld + j] += e_t * (p.tau * alpha) * g * (1.0f - g);
        }
        if (WARP_GROUP) __syncwarp(); else __syncthreads();
    }
    __syncthreads();
    float* out = p.ds_out + (size_t)m * d * d;
    for (int e = tid; e < d * d; e += blockDim.x) {
        int i = e / d, j = e % d;
        float sum = 0.0f;
        for (int g = 0; g < n_groups; ++g) sum += smem[d * d + 2 * d * k + (size_t)g * 6 * mat + 5 * mat + i * ld + j];
        out[e] = sum;
    }
    fuse_arrive(p.fuse, m, smem);
}

inline size_t acyclic_smem(int d, int k, int n_groups) {
    return ((size_t)d * d + 2 * (size_t)d * k + (size_t)n_groups * 6 * d * (d | 1)) * sizeof(float);
}

// h(G) = tr((I + G/d)^d) - d for caller-supplied graphs (hook for graph_utils.py:8-28); one CTA per graph.
__global__ void __launch_bounds__(256) k_acyclic_value(const float* g, int n, int d, float* h_out) {
    extern __shared__ __align__(16) float smem[];
    const int ld = d | 1, mat = d * ld, tid = threadIdx.x;
    float* bZ0 = smem; float* bZ1 = smem + mat; float* bR0 = smem + 2 * mat; float* bR1 = smem + 3 * mat;
    const float* gm = g + (size_t)blockIdx.x * d * d;
    const float inv_d = 1.0f / (float)d;
    for (int e = tid; e < d * d; e += blockDim.x) {
        int i = e / d, j = e % d;
        bZ0[i * ld + j] = (i == j ? 1.0f : 0.0f) + inv_d * gm[e];
    }
    __syncthreads();
    float* zc = bZ0; float* zn = bZ1; float* rc = nullptr; float* rn = bR0;
    int nn = d;
    while (nn > 0) {
        if (nn & 1) {
            if (rc == nullptr) { for (int e = tid; e < mat; e += blockDim.x) rn[e] = zc[e]; }
            else group_matmul(rc, zc, rn, d, ld, tid, blockDim.x);
            float* tmp = rc; rc = rn; rn = (tmp == nullptr) ? bR1 : tmp;
            __syncthreads();
        }
        nn >>= 1;
        if (nn > 0) {
            group_matmul(zc, zc, zn, d, ld, tid, blockDim.x);
            float* tmp = zc; zc = zn; zn = tmp;
            __syncthreads();
        }
    }
    if (tid < 32) {
        float tr = 0.0f;
        for (int i = tid; i < d; i += 32) tr += rc[i * ld + i];
        tr = warp_sum(tr);
        if (tid == 0) h_out[blockIdx.x] = tr - (float)d;
    }
}

// ------------------------------------------------------------------------------------------
// progress summary of a particle set for the callback path (dibs.py:661-692: the reference's visualize_callback
// computes particle_to_g_lim, edge_probs and `(elwise_acyclic_constr_nograd(gs) > 0).sum()` on the host side of a
// blocking transfer).  Here: one CTA per particle -> edge probabilities, G_lim, h(G_lim) with the same power loop
// as k_acyclic_value; a second kernel reduces over particles in fixed order into [#cyclic, n, mean edge probs];
// the caller copies that record to pinned host memory asynchronously -- the step stream never waits for the host.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_particle_summary(const float* z, int z_ld, int n, int d, int k, float alpha,
                                                          float* p_out, float* h_out) {
    extern __shared__ __align__(16) float smem[];
    const int ld = d | 1, mat = d * ld, tid = threadIdx.x;
    float* bZ0 = smem; float* bZ1 = smem + mat; float* bR0 = smem + 2 * mat; float* bR1 = smem + 3 * mat;
    float* sZ = smem + 4 * mat;                       // [2*d*k]
    const float* zrow = z + (size_t)blockIdx.x * z_ld;
    for (int e = tid; e < 2 * d * k; e += blockDim.x) sZ[e] = zrow[e];
    __syncthreads();
    const float inv_d = 1.0f / (float)d;
    for (int e = tid; e < d * d; e += blockDim.x) {
        const int i = e / d, j = e % d;
        float acc = 0.0f;
        for (int kk = 0; kk < k; ++kk) acc = fmaf(sZ[(i * k + kk) * 2], sZ[(j * k + kk) * 2 + 1], acc);
        p_out[(size_t)blockIdx.x * d * d + e] = (i == j) ? 0.0f : sigmoidf_ref(alpha * acc);
        const float g = (i != j && acc > 0.0f) ? 1.0f : 0.0f;                // particle_to_g_lim (dibs.py:84-99)
        bZ0[i * ld + j] = (i == j ? 1.0f : 0.0f) + inv_d * g;
    }
    __syncthreads();
    float* zc = bZ0; float* zn = bZ1; float* rc = nullptr; float* rn = bR0;
    int nn = d;
    while (nn > 0) {
        if (nn & 1) {
            if (rc == nullptr) { for (int e = tid; e < mat; e += blockDim.x) rn[e] = zc[e]; }
            else group_matmul(rc, zc, rn, d, ld, tid, blockDim.x);
            float* tmp = rc; rc = rn; rn = (tmp == nullptr) ? bR1 : tmp;
            __syncthreads();
        }
        nn >>= 1;
        if (nn > 0) {
            group_matmul(zc, zc, zn, d, ld, tid, blockDim.x);
            float* tmp = zc; zc = zn; zn = tmp;
            __syncthreads();
        }
    }
    if (tid < 32) {
        float tr = 0.0f;
        for (int i = tid; i < d; i += 32) tr += rc[i * ld + i];
        tr = warp_sum(tr);
        if (tid == 0) h_out[blockIdx.x] = tr - (float)d;
    }
}

// out[0] = #particles with h(G_lim) > 0, out[1] = n, out[2 + e] = mean over particles of edge_probs[e]
__global__ void __launch_bounds__(256) k_summary_reduce(const float* p_all, const float* h_all, int n, int dd, float* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < dd) {
        float s = 0.0f;
        for (int m = 0; m < n; ++m) s += p_all[(size_t)m * dd + e];
        out[2 + e] = s / (float)n;
    }
    if (e == 0) {
        int c = 0;
        for (int m = 0; m < n; ++m) c += h_all[m] > 0.0f ? 1 : 0;
        out[0] = (float)c; out[1] = (float)n;
    }
}

// ------------------------------------------------------------------------------------------
// assemble: partials -> d log p / dZ (and d/dTheta) for one particle.  Hooks: one CTA per particle runs the
// assemble step (assemble.cuh); the step loop runs it inside the gradient kernels (fuse_arrive)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble_grad(const __grid_constant__ AsmParams p) {
    extern __shared__ __align__(16) float smem[];
    assemble_particle(p, blockIdx.x, smem);
}

// edge_probs / particle_to_g_lim (dibs.py:84-99,168-184): the edge-probability pass as a stand-alone, HBM-bound
// kernel (algorithmic traffic: read Z once, 8dk bytes, write P once, 4d^2 bytes per particle; AI = d/6 flop/B).
// One WARP per particle, 8 particles in flight per CTA, grid-stride over particles:
//   * the latent row arrives as coalesced 128-bit loads and is de-interleaved into sU[kk][i], sV[kk][j] (leading
//     dimension LD = d rounded up to 4, so a thread's 4 consecutive i / j are one 128-bit shared-memory read);
//   * lane (ti, tj) owns a 4 x 4 tile of U V^T: two LDS.128 per 16 FMAs -- without the register tile the kernel
//     is bound by shared-memory bandwidth (2 loads per FMA) at ~12 % of the HBM roofline (measured, round 2);
//   * the d x d result is staged in shared memory and leaves as coalesced 128-bit stores.
constexpr int EP_WARPS = 8;     // particles in flight per CTA (fewer when a particle's tiles need more shared memory)

inline size_t edge_probs_smem_per_warp(int d, int k) {
    const int ld = (d + 3) & ~3;
    return ((size_t)2 * k * ld + (size_t)ld * ld + 4) * sizeof(float);
}
inline int edge_probs_warps(int d, int k) {
    int w = (int)((200 * 1024) / edge_probs_smem_per_warp(d, k));
    return w < 1 ? 1 : (w > EP_WARPS ? EP_WARPS : w);
}

// raw != 0: write the raw scores U V^T themselves (the step's edge-probability pass: the gradient kernels of the next
// step apply sigma(alpha .) with the reference's rounding in their prologues)
__global__ void __launch_bounds__(EP_WARPS * 32) k_edge_probs(const float* __restrict__ z, int z_ld, int n, int d, int k,
                                                             float alpha, float* __restrict__ p_out,
                                                             int32_t* __restrict__ g_lim_out, int raw) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ld = (d + 3) & ~3, tq = ld >> 2;                 // tq x tq register tiles of 4 x 4
    const int per_warp = 2 * k * ld + ld * ld + 4;
    float* sU = smem + (size_t)warp * per_warp;                // [k][ld]
    float* sV = sU + k * ld;                                   // [k][ld]
    float* sP = sV + k * ld;                                   // [d][ld] staged result
    const int dk2 = 2 * d * k, dd = d * d;
    const bool vec_in = ((z_ld & 3) == 0) && ((dk2 & 3) == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
    const int n_warps = blockDim.x >> 5;
    for (int m = blockIdx.x * n_warps + warp; m < n; m += gridDim.x * n_warps) {
        const float* zrow = z + (size_t)m * z_ld;
        // zero the padding columns once per particle (cheap; keeps the tiles branch-free)
        for (int e = lane; e < 2 * k * (ld - d); e += 32) {
            const int kk = e / (ld - d), c = d + e % (ld - d);
            (kk < k ? sU : sV)[(kk % k) * ld + c] = 0.0f;
        }
        if (vec_in) {
            for (int e4 = lane; e4 < dk2 / 4; e4 += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(zrow) + e4);
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = 4 * e4 + u, i = e / (2 * k), r = e - i * 2 * k;
                    ((r & 1) ? sV : sU)[(r >> 1) * ld + i] = vv[u];
                }
            }
        } else {
            for (int e = lane; e < dk2; e += 32) {
                const int i = e / (2 * k), r = e - i * 2 * k;
                ((r & 1) ? sV : sU)[(r >> 1) * ld + i] = zrow[e];
            }
        }
        __syncwarp();
        for (int tile = lane; tile < tq * tq; tile += 32) {
            const int ti = tile / tq, tj = tile - ti * tq;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
            const float* up = sU + 4 * ti;
            const float* vp = sV + 4 * tj;
            for (int kk = 0; kk < k; ++kk, up += ld, vp += ld) {
                const float4 u4 = *reinterpret_cast<const float4*>(up);
                const float4 v4 = *reinterpret_cast<const float4*>(vp);
                const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(uu[a], vv[b], acc[a][b]);     // same kk order as every other U V^T
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = 4 * ti + a, j = 4 * tj + b;
                    float v;
                    if (raw) v = acc[a][b];
                    // stand-alone probabilities: fast exponential + approximate division (~1e-6 relative; the bit-exact
                    // Bernoulli thresholds of the step never come from here)
                    else if (p_out) v = (i == j) ? 0.0f : __fdividef(1.0f, 1.0f + __expf(-alpha * acc[a][b]));
                    else v = __int_as_float((i != j && acc[a][b] > 0.0f) ? 1 : 0);
                    if (i < d && j < d) sP[i * d + j] = v;                                      // packed [d][d]
                }
        }
        __syncwarp();
        float* out = p_out ? p_out + (size_t)m * dd : reinterpret_cast<float*>(g_lim_out + (size_t)m * dd);
        if (((dd & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            for (int e4 = lane; e4 < dd / 4; e4 += 32)
                __stcs(reinterpret_cast<float4*>(out) + e4, *reinterpret_cast<const float4*>(sP + 4 * e4));
        } else {
            for (int e = lane; e < dd; e += 32) out[e] = sP[e];
        }
        __syncwarp();
    }
}

// sample_g / soft-graph hooks: materialise the graphs the fused kernels generate on the fly
__global__ void __launch_bounds__(256) k_sample_graphs(const float* p_or_z, int ld, const uint32_t* keys, int d, int k,
                                                       int n_samples, int hard, float alpha, float tau,
                                                       int partitionable, int32_t* g_hard, float* g_soft) {
    extern __shared__ __align__(16) float smem[];
    const int m = blockIdx.x, tid = threadIdx.x;
    float* sA = smem; float* sZ = smem + d * d;
    if (hard) {
        for (int e = tid; e < d * d; e += blockDim.x) sA[e] = p_or_z[(size_t)m * ld + e];
    } else {
        const float* zrow = p_or_z + (size_t)m * ld;
        for (int e = tid; e < 2 * d * k; e += blockDim.x) sZ[e] = zrow[e];
        __syncthreads();
        for (int e = tid; e < d * d; e += blockDim.x) {
            int i = e / d, j = e % d;
            float acc = 0.0f;
            for (int kk = 0; kk < k; ++kk) acc = fmaf(sZ[(i * k + kk) * 2], sZ[(j * k + kk) * 2 + 1], acc);
            sA[e] = alpha * acc;
        }
    }
    __syncthreads();
    const uint2 key = make_uint2(keys[2 * m], keys[2 * m + 1]);
    const uint32_t n_total = (uint32_t)n_samples * d * d;
    for (uint32_t e = tid; e < n_total; e += blockDim.x) {
        int ij = e % (d * d), i = ij / d, j = ij % d;
        uint32_t bits = jax_bits(key, e, n_total, partitionable);
        if (hard) g_hard[(size_t)m * n_total + e] = (i != j && bits_to_unit(bits) < sA[ij]) ? 1 : 0;
        else g_soft[(size_t)m * n_total + e] = (i == j) ? 0.0f : sigmoidf_ref(tau * (logistic_from_bits(bits) + sA[ij]));
    }
}

}  // namespace dibs
