// All-pairs part of the SVGD step: squared-exponential kernel matrix on (Z, Theta), the transform phi and
// the optimizer update, on a rank's slab of rows against all (gathered) particles.
//
// replaces: dibs/kernel.py:20-30,52-71; dibs/inference/svgd.py:165-224 (marginal), :537-670 (joint),
// :265,718-719 + jax.example_libraries.optimizers.{sgd,rmsprop}.
// Closed form: grad_{x_j} k(x_j, x_i) = -(2/h)(x_j - x_i) k_ji, kept in difference form
// like the reference's autodiff so that near-neighbours do not cancel.
#pragma once
#include "common.cuh"

namespace dibs {

struct PairParams {
    const float* x_all; int ld;          // all particles: row = [Z (dz) | Theta (dth) | ...], stride ld
    const float* g_all; int g_ld;        // log-prob gradients, same column layout
    int n_all;                           // M
    int row0, n_rows;                    // this rank's rows [row0, row0 + n_rows)
    int dz, dth;
    int n_split, split_len;              // split of the feature axis for the distance pass
    float* dist_part;                    // [n_split][2][n_rows][n_all] partial squared distances (z, theta)
    float* kz; float* kt; float* kfull;  // [n_rows][n_all]
    float h_z, h_t, scale_z, scale_t;
    // update
    float* x_next; int next_ld;          // where the updated local rows go (null: phi only)
    float* v; int v_ld;                  // RMSprop state [n_rows][dz+dth]
    float* phi_out; int phi_ld;          // optional [n_rows][dz+dth]
    int optimizer; float stepsize;
    StepState* st; int n_step_splits; int n_particles; int partitionable;   // advanced by block 0 when st != null
};

// ---- pass 1: partial squared distances, tile 64x64, 256 threads x (4x4), feature axis split across blockIdx.z
constexpr int KT = 64;   // tile edge
constexpr int KF = 32;   // features per smem stage

__global__ void __launch_bounds__(256) k_pair_dist(PairParams p) {
    __shared__ float sI[KF][KT + 1];
    __shared__ float sJ[KF][KT + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int j0 = blockIdx.x * KT, i0 = blockIdx.y * KT, sp = blockIdx.z;
    const int D = p.dz + p.dth;
    const int f_begin = sp * p.split_len, f_end = min(D, f_begin + p.split_len);
    float accz[4][4], acct[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { accz[a][b] = 0.0f; acct[a][b] = 0.0f; }

    for (int f0 = f_begin; f0 < f_end; f0 += KF) {
        // stage KF features of 64 i-rows and 64 j-rows (coalesced along the feature axis)
        for (int e = tid; e < KT * KF; e += 256) {
            int r = e / KF, f = e % KF;
            int fi = f0 + f;
            bool okf = fi < f_end;
            int gi = i0 + r, gj = j0 + r;
            sI[f][r] = (okf && gi < p.n_rows) ? p.x_all[(size_t)(p.row0 + gi) * p.ld + fi] : 0.0f;
            sJ[f][r] = (okf && gj < p.n_all) ? p.x_all[(size_t)gj * p.ld + fi] : 0.0f;
        }
        __syncthreads();
        const int nf = min(KF, f_end - f0);
        for (int f = 0; f < nf; ++f) {
            const bool is_z = (f0 + f) < p.dz;
            float xi[4], xj[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { xi[a] = sI[f][ty * 4 + a]; xj[a] = sJ[f][tx * 4 + a]; }
            if (is_z) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) { float df = xi[a] - xj[b]; accz[a][b] = fmaf(df, df, accz[a][b]); }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) { float df = xi[a] - xj[b]; acct[a][b] = fmaf(df, df, acct[a][b]); }
            }
        }
        __syncthreads();
    }
    const size_t plane = (size_t)p.n_rows * p.n_all;
    float* oz = p.dist_part + ((size_t)sp * 2) * plane;
    float* ot = oz + plane;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int gi = i0 + ty * 4 + a, gj = j0 + tx * 4 + b;
            if (gi < p.n_rows && gj < p.n_all) {
                oz[(size_t)gi * p.n_all + gj] = accz[a][b];
                ot[(size_t)gi * p.n_all + gj] = acct[a][b];
            }
        }
}

// ---- pass 2: sum the feature splits in fixed order, apply the SE kernels (kernel.py:30,66-71)
__global__ void __launch_bounds__(256) k_pair_finish(PairParams p) {
    const size_t plane = (size_t)p.n_rows * p.n_all;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
        float dz = 0.0f, dt = 0.0f;
        for (int s = 0; s < p.n_split; ++s) {
            dz += p.dist_part[((size_t)s * 2) * plane + e];
            dt += p.dist_part[((size_t)s * 2 + 1) * plane + e];
        }
        float kz = p.scale_z * expf(-dz / p.h_z);
        float kt = p.dth > 0 ? p.scale_t * expf(-dt / p.h_t) : 0.0f;
        p.kz[e] = kz;
        if (p.kt) p.kt[e] = kt;
        p.kfull[e] = kz + kt;
    }
}

// ---- pass 3: phi_i = -(1/M) sum_j [ K_ij g_j - (2/h) Kterm_ij (x_j - x_i) ]  + optimizer update
constexpr int PT_I = 32, PT_C = 32, PT_J = 32;   // tile: 32 rows x 32 feature columns, 64 threads x (4x4)

__global__ void __launch_bounds__(64) k_phi_update(PairParams p) {
    __shared__ float sK[PT_J][PT_I + 1];    // K_full[i][j] transposed: [j][i]
    __shared__ float sKt[PT_J][PT_I + 1];   // K term (z or theta block)
    __shared__ float sXj[PT_J][PT_C + 1];
    __shared__ float sGj[PT_J][PT_C + 1];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int i0 = blockIdx.y * PT_I;
    const int D = p.dz + p.dth;
    // a column tile never straddles the Z | Theta boundary: grid.x = ceil(dz/32) + ceil(dth/32)
    const int nzt = (p.dz + PT_C - 1) / PT_C;
    const bool z_block = (int)blockIdx.x < nzt;
    const int c0 = z_block ? blockIdx.x * PT_C : p.dz + ((int)blockIdx.x - nzt) * PT_C;
    const int c_end = z_block ? p.dz : D;
    const float* kterm = z_block ? p.kz : p.kt;
    const float h = z_block ? p.h_z : p.h_t;

    float xi[4][4], drive[4][4], rep[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int gi = i0 + ty * 4 + a, gc = c0 + tx * 4 + b;
            xi[a][b] = (gi < p.n_rows && gc < c_end) ? p.x_all[(size_t)(p.row0 + gi) * p.ld + gc] : 0.0f;
            drive[a][b] = 0.0f; rep[a][b] = 0.0f;
        }

    for (int j0 = 0; j0 < p.n_all; j0 += PT_J) {
        for (int e = tid; e < PT_J * PT_I; e += 64) {
            int i = e / PT_J, j = e % PT_J;      // coalesced along j (row-major K slab)
            int gi = i0 + i, gj = j0 + j;
            bool ok = gi < p.n_rows && gj < p.n_all;
            sK[j][i] = ok ? p.kfull[(size_t)gi * p.n_all + gj] : 0.0f;
            sKt[j][i] = ok ? kterm[(size_t)gi * p.n_all + gj] : 0.0f;
        }
        for (int e = tid; e < PT_J * PT_C; e += 64) {
            int j = e / PT_C, c = e % PT_C;
            int gj = j0 + j, gc = c0 + c;
            bool ok = gj < p.n_all && gc < c_end;
            sXj[j][c] = ok ? p.x_all[(size_t)gj * p.ld + gc] : 0.0f;
            sGj[j][c] = ok ? p.g_all[(size_t)gj * p.g_ld + gc] : 0.0f;
        }
        __syncthreads();
        const int nj = min(PT_J, p.n_all - j0);
        for (int j = 0; j < nj; ++j) {
            float kf[4], kt[4], xj[4], gj[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { kf[a] = sK[j][ty * 4 + a]; kt[a] = sKt[j][ty * 4 + a]; }
#pragma unroll
            for (int b = 0; b < 4; ++b) { xj[b] = sXj[j][tx * 4 + b]; gj[b] = sGj[j][tx * 4 + b]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    drive[a][b] = fmaf(kf[a], gj[b], drive[a][b]);
                    rep[a][b] = fmaf(kt[a], xj[b] - xi[a][b], rep[a][b]);
                }
        }
        __syncthreads();
    }
    const float inv_m = 1.0f / (float)p.n_all;
    const float c2 = -2.0f / h;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int gi = i0 + ty * 4 + a, gc = c0 + tx * 4 + b;
            if (gi >= p.n_rows || gc >= c_end) continue;
            // -(weighted_gradient_ascent + repulsion).mean(axis=0)   (svgd.py:212-216)
            float phi = -(drive[a][b] + c2 * rep[a][b]) * inv_m;
            if (p.phi_out) p.phi_out[(size_t)gi * p.phi_ld + gc] = phi;
            if (p.x_next) {
                float x = xi[a][b];
                if (p.optimizer == 1) {
                    // rmsprop(step, gamma=0.9, eps=1e-8): v = v*gamma + g^2*(1-gamma); x -= step*g/sqrt(v+eps)
                    float* vp = p.v + (size_t)gi * p.v_ld + gc;
                    float v = __fadd_rn(__fmul_rn(*vp, 0.9f), __fmul_rn(__fmul_rn(phi, phi), 0.1f));
                    *vp = v;
                    x = __fsub_rn(x, __fdiv_rn(__fmul_rn(p.stepsize, phi), __fsqrt_rn(__fadd_rn(v, 1e-8f))));
                } else {
                    x = __fsub_rn(x, __fmul_rn(p.stepsize, phi));   // sgd: x - step * g
                }
                p.x_next[(size_t)gi * p.next_ld + gc] = x;
            }
        }
    // carry the loop state: key <- after this step's (M+1)-way splits, t <- t + 1   (svgd.py:245,251,272)
    if (p.st && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        uint2 key = make_uint2(p.st->key[0], p.st->key[1]);
        for (int w = 0; w < p.n_step_splits; ++w) key = jax_split_row(key, 0u, (uint32_t)p.n_particles + 1u, p.partitionable);
        p.st->key[0] = key.x; p.st->key[1] = key.y;
        p.st->t += 1;
    }
}

}  // namespace dibs
