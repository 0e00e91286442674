// All-pairs part of the SVGD step: squared-exponential kernel matrix on (Z, Theta), the transform phi and
// the optimizer update, on a rank's slab of rows against all (gathered) particles.
//
// replaces: dibs/kernel.py:20-30,52-71; dibs/inference/svgd.py:165-224 (marginal), :537-670 (joint),
// :265,718-719 + jax.example_libraries.optimizers.{sgd,rmsprop}.
// Closed form: grad_{x_j} k(x_j, x_i) = -(2/h)(x_j - x_i) k_ji, kept in difference form
// like the reference's autodiff so that near-neighbours do not cancel.
#pragma once
#include "common.cuh"
#include "kernels_peer.cuh"

namespace dibs {

struct PairParams {
    const float* x_all; int ld;          // all particles: row = [Z (dz) | Theta (dth) | ...], stride ld
    const float* g_all; int g_ld;        // log-prob gradients, same column layout
    int n_all;                           // M
    int row0, n_rows;                    // this rank's rows [row0, row0 + n_rows)
    int dz, dth;
    int n_split, n_split_z;              // feature splits of the distance pass: the first n_split_z cut Z, the rest Theta
    int split_len_z, split_len_t;        // features per split (multiples of 32)
    float* dist_part;                    // [n_split][n_rows][n_all] partial squared distances
    int n_jsplit, j_len;                 // phi: slices of the j axis (j_len on the global index, multiple of 32)
    float* phi_part;                     // [n_jsplit][n_rows][dz+dth] per-slice sums of drive - (2/h) repulsion
    PeerWait wait_x, wait_g;             // peer-memory exchange: flags to wait on before reading x_all / g_all
    float* kz; float* kt; float* kfull;  // [n_rows][n_all]
    float h_z, h_t, scale_z, scale_t;
};

// ---- pass 1: partial squared distances.  Tile 64 x 64 outputs, 256 threads x (4 x 4) with INTERLEAVED ownership
// (rows ty + 16a, columns tx + 16b) so that 128-bit shared-memory reads along the feature axis are conflict-free
// with a row stride of 36 floats.  blockIdx.z = feature split; a split never straddles the Z | Theta boundary
// (the first n_split_z splits cut the Z features, the rest the Theta features).
constexpr int KT = 64;    // tile edge
constexpr int KF = 32;    // features per shared-memory stage
constexpr int KFP = 36;   // padded row stride (floats): 16-byte aligned, quarter-warp conflict-free

__global__ void __launch_bounds__(256) k_pair_dist(PairParams p) {
    __shared__ __align__(16) float sI[KT * KFP];
    __shared__ __align__(16) float sJ[KT * KFP];
    peer_wait(p.wait_x);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int j0 = blockIdx.x * KT, i0 = blockIdx.y * KT, sp = blockIdx.z;
    const bool z_part = sp < p.n_split_z;
    const int base = z_part ? 0 : p.dz;
    const int len = z_part ? p.split_len_z : p.split_len_t;
    const int f_begin = base + (z_part ? sp : sp - p.n_split_z) * len;
    const int f_end = min(z_part ? p.dz : p.dz + p.dth, f_begin + len);
    // accumulators are packed pairs {sum over even features, sum over odd features} (FFMA2), folded at the end
    f32x2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0ull;

    for (int f0 = f_begin; f0 < f_end; f0 += KF) {
        // stage KF features of 64 i-rows and 64 j-rows; a thread copies 2 x float4 per matrix (coalesced along f)
        for (int e = tid; e < KT * (KF / 4); e += 256) {
            const int r = e >> 3, f4 = (e & 7) * 4;
            const int fi = f0 + f4;
            const int gi = i0 + r, gj = j0 + r;
            float4 vi = make_float4(0.f, 0.f, 0.f, 0.f), vj = vi;
            if (gi < p.n_rows) {
                const float* src = p.x_all + (size_t)(p.row0 + gi) * p.ld + fi;
                if (fi + 3 < f_end && ((((size_t)(p.row0 + gi) * p.ld + fi) & 3) == 0)) vi = *reinterpret_cast<const float4*>(src);
                else {
                    if (fi < f_end) vi.x = src[0];
                    if (fi + 1 < f_end) vi.y = src[1];
                    if (fi + 2 < f_end) vi.z = src[2];
                    if (fi + 3 < f_end) vi.w = src[3];
                }
            }
            if (gj < p.n_all) {
                const float* src = p.x_all + (size_t)gj * p.ld + fi;
                if (fi + 3 < f_end && ((((size_t)gj * p.ld + fi) & 3) == 0)) vj = *reinterpret_cast<const float4*>(src);
                else {
                    if (fi < f_end) vj.x = src[0];
                    if (fi + 1 < f_end) vj.y = src[1];
                    if (fi + 2 < f_end) vj.z = src[2];
                    if (fi + 3 < f_end) vj.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(&sI[r * KFP + f4]) = vi;
            *reinterpret_cast<float4*>(&sJ[r * KFP + f4]) = vj;
        }
        __syncthreads();
#pragma unroll
        for (int f4 = 0; f4 < KF; f4 += 4) {
            ulonglong2 xi[4], xj[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xi[a] = *reinterpret_cast<const ulonglong2*>(&sI[(ty + 16 * a) * KFP + f4]);
#pragma unroll
            for (int b = 0; b < 4; ++b) xj[b] = *reinterpret_cast<const ulonglong2*>(&sJ[(tx + 16 * b) * KFP + f4]);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    f32x2 df;
                    df = sub2(xi[a].x, xj[b].x); acc[a][b] = fma2(df, df, acc[a][b]);
                    df = sub2(xi[a].y, xj[b].y); acc[a][b] = fma2(df, df, acc[a][b]);
                }
        }
        __syncthreads();
    }
    const size_t plane = (size_t)p.n_rows * p.n_all;
    float* o = p.dist_part + (size_t)sp * plane;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gi = i0 + ty + 16 * a, gj = j0 + tx + 16 * b;
            if (gi < p.n_rows && gj < p.n_all) o[(size_t)gi * p.n_all + gj] = lo2(acc[a][b]) + hi2(acc[a][b]);
        }
}

// ---- pass 2: sum the feature splits in fixed order, apply the SE kernels (kernel.py:30,66-71)
__global__ void __launch_bounds__(256) k_pair_finish(PairParams p) {
    const size_t plane = (size_t)p.n_rows * p.n_all;
    const int nz = p.n_split_z, nt = p.n_split - p.n_split_z;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
        float dz = 0.0f, dt = 0.0f;
#pragma unroll 4
        for (int s = 0; s < nz; ++s) dz += p.dist_part[(size_t)s * plane + e];
#pragma unroll 4
        for (int s = 0; s < nt; ++s) dt += p.dist_part[(size_t)(nz + s) * plane + e];
        float kz = p.scale_z * expf(-dz / p.h_z);
        float kt = p.dth > 0 ? p.scale_t * expf(-dt / p.h_t) : 0.0f;
        p.kz[e] = kz;
        if (p.kt) p.kt[e] = kt;
        p.kfull[e] = kz + kt;
    }
}

// ---- pass 3: partial sums of phi_i = -(1/M) sum_j [ K_ij g_j - (2/h) Kterm_ij (x_j - x_i) ] over one slice of j.
// Tile: 32 rows x 64 feature columns, 128 threads x (4 rows x 4 columns); per particle j a thread issues four
// 128-bit shared-memory loads (K, Kterm for its 4 rows; x_j, g_j for its 4 columns) for 48 FP instructions.
// blockIdx.z = slice of the j axis (fixed length j_len on the GLOBAL particle index, so the summation order --
// and with it every bit of the result -- does not depend on how many ranks share the particles).
constexpr int PT_I = 32, PT_C = 64, PT_J = 32;
constexpr int PT_KP = 36;   // padded stride of the transposed K tiles [j][i]

__global__ void __launch_bounds__(128) k_phi_partial(PairParams p) {
    // K tiles transposed to [j][i] with every entry DUPLICATED ({k, k}): the packed FFMA2 needs the row weight in
    // both halves of a register pair, and two 128-bit loads are cheaper than four register moves per (j, row)
    __shared__ __align__(16) float sK[PT_J * PT_KP * 2];     // K_full
    __shared__ __align__(16) float sKt[PT_J * PT_KP * 2];    // K term (z or theta block)
    __shared__ __align__(16) float sXj[PT_J * PT_C];
    __shared__ __align__(16) float sGj[PT_J * PT_C];
    peer_wait(p.wait_g);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * PT_I;
    const int D = p.dz + p.dth;
    // a column tile never straddles the Z | Theta boundary: grid.x = ceil(dz/64) + ceil(dth/64)
    const int nzt = (p.dz + PT_C - 1) / PT_C;
    const bool z_block = (int)blockIdx.x < nzt;
    const int c0 = z_block ? blockIdx.x * PT_C : p.dz + ((int)blockIdx.x - nzt) * PT_C;
    const int c_end = z_block ? p.dz : D;
    const float* kterm = z_block ? p.kz : p.kt;
    const float h = z_block ? p.h_z : p.h_t;
    const int j_begin = blockIdx.z * p.j_len;
    const int j_end = min(p.n_all, j_begin + p.j_len);

    f32x2 xi[4][2], drive[4][2], rep[4][2];      // [row][column pair]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int gi = i0 + ty * 4 + a, gc = c0 + tx * 4 + 2 * b;
            const float* xr = p.x_all + (size_t)(p.row0 + gi) * p.ld;
            const float v0 = (gi < p.n_rows && gc < c_end) ? xr[gc] : 0.0f;
            const float v1 = (gi < p.n_rows && gc + 1 < c_end) ? xr[gc + 1] : 0.0f;
            xi[a][b] = pack2(v0, v1);
            drive[a][b] = 0ull; rep[a][b] = 0ull;
        }

    for (int j0 = j_begin; j0 < j_end; j0 += PT_J) {
        // K tiles: coalesced along j in global memory, transposed into [j][i]
        for (int e = tid; e < PT_I * PT_J; e += 128) {
            int i = e / PT_J, j = e % PT_J;
            int gi = i0 + i, gj = j0 + j;
            bool ok = gi < p.n_rows && gj < j_end;
            const float kf = ok ? p.kfull[(size_t)gi * p.n_all + gj] : 0.0f;
            const float kt = ok ? kterm[(size_t)gi * p.n_all + gj] : 0.0f;
            *reinterpret_cast<float2*>(&sK[(j * PT_KP + i) * 2]) = make_float2(kf, kf);
            *reinterpret_cast<float2*>(&sKt[(j * PT_KP + i) * 2]) = make_float2(kt, kt);
        }
        for (int e = tid; e < PT_J * (PT_C / 4); e += 128) {
            const int j = e / (PT_C / 4), c4 = (e % (PT_C / 4)) * 4;
            const int gj = j0 + j, gc = c0 + c4;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
            if (gj < j_end) {
                const size_t ox = (size_t)gj * p.ld + gc, og = (size_t)gj * p.g_ld + gc;
                if (gc + 3 < c_end &&
                    ((reinterpret_cast<uintptr_t>(p.x_all + ox) | reinterpret_cast<uintptr_t>(p.g_all + og)) & 15) == 0) {
                    xv = *reinterpret_cast<const float4*>(p.x_all + ox);
                    gv = *reinterpret_cast<const float4*>(p.g_all + og);
                } else {
                    if (gc < c_end) { xv.x = p.x_all[ox]; gv.x = p.g_all[og]; }
                    if (gc + 1 < c_end) { xv.y = p.x_all[ox + 1]; gv.y = p.g_all[og + 1]; }
                    if (gc + 2 < c_end) { xv.z = p.x_all[ox + 2]; gv.z = p.g_all[og + 2]; }
                    if (gc + 3 < c_end) { xv.w = p.x_all[ox + 3]; gv.w = p.g_all[og + 3]; }
                }
            }
            *reinterpret_cast<float4*>(&sXj[j * PT_C + c4]) = xv;
            *reinterpret_cast<float4*>(&sGj[j * PT_C + c4]) = gv;
        }
        __syncthreads();
        const int nj = min(PT_J, j_end - j0);
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const ulonglong2 kfa = *reinterpret_cast<const ulonglong2*>(&sK[(j * PT_KP + ty * 4) * 2]);
            const ulonglong2 kfb = *reinterpret_cast<const ulonglong2*>(&sK[(j * PT_KP + ty * 4 + 2) * 2]);
            const ulonglong2 kta = *reinterpret_cast<const ulonglong2*>(&sKt[(j * PT_KP + ty * 4) * 2]);
            const ulonglong2 ktb = *reinterpret_cast<const ulonglong2*>(&sKt[(j * PT_KP + ty * 4 + 2) * 2]);
            const ulonglong2 xj = *reinterpret_cast<const ulonglong2*>(&sXj[j * PT_C + tx * 4]);
            const ulonglong2 gj = *reinterpret_cast<const ulonglong2*>(&sGj[j * PT_C + tx * 4]);
            const f32x2 kf[4] = {kfa.x, kfa.y, kfb.x, kfb.y}, kt[4] = {kta.x, kta.y, ktb.x, ktb.y};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                drive[a][0] = fma2(kf[a], gj.x, drive[a][0]);
                drive[a][1] = fma2(kf[a], gj.y, drive[a][1]);
                rep[a][0] = fma2(kt[a], sub2(xj.x, xi[a][0]), rep[a][0]);
                rep[a][1] = fma2(kt[a], sub2(xj.y, xi[a][1]), rep[a][1]);
            }
        }
        __syncthreads();
    }
    // weighted_gradient_ascent + repulsion of this j slice   (svgd.py:212-216); the mean and the sign follow in
    // k_opt_update once all slices are summed
    const float c2 = -2.0f / h;
    float* part = p.phi_part + (size_t)blockIdx.z * p.n_rows * D;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gi = i0 + ty * 4 + a, gc = c0 + tx * 4;
        if (gi >= p.n_rows) continue;
        float* o = part + (size_t)gi * D + gc;
        const float v0 = fmaf(c2, lo2(rep[a][0]), lo2(drive[a][0])), v1 = fmaf(c2, hi2(rep[a][0]), hi2(drive[a][0]));
        const float v2 = fmaf(c2, lo2(rep[a][1]), lo2(drive[a][1])), v3 = fmaf(c2, hi2(rep[a][1]), hi2(drive[a][1]));
        if (gc + 3 < c_end && ((((size_t)blockIdx.z * p.n_rows + gi) * D + gc) & 3) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
        } else {
            if (gc < c_end) o[0] = v0;
            if (gc + 1 < c_end) o[1] = v1;
            if (gc + 2 < c_end) o[2] = v2;
            if (gc + 3 < c_end) o[3] = v3;
        }
    }
}

// ---- pass 3, experiment (DIBS_B200_PHI_TILE=1; default off until timed): 64 rows x 64 columns per CTA, 128 threads x
// (8 rows x 4 columns).  Per particle j a thread issues six 128-bit loads (K and Kterm for its 8 rows, x_j and g_j for
// its 4 columns) for 48 packed FP instructions -- twice the arithmetic per shared-memory byte of the 32-row kernel;
// K tiles are stored once (no duplication), the row weight enters the FFMA2 as a broadcast scalar.
// Same j slicing and the same per-element accumulation order as k_phi_partial -> bit-identical partial sums.
constexpr int PB_I = 64;
constexpr int PB_KP = 68;   // padded stride of the transposed K tiles [j][i]

__global__ void __launch_bounds__(128) k_phi_partial_big(PairParams p) {
    __shared__ __align__(16) float sK[PT_J * PB_KP];      // K_full[i][j] transposed: [j][i]
    __shared__ __align__(16) float sKt[PT_J * PB_KP];     // K term (z or theta block)
    __shared__ __align__(16) float sXj[PT_J * PT_C];
    __shared__ __align__(16) float sGj[PT_J * PT_C];
    peer_wait(p.wait_g);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // ty 0..7: rows 8 ty .. 8 ty + 7
    const int i0 = blockIdx.y * PB_I;
    const int D = p.dz + p.dth;
    const int nzt = (p.dz + PT_C - 1) / PT_C;
    const bool z_block = (int)blockIdx.x < nzt;
    const int c0 = z_block ? blockIdx.x * PT_C : p.dz + ((int)blockIdx.x - nzt) * PT_C;
    const int c_end = z_block ? p.dz : D;
    const float* kterm = z_block ? p.kz : p.kt;
    const float h = z_block ? p.h_z : p.h_t;
    const int j_begin = blockIdx.z * p.j_len;
    const int j_end = min(p.n_all, j_begin + p.j_len);

    f32x2 xi[8][2], drive[8][2], rep[8][2];               // [row][column pair]
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int gi = i0 + ty * 8 + a, gc = c0 + tx * 4 + 2 * b;
            const float* xr = p.x_all + (size_t)(p.row0 + gi) * p.ld;
            const float v0 = (gi < p.n_rows && gc < c_end) ? xr[gc] : 0.0f;
            const float v1 = (gi < p.n_rows && gc + 1 < c_end) ? xr[gc + 1] : 0.0f;
            xi[a][b] = pack2(v0, v1);
            drive[a][b] = 0ull; rep[a][b] = 0ull;
        }

    for (int j0 = j_begin; j0 < j_end; j0 += PT_J) {
        for (int e = tid; e < PB_I * PT_J; e += 128) {
            const int i = e / PT_J, j = e % PT_J;
            const int gi = i0 + i, gj = j0 + j;
            const bool ok = gi < p.n_rows && gj < j_end;
            sK[j * PB_KP + i] = ok ? p.kfull[(size_t)gi * p.n_all + gj] : 0.0f;
            sKt[j * PB_KP + i] = ok ? kterm[(size_t)gi * p.n_all + gj] : 0.0f;
        }
        for (int e = tid; e < PT_J * (PT_C / 4); e += 128) {
            const int j = e / (PT_C / 4), c4 = (e % (PT_C / 4)) * 4;
            const int gj = j0 + j, gc = c0 + c4;
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), gv = xv;
            if (gj < j_end) {
                const size_t ox = (size_t)gj * p.ld + gc, og = (size_t)gj * p.g_ld + gc;
                if (gc + 3 < c_end &&
                    ((reinterpret_cast<uintptr_t>(p.x_all + ox) | reinterpret_cast<uintptr_t>(p.g_all + og)) & 15) == 0) {
                    xv = *reinterpret_cast<const float4*>(p.x_all + ox);
                    gv = *reinterpret_cast<const float4*>(p.g_all + og);
                } else {
                    if (gc < c_end) { xv.x = p.x_all[ox]; gv.x = p.g_all[og]; }
                    if (gc + 1 < c_end) { xv.y = p.x_all[ox + 1]; gv.y = p.g_all[og + 1]; }
                    if (gc + 2 < c_end) { xv.z = p.x_all[ox + 2]; gv.z = p.g_all[og + 2]; }
                    if (gc + 3 < c_end) { xv.w = p.x_all[ox + 3]; gv.w = p.g_all[og + 3]; }
                }
            }
            *reinterpret_cast<float4*>(&sXj[j * PT_C + c4]) = xv;
            *reinterpret_cast<float4*>(&sGj[j * PT_C + c4]) = gv;
        }
        __syncthreads();
        const int nj = min(PT_J, j_end - j0);
#pragma unroll 2
        for (int j = 0; j < nj; ++j) {
            const float4 kfa = *reinterpret_cast<const float4*>(&sK[j * PB_KP + ty * 8]);
            const float4 kfb = *reinterpret_cast<const float4*>(&sK[j * PB_KP + ty * 8 + 4]);
            const float4 kta = *reinterpret_cast<const float4*>(&sKt[j * PB_KP + ty * 8]);
            const float4 ktb = *reinterpret_cast<const float4*>(&sKt[j * PB_KP + ty * 8 + 4]);
            const ulonglong2 xj = *reinterpret_cast<const ulonglong2*>(&sXj[j * PT_C + tx * 4]);
            const ulonglong2 gj = *reinterpret_cast<const ulonglong2*>(&sGj[j * PT_C + tx * 4]);
            const float kf[8] = {kfa.x, kfa.y, kfa.z, kfa.w, kfb.x, kfb.y, kfb.z, kfb.w};
            const float kt[8] = {kta.x, kta.y, kta.z, kta.w, ktb.x, ktb.y, ktb.z, ktb.w};
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const f32x2 kfa2 = pack2(kf[a], kf[a]), kta2 = pack2(kt[a], kt[a]);
                drive[a][0] = fma2(kfa2, gj.x, drive[a][0]);
                drive[a][1] = fma2(kfa2, gj.y, drive[a][1]);
                rep[a][0] = fma2(kta2, sub2(xj.x, xi[a][0]), rep[a][0]);
                rep[a][1] = fma2(kta2, sub2(xj.y, xi[a][1]), rep[a][1]);
            }
        }
        __syncthreads();
    }
    const float c2 = -2.0f / h;
    float* part = p.phi_part + (size_t)blockIdx.z * p.n_rows * D;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int gi = i0 + ty * 8 + a, gc = c0 + tx * 4;
        if (gi >= p.n_rows) continue;
        float* o = part + (size_t)gi * D + gc;
        const float v0 = fmaf(c2, lo2(rep[a][0]), lo2(drive[a][0])), v1 = fmaf(c2, hi2(rep[a][0]), hi2(drive[a][0]));
        const float v2 = fmaf(c2, lo2(rep[a][1]), lo2(drive[a][1])), v3 = fmaf(c2, hi2(rep[a][1]), hi2(drive[a][1]));
        if (gc + 3 < c_end && ((((size_t)blockIdx.z * p.n_rows + gi) * D + gc) & 3) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(v0, v1, v2, v3);
        } else {
            if (gc < c_end) o[0] = v0;
            if (gc + 1 < c_end) o[1] = v1;
            if (gc + 2 < c_end) o[2] = v2;
            if (gc + 3 < c_end) o[3] = v3;
        }
    }
}

// ---- pass 4: per particle, sum the j slices in fixed order -> phi, optimizer step, and -- because the whole new
// latent row is in shared memory at that point -- the NEXT step's raw scores U V^T (the edge-probability pass,
// dibs.py:179-181; the next step's sub-keys and loop state come from k_assemble_grad).
struct UpdateParams {
    const float* phi_part; int n_jsplit;      // [n_jsplit][n_rows][D]
    int n_rows, dz, dth, n_all;
    const float* x_cur; int ld;               // this rank's rows of the current packed buffer
    float* x_next; int next_ld;               // updated rows (null: phi only)
    float* v; int v_ld;
    float* phi_out; int phi_ld;
    int optimizer; float stepsize;
    // next step's raw scores U V^T (null: skip)
    int d, k;
    float* scores;
    // peer-memory exchange fused into the kernel: the updated row also goes into every peer's particle buffer
    int row0; PeerPush push;
};

__global__ void __launch_bounds__(256) k_opt_update(UpdateParams p) {
    extern __shared__ __align__(16) float smem[];
    const int m = blockIdx.x, tid = threadIdx.x;
    const int D = p.dz + p.dth, d = p.d, k = p.k;
    float* sU = smem; float* sV = smem + k * d;       // new Z, de-interleaved and transposed: sU[kk][i], sV[kk][j]
    const float inv_m = 1.0f / (float)p.n_all;
    const size_t plane = (size_t)p.n_rows * D;
    const float* part = p.phi_part + (size_t)m * D;
    // four elements per thread per pass and the slice loop unrolled: every global read of a pass is in flight
    // before the first one is consumed (the kernel is a chain of L2 latencies otherwise)
    constexpr int UNR = 4;
    for (int e0 = tid; e0 < D; e0 += UNR * blockDim.x) {
        float sum[UNR], xv[UNR], vv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int e = e0 + u * blockDim.x;
            sum[u] = 0.0f;
            xv[u] = e < D ? p.x_cur[(size_t)m * p.ld + e] : 0.0f;
            vv[u] = (e < D && p.x_next && p.optimizer == 1) ? p.v[(size_t)m * p.v_ld + e] : 0.0f;
        }
#pragma unroll 4
        for (int s = 0; s < p.n_jsplit; ++s) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e < D) sum[u] += part[(size_t)s * plane + e];
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e >= D) continue;
            // -(weighted_gradient_ascent + repulsion).mean(axis=0)   (svgd.py:212-216)
            const float phi = -sum[u] * inv_m;
            if (p.phi_out) p.phi_out[(size_t)m * p.phi_ld + e] = phi;
            float x = xv[u];
            if (p.x_next) {
                if (p.optimizer == 1) {
                    // rmsprop(step, gamma=0.9, eps=1e-8): v = v*gamma + g^2*(1-gamma); x -= step*g/sqrt(v+eps)
                    const float v = __fadd_rn(__fmul_rn(vv[u], 0.9f), __fmul_rn(__fmul_rn(phi, phi), 0.1f));
                    p.v[(size_t)m * p.v_ld + e] = v;
                    x = __fsub_rn(x, __fdiv_rn(__fmul_rn(p.stepsize, phi), __fsqrt_rn(__fadd_rn(v, 1e-8f))));
                } else {
                    x = __fsub_rn(x, __fmul_rn(p.stepsize, phi));   // sgd: x - step * g
                }
                p.x_next[(size_t)m * p.next_ld + e] = x;
                if (p.push.world) peer_store(p.push, (size_t)(p.row0 + m) * p.next_ld + e, x);
            }
            if (p.scores && e < p.dz) {
                const int i = e / (2 * k), r = e - i * 2 * k, kk = r >> 1;
                ((r & 1) ? sV : sU)[kk * d + i] = x;
            }
        }
    }
    if (p.push.world) peer_signal(p.push, gridDim.x);
    if (!p.scores) return;
    __syncthreads();
    float* out = p.scores + (size_t)m * d * d;
    for (int e = tid; e < d * d; e += blockDim.x) {
        const int i = e / d, j = e - i * d;
        float acc = 0.0f;
        for (int kk = 0; kk < k; ++kk) acc = fmaf(sU[kk * d + i], sV[kk * d + j], acc);
        out[e] = acc;
    }
}

}  // namespace dibs
