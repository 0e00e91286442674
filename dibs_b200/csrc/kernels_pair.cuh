// All-pairs part of the SVGD step: squared-exponential kernel matrix on (Z, Theta), the transform phi and
// the optimizer update, on a rank's slab of rows against all (gathered) particles.
//
// replaces: dibs/kernel.py:20-30,52-71; dibs/inference/svgd.py:165-224 (marginal), :537-670 (joint),
// :265,718-719 + jax.example_libraries.optimizers.{sgd,rmsprop}.
// Closed form: grad_{x_j} k(x_j, x_i) = -(2/h)(x_j - x_i) k_ji, kept in difference form
// like the reference's autodiff so that near-neighbours do not cancel.
//
// Three kernels.  The distance pass splits the feature axis over CTAs and a second, fully parallel kernel sums the
// split planes and applies exp (both sit on the side branch of the step graph, under the gradient phase -- measured:
// finishing a tile inside its last-arriving split CTA is 4x slower at 256 particles, 16 busy SMs instead of 148).
// The phi pass splits the j axis over CTAs, counts arrivals per OUTPUT TILE, and the CTA that completes a tile's count
// sums the slices in FIXED order and applies the epilogue (mean, optimizer step, peer push): one kernel less on the
// critical path.  Summation orders are the plane index, never the arrival order, and the slicing is a function of
// (M, D) only -- results are bit-identical for any number of GPUs and from run to run.
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
    uint32_t* phi_cnt;                   // [row tiles][column tiles] arrival counters of the phi pass
    PeerWait wait_x, wait_g;             // peer-memory exchange: flags to wait on before reading x_all / g_all
    float* kz; float* kt; float* kfull;  // [n_rows][n_all]
    // tensor-core phi: the same three planes split into TF32 hi / lo parts by the kernel that produces them, so the
    // phi kernel's TMA loads are MMA operands as they land ([6][n_rows][n_all]: K hi, K lo, K_z hi, K_z lo, K_t hi,
    // K_t lo; null: not needed)
    float* k_split;
    float h_z, h_t, scale_z, scale_t;
    // epilogue of the phi pass (optimizer step; svgd.py:265,718-719)
    float* x_next; int next_ld;          // updated rows of this rank (null: phi only)
    float* v; int v_ld;                  // RMSprop second moments [n_rows][D]
    float* phi_out; int phi_ld;          // hooks: phi itself [n_rows][D] (null: skip)
    int optimizer; float stepsize;
    PeerPush push_x;                     // fused exchange: updated rows also go into every peer's next particle buffer
};

// ---- pass 1: squared distances -> K.  Tile 64 x 64 outputs, 256 threads x (4 x 4) with INTERLEAVED ownership
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

// ---- pass 1b: sum the feature splits in fixed order, apply the SE kernels (kernel.py:30,66-71)
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
        const float kf = kz + kt;
        p.kfull[e] = kf;
        if (p.k_split) {
            // x = hi + lo exactly, hi = x rounded to the nearest TF32 value (kernels_phi_mma.cuh)
            const float fh = __uint_as_float((__float_as_uint(kf) + 0x1000u) & 0xFFFFE000u);
            const float zh = __uint_as_float((__float_as_uint(kz) + 0x1000u) & 0xFFFFE000u);
            const float th = __uint_as_float((__float_as_uint(kt) + 0x1000u) & 0xFFFFE000u);
            p.k_split[e] = fh; p.k_split[plane + e] = kf - fh;
            p.k_split[2 * plane + e] = zh; p.k_split[3 * plane + e] = kz - zh;
            if (p.kt) { p.k_split[4 * plane + e] = th; p.k_split[5 * plane + e] = kt - th; }
        }
    }
}

// ---- pass 2: phi_i = -(1/M) sum_j [ K_ij g_j - (2/h) Kterm_ij (x_j - x_i) ], then the optimizer step.
// Tile: (8 RPT) rows x 64 feature columns per CTA, 128 threads x (RPT rows x 4 columns).  RPT = 8 (64-row tiles): per
// particle j a thread issues six 128-bit shared-memory loads (K and Kterm for its 8 rows, x_j and g_j for its 4
// columns) for 48 packed FP instructions -- 12 % faster than RPT = 4 at 1024 rows per rank; RPT = 4 (32-row tiles)
// where a rank owns few rows, so that twice as many CTAs share the epilogue.  The row weight enters the FFMA2 as a
// broadcast scalar.  Every output element is the same fixed-order sum whatever the tile height.
// blockIdx.z = slice of the j axis (fixed length j_len on the GLOBAL particle index).  The last slice CTA of a tile to
// arrive sums the slices in order -> phi, applies RMSprop / SGD to its 64 x 64 block of the rank's rows and -- on
// several GPUs -- stores the updated block into every peer's next particle buffer (fused exchange).
constexpr int PT_C = 64, PT_J = 32;

// Epilogue of the phi pass for one output tile -- rows [i0, i0 + tile_rows) of the rank's slab x the (up to) 64 columns
// [c0, c_end) -- run by ALL threads of the tile's last-arriving j-slice CTA (SIMT and tensor-core kernels alike):
// fixed-order sum of the slice planes -> phi = -(sum)/M (svgd.py:212-216), RMSprop / SGD step (svgd.py:265,718-719;
// jax.example_libraries.optimizers), updated rows to this rank's next buffer and -- on several GPUs -- to every
// peer's.  16 consecutive threads cover one row's 64 columns: 128-bit accesses, 256 contiguous bytes per row.
// optimizer step of one element (IEEE-ordered like jax.example_libraries.optimizers; svgd.py:265,718-719)
__device__ __forceinline__ void phi_opt_step(int optimizer, float stepsize, float phi, float x, float v, float& xn, float& vn) {
    if (optimizer == 1) {
        // rmsprop(step, gamma=0.9, eps=1e-8): v = v*gamma + g^2*(1-gamma); x -= step*g/sqrt(v+eps)
        vn = __fadd_rn(__fmul_rn(v, 0.9f), __fmul_rn(__fmul_rn(phi, phi), 0.1f));
        xn = __fsub_rn(x, __fdiv_rn(__fmul_rn(stepsize, phi), __fsqrt_rn(__fadd_rn(vn, 1e-8f))));
    } else {
        vn = 0.0f;
        xn = __fsub_rn(x, __fmul_rn(stepsize, phi));   // sgd: x - step * g
    }
}

// elements [gc, min(gc + 4, c_end)) of row gi, one at a time: tile edges and layouts whose rows are not 16-byte
// aligned (odd n_vars).  Deliberately out of line and rolled -- see phi_finish_tile on code size
static __device__ __noinline__ void phi_finish_scalar(const PairParams& p, int gi, int gc, int c_end, int D, size_t plane,
                                                      float inv_m, bool rms) {
#pragma unroll 1
    for (int c = gc; c < min(gc + 4, c_end); ++c) {
        const float* pr = p.phi_part + (size_t)gi * D + c;
        float sum = 0.0f;
#pragma unroll 1
        for (int s = 0; s < p.n_jsplit; ++s) sum += __ldcg(pr + (size_t)s * plane);
        const float phi = -sum * inv_m;                      // -(weighted_gradient_ascent + repulsion).mean(axis=0)
        if (p.phi_out) p.phi_out[(size_t)gi * p.phi_ld + c] = phi;
        if (!p.x_next) continue;
        float* vp = rms ? p.v + (size_t)gi * p.v_ld + c : nullptr;
        float xn, vn;
        phi_opt_step(p.optimizer, p.stepsize, phi, p.x_all[(size_t)(p.row0 + gi) * p.ld + c], rms ? *vp : 0.0f, xn, vn);
        if (rms) *vp = vn;
        p.x_next[(size_t)gi * p.next_ld + c] = xn;
        if (p.push_x.world) peer_store(p.push_x, (size_t)(p.row0 + gi) * p.next_ld + c, xn);
    }
}

struct PhiFinishLoads { float4 q[4], x, v; };

// The tile's last-arriving CTA runs this ONCE: it is cold code, and what it costs is instruction fetch (the SM's
// 32 KB instruction cache holds the main loop; an earlier, fully unrolled version of this pass was 65 KB of SASS and
// took 25 us per 128 x 64 tile, hardly any of it memory latency).  Hence: ONE rolled loop over 4-column items,
// software-pipelined by hand -- the (up to) six 128-bit reads of the next item are issued before the current one is
// finished -- and everything irregular pushed into phi_finish_scalar.
__device__ __forceinline__ void phi_finish_tile(const PairParams& p, int i0, int tile_rows, int c0, int c_end, int tid, int nthr) {
    const int D = p.dz + p.dth;
    const size_t plane = (size_t)p.n_rows * D;
    const float inv_m = 1.0f / (float)p.n_all;
    const int limit = tile_rows * 16, n_jsplit = p.n_jsplit;
    const bool rms = p.x_next && p.optimizer == 1;
    // 128-bit path: every row of every buffer the item touches starts 16-byte aligned
    const bool aligned = (((D | p.ld | (p.x_next ? p.next_ld : 0) | (rms ? p.v_ld : 0) | c0) & 3) == 0) && ((plane & 3) == 0) &&
                         (((reinterpret_cast<uintptr_t>(p.phi_part) | reinterpret_cast<uintptr_t>(p.x_all) |
                            reinterpret_cast<uintptr_t>(p.x_next) | reinterpret_cast<uintptr_t>(p.v)) & 15) == 0) && !p.phi_out;
    auto vec_item = [&](int idx) -> bool {
        const int gi = i0 + (idx >> 4), gc = c0 + (idx & 15) * 4;
        return idx < limit && aligned && gi < p.n_rows && gc + 3 < c_end;
    };
    auto load = [&](PhiFinishLoads& L, int idx) {
        if (!vec_item(idx)) return;
        const int gi = i0 + (idx >> 4), gc = c0 + (idx & 15) * 4;
        const float* pr = p.phi_part + (size_t)gi * D + gc;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (u < n_jsplit) L.q[u] = __ldcg(reinterpret_cast<const float4*>(pr + (size_t)u * plane));
        L.x = *reinterpret_cast<const float4*>(p.x_all + (size_t)(p.row0 + gi) * p.ld + gc);
        if (rms) L.v = *reinterpret_cast<const float4*>(p.v + (size_t)gi * p.v_ld + gc);
    };
    PhiFinishLoads cur, nxt;
    load(cur, tid);
#pragma unroll 1
    for (int idx = tid; idx < limit; idx += nthr) {
        load(nxt, idx + nthr);
        const int gi = i0 + (idx >> 4), gc = c0 + (idx & 15) * 4;
        if (vec_item(idx)) {
            float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (u < n_jsplit) { sum[0] += cur.q[u].x; sum[1] += cur.q[u].y; sum[2] += cur.q[u].z; sum[3] += cur.q[u].w; }
#pragma unroll 1
            for (int s = 4; s < n_jsplit; ++s) {               // more than four slices: the rest in order
                const float4 q = __ldcg(reinterpret_cast<const float4*>(p.phi_part + (size_t)s * plane + (size_t)gi * D + gc));
                sum[0] += q.x; sum[1] += q.y; sum[2] += q.z; sum[3] += q.w;
            }
            const float xc[4] = {cur.x.x, cur.x.y, cur.x.z, cur.x.w};
            const float vo[4] = {rms ? cur.v.x : 0.0f, rms ? cur.v.y : 0.0f, rms ? cur.v.z : 0.0f, rms ? cur.v.w : 0.0f};
            float xn[4], vn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) phi_opt_step(p.optimizer, p.stepsize, -sum[u] * inv_m, xc[u], vo[u], xn[u], vn[u]);
            if (p.x_next) {
                if (rms) *reinterpret_cast<float4*>(p.v + (size_t)gi * p.v_ld + gc) = make_float4(vn[0], vn[1], vn[2], vn[3]);
                const float4 x4 = make_float4(xn[0], xn[1], xn[2], xn[3]);
                *reinterpret_cast<float4*>(p.x_next + (size_t)gi * p.next_ld + gc) = x4;
                if (p.push_x.world) {
                    const size_t off_all = (size_t)(p.row0 + gi) * p.next_ld + gc;      // in a whole particle buffer
#pragma unroll 1
                    for (int q = 0; q < p.push_x.world; ++q)
                        if (q != p.push_x.rank) *reinterpret_cast<float4*>(p.push_x.dst[q] + off_all) = x4;
                }
            }
        } else if (gi < p.n_rows && gc < c_end) {
            phi_finish_scalar(p, gi, gc, c_end, D, plane, inv_m, rms);
        }
        cur = nxt;
    }
}

template <int RPT>
__global__ void __launch_bounds__(128) k_phi(const __grid_constant__ PairParams p) {
    constexpr int PB_I = 8 * RPT;       // rows per tile
    constexpr int PB_KP = PB_I + 4;     // padded stride of the transposed K tiles [j][i]
    __shared__ __align__(16) float sK[PT_J * PB_KP];      // K_full[i][j] transposed: [j][i]
    __shared__ __align__(16) float sKt[PT_J * PB_KP];     // K term (z or theta block)
    __shared__ __align__(16) float sXj[PT_J * PT_C];
    __shared__ __align__(16) float sGj[PT_J * PT_C];
    __shared__ int s_last;
    peer_wait(p.wait_g);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // ty 0..7: rows RPT ty .. RPT ty + RPT - 1
    const int i0 = blockIdx.y * PB_I;
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

    f32x2 xi[RPT][2], drive[RPT][2], rep[RPT][2];         // [row][column pair]
#pragma unroll
    for (int a = 0; a < RPT; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int gi = i0 + ty * RPT + a, gc = c0 + tx * 4 + 2 * b;
            const float* xr = p.x_all + (size_t)(p.row0 + gi) * p.ld;
            const float v0 = (gi < p.n_rows && gc < c_end) ? xr[gc] : 0.0f;
            const float v1 = (gi < p.n_rows && gc + 1 < c_end) ? xr[gc + 1] : 0.0f;
            xi[a][b] = pack2(v0, v1);
            drive[a][b] = 0ull; rep[a][b] = 0ull;
        }

    for (int j0 = j_begin; j0 < j_end; j0 += PT_J) {
        // K tiles: coalesced along j in global memory, transposed into [j][i]
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
            float kf[RPT], kt[RPT];
#pragma unroll
            for (int q = 0; q < RPT / 4; ++q) {
                const float4 kfa = *reinterpret_cast<const float4*>(&sK[j * PB_KP + ty * RPT + 4 * q]);
                const float4 kta = *reinterpret_cast<const float4*>(&sKt[j * PB_KP + ty * RPT + 4 * q]);
                kf[4 * q] = kfa.x; kf[4 * q + 1] = kfa.y; kf[4 * q + 2] = kfa.z; kf[4 * q + 3] = kfa.w;
                kt[4 * q] = kta.x; kt[4 * q + 1] = kta.y; kt[4 * q + 2] = kta.z; kt[4 * q + 3] = kta.w;
            }
            const ulonglong2 xj = *reinterpret_cast<const ulonglong2*>(&sXj[j * PT_C + tx * 4]);
            const ulonglong2 gj = *reinterpret_cast<const ulonglong2*>(&sGj[j * PT_C + tx * 4]);
#pragma unroll
            for (int a = 0; a < RPT; ++a) {
                const f32x2 kfa2 = pack2(kf[a], kf[a]), kta2 = pack2(kt[a], kt[a]);
                drive[a][0] = fma2(kfa2, gj.x, drive[a][0]);
                drive[a][1] = fma2(kfa2, gj.y, drive[a][1]);
                rep[a][0] = fma2(kta2, sub2(xj.x, xi[a][0]), rep[a][0]);
                rep[a][1] = fma2(kta2, sub2(xj.y, xi[a][1]), rep[a][1]);
            }
        }
        __syncthreads();
    }
    // weighted_gradient_ascent + repulsion of this j slice   (svgd.py:212-216)
    const float c2 = -2.0f / h;
    const size_t plane = (size_t)p.n_rows * D;
    float* part = p.phi_part + (size_t)blockIdx.z * plane;
    const int gc = c0 + tx * 4;
#pragma unroll
    for (int a = 0; a < RPT; ++a) {
        const int gi = i0 + ty * RPT + a;
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
    // ---- the tile's last j-slice CTA: fixed-order sum of the slices -> phi, optimizer step, peer push
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        uint32_t* cnt = p.phi_cnt + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        const unsigned old = atomicAdd(cnt, 1u);
        s_last = (old == (unsigned)p.n_jsplit - 1u) ? 1 : 0;
        if (s_last) *cnt = 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    phi_finish_tile(p, i0, PB_I, c0, c_end, tid, 128);
    if (p.push_x.world) peer_signal(p.push_x, gridDim.x * gridDim.y);      // one signal per output tile
}

}  // namespace dibs
