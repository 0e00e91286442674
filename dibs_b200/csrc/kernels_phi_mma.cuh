// phi on the 5th-generation tensor cores: the two (M_loc x M)(M x D) contractions of the SVGD transform as
// tcgen05.mma (kind::tf32, fp32 accumulators in TMEM) with operands staged by TMA, in split precision (3 x TF32).
//
// replaces: dibs/inference/svgd.py:194-224 (_z_update / _parallel_update_z), :591-670 (joint: _z_update,
// _theta_update, _parallel_update_*) and the kernel gradients of :179-192,554-588; closed form SURVEY App. B-9:
//   phi_i = -(1/M) [ (K G)_i - (2/h) ( (K* X)_i - rowsum_i(K*) x_i ) ]
// with K = K_z + K_theta (drive term) and K* = K_z for the Z block, K_theta for the Theta block (repulsion).
//
// Why split precision: TF32 keeps 11 significand bits; parity with the fp32 reference (1e-5) needs ~fp32 products.
// Every operand x is split exactly into hi = x rounded to the nearest TF32 value and lo = x - hi (at most 12
// significant bits); a b ~= hi_a hi_b + hi_a lo_b + lo_a hi_b, error ~2^-23 |a b|, all three products accumulated in
// fp32 in TMEM.  The repulsion term is evaluated in GEMM form (K* X minus the row sum times
// x_i); its cancellation only bites when every particle a row has weight on sits within ~1e-4 of it, where the term
// itself vanishes against the drive term (checked against the difference-form oracle in tests/test_gpu_parity.py).
//
// Work decomposition: CTA = (128-row tile of the rank's rows, 64-column tile of [Z | Theta], slice of the j axis);
// the j slices are those of the SIMT kernel (fixed length on the GLOBAL particle index), partial tiles go to the same
// phi_part planes and the same last-arriver epilogue sums them in fixed order and applies the optimizer step.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2-5 = operand
// splitters during the main loop (hi / lo parts of the G and X tiles; thread = row of the K* tile for rowsum(K*) in
// fp32; the K planes arrive already split from k_pair_finish) and the epilogue afterwards (thread = TMEM lane = row).
//
// Shared-memory operand layouts are the canonical 128-byte-swizzled ones a plain 2-D TMA box produces:
//   A = K tiles, K-major:   [128 rows][32 j] fp32, rows of 128 B, 16-byte chunks XOR-swizzled by (row & 7);
//                           UMMA descriptor: SWIZZLE_128B, SBO = 1024 B (8 rows), K step of 8 = +32 B.
//   B = G / X tiles, MN-major: two panels of [32 j][32 columns] fp32, rows of 128 B, 32-byte chunks XOR-swizzled by
//                           (j & 3) (TMA mode SWIZZLE_128B_ATOM_32B): for an MN-major 32-bit operand the tensor core
//                           accepts only this layout (UMMA SWIZZLE_128B_BASE32B; with plain SWIZZLE_128B the MMAs
//                           silently contribute zeros -- found the hard way); LBO = 4096 B (next 32-column panel),
//                           SBO = 512 B (next 4 j), K step of 8 = +1024 B.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels_pair.cuh"

namespace dibs {

constexpr int MM_ROWS = 128;      // UMMA M
constexpr int MM_COLS = 64;       // UMMA N (= PT_C: column tiles never straddle the Z | Theta boundary)
constexpr int MM_KS = 32;         // j per pipeline stage (one 128-byte swizzle row of A)
constexpr int MM_STAGES = 2;
constexpr int MM_THREADS = 192;
constexpr int MM_A_BYTES = MM_ROWS * MM_KS * 4;       // 16 KB
constexpr int MM_B_BYTES = MM_KS * MM_COLS * 4;       //  8 KB (two panels of 4 KB)
// per stage: A1 hi/lo, A2 hi/lo, B1 hi/lo, B2 hi/lo
constexpr int MM_STAGE_BYTES = 4 * MM_A_BYTES + 4 * MM_B_BYTES;      // 96 KB
constexpr int MM_SMEM_BYTES = MM_STAGES * MM_STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
constexpr int MM_TMEM_COLS = 128;                     // D1 (drive) 64 columns + D2 (K* X) 64 columns

struct PhiMmaMaps {
    CUtensorMap a[6];                 // K hi, K lo, K_z hi, K_z lo, K_theta hi, K_theta lo  [n_rows][n_all] (PairParams::k_split)
    CUtensorMap b_x, b_g;             // particles / gradients [n_all][ld]
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded: a pipeline that never completes (bad descriptor, faulted TMA) must abort the kernel, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    unsigned spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && (++spins & 0xFFFFu) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] B[smem], one 128 x 64 x 8 TF32 MMA
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units, version 1
// (Blackwell), layout type 2 = SWIZZLE_128B (16-byte chunks XORed with the row, K-major A), 1 = SWIZZLE_128B_BASE32B
// (32-byte chunks XORed with row & 3 -- the only layout the tensor core accepts for an MN-major 32-bit operand)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;       // version_
    d |= (uint64_t)layout_type << 61;
    return d;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return umma_desc(smem_addr, lbo_bytes, sbo_bytes, 2u);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, TF32 x TF32, A K-major, B MN-major, N, M
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// x -> (hi, lo): hi = x rounded to the nearest TF32 value (13 low mantissa bits zero: the tensor core reads it exactly
// whatever its own conversion does), lo = x - hi, exact in fp32 and at most 12 significant bits.
// Measured against fp64 on the phi contraction: 3 products ~2e-6 of the largest entry; hi x hi alone 2e-3.
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
    hi.x = tf32_rn(v.x); lo.x = v.x - hi.x;
    hi.y = tf32_rn(v.y); lo.y = v.y - hi.y;
    hi.z = tf32_rn(v.z); lo.z = v.z - hi.z;
    hi.w = tf32_rn(v.w); lo.w = v.w - hi.w;
}

#ifdef DIBS_PHI_TRACE
// pipeline trace of a debug build (tools/phi_trace.py): 64 clock samples per CTA, never compiled into the product library
__device__ unsigned long long g_phi_trace[1024 * 64];
#define PHI_TR(slot) do { if (tr) tr[slot] = (unsigned long long)clock64(); } while (0)
#else
#define PHI_TR(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(MM_THREADS, 1)
k_phi_mma(const __grid_constant__ PairParams p, const __grid_constant__ PhiMmaMaps maps) {
    extern __shared__ uint8_t mm_smem_raw[];
    // 1024-byte alignment: the swizzle pattern is a function of the absolute shared-memory address
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mm_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MM_STAGES * MM_STAGE_BYTES);
    uint64_t* bar_full = bars;                       // [MM_STAGES] TMA bytes have landed
    uint64_t* bar_split = bars + MM_STAGES;          // [MM_STAGES] hi / lo tiles are ready for the tensor core
    uint64_t* bar_empty = bars + 2 * MM_STAGES;      // [MM_STAGES] the MMAs reading the stage have completed
    uint64_t* bar_accum = bars + 3 * MM_STAGES;      // all MMAs of the tile have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MM_STAGES + 1);
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef DIBS_PHI_TRACE
    const int tr_cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    unsigned long long* tr = tr_cta < 1024 ? g_phi_trace + tr_cta * 64 : nullptr;
    if (tr && tid == 0) {
        unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        tr[0] = global_timer_ns(); tr[1] = sm; tr[50] = (unsigned long long)clock64();
    }
#endif
    const int D = p.dz + p.dth;
    const int nzt = (p.dz + MM_COLS - 1) / MM_COLS;
    const bool z_block = (int)blockIdx.x < nzt;
    const int c0 = z_block ? blockIdx.x * MM_COLS : p.dz + ((int)blockIdx.x - nzt) * MM_COLS;
    const int c_end = z_block ? p.dz : D;
    const int i0 = blockIdx.y * MM_ROWS;
    const int j_begin = blockIdx.z * p.j_len;
    const int j_end = min(p.n_all, j_begin + p.j_len);
    const int n_it = (j_end - j_begin + MM_KS - 1) / MM_KS;
    const bool same_a = (p.dth == 0);                // marginal: K* == K, one A tile feeds both products
    const CUtensorMap* map_a2 = z_block ? &maps.a[2] : &maps.a[4];        // K* hi (lo = the next map)
    const float h = z_block ? p.h_z : p.h_t;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[0])) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map_a2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b_g)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b_x)) : "memory");
        for (int s = 0; s < MM_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_split + s, 4); mbar_init(bar_empty + s, 1); }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(MM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    peer_wait(p.wait_g);                             // (contains a __syncthreads) gradients of every rank have landed
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto stage_ptr = [&](int s, int which) -> uint8_t* {       // which: 0 A1h 1 A1l 2 A2h 3 A2l 4 B1h 5 B1l 6 B2h 7 B2l
        uint8_t* b = smem + (size_t)s * MM_STAGE_BYTES;
        return which < 4 ? b + which * MM_A_BYTES : b + 4 * MM_A_BYTES + (which - 4) * MM_B_BYTES;
    };

    float rs = 0.0f;                                 // rowsum of K* over the slice (splitter threads: one row each)
    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t bytes = (same_a ? 2 : 4) * MM_A_BYTES + 2 * MM_B_BYTES;
            for (int it = 0; it < n_it; ++it) {
                const int s = it % MM_STAGES;
                if (it >= MM_STAGES) mbar_wait(bar_empty + s, ((it / MM_STAGES) - 1) & 1);
                const int j0 = j_begin + it * MM_KS;
                if (it < 8) PHI_TR(8 + it);
                mbar_expect_tx(bar_full + s, bytes);
                tma_load_2d(stage_ptr(s, 0), &maps.a[0], bar_full + s, j0, i0);
                tma_load_2d(stage_ptr(s, 1), &maps.a[1], bar_full + s, j0, i0);
                if (!same_a) {
                    tma_load_2d(stage_ptr(s, 2), map_a2, bar_full + s, j0, i0);
                    tma_load_2d(stage_ptr(s, 3), map_a2 + 1, bar_full + s, j0, i0);
                }
                tma_load_2d(stage_ptr(s, 4), &maps.b_g, bar_full + s, c0, j0);
                tma_load_2d(stage_ptr(s, 4) + MM_B_BYTES / 2, &maps.b_g, bar_full + s, c0 + 32, j0);
                tma_load_2d(stage_ptr(s, 6), &maps.b_x, bar_full + s, c0, j0);
                tma_load_2d(stage_ptr(s, 6) + MM_B_BYTES / 2, &maps.b_x, bar_full + s, c0 + 32, j0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(MM_ROWS, MM_COLS);
            const uint32_t d1 = tmem_base, d2 = tmem_base + MM_COLS;
            const uint64_t da_base = umma_desc_sw128(smem_u32(stage_ptr(0, 0)), 16, 1024);
            const uint64_t db_base = umma_desc(smem_u32(stage_ptr(0, 4)), MM_B_BYTES / 2, 512, 1u);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % MM_STAGES;
                mbar_wait(bar_split + s, (it / MM_STAGES) & 1);
                if (it < 8) PHI_TR(32 + it);
                tc_fence_after();
                // descriptors of the stage's first K step (one thread issues everything: keep its instruction count low);
                    // a K step adds a constant to the start-address field (16-byte units, no carry out of the 14 bits:
                    // the stage buffers are far below 256 KB)
                const uint64_t sa = (uint64_t)(((uint32_t)s * MM_STAGE_BYTES) >> 4);
                const uint64_t ea1h = da_base + sa, ea1l = ea1h + (MM_A_BYTES >> 4);
                const uint64_t ea2h = same_a ? ea1h : ea1h + 2 * (MM_A_BYTES >> 4), ea2l = same_a ? ea1l : ea1h + 3 * (MM_A_BYTES >> 4);
                const uint64_t eb1h = db_base + sa, eb1l = eb1h + (MM_B_BYTES >> 4);
                const uint64_t eb2h = eb1h + 2 * (MM_B_BYTES >> 4), eb2l = eb1h + 3 * (MM_B_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < MM_KS / 8; ++ks) {
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    // A: K-major, +32 B per K step inside the 128-byte swizzle row; SBO = 8 rows
                    const uint64_t da1h = ea1h + 2 * ks, da1l = ea1l + 2 * ks, da2h = ea2h + 2 * ks, da2l = ea2l + 2 * ks;
                    // B: MN-major (SWIZZLE_128B_BASE32B atoms of 4 j x 128 B), +1024 B per K step (8 rows of 128 B);
                    // LBO = next 32-column panel, SBO = next 4 j
                    const uint64_t db1h = eb1h + 64 * ks, db1l = eb1l + 64 * ks, db2h = eb2h + 64 * ks, db2l = eb2l + 64 * ks;
                    tc_mma_tf32(d1, da1h, db1h, idesc, acc);      // K G
                    tc_mma_tf32(d1, da1h, db1l, idesc, 1u);
                    tc_mma_tf32(d1, da1l, db1h, idesc, 1u);
                    tc_mma_tf32(d2, da2h, db2h, idesc, acc);      // K* X
                    tc_mma_tf32(d2, da2h, db2l, idesc, 1u);
                    tc_mma_tf32(d2, da2l, db2h, idesc, 1u);
                }
                tc_commit(bar_empty + s);             // the stage may be refilled once these MMAs have read it
                if (it < 8) PHI_TR(40 + it);
            }
            tc_commit(bar_accum);                     // accumulators complete
        }
    } else {
        // ===== operand splitters (128 threads): hi / lo tiles in place, rowsum(K*) =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may read: rows 32 q .. 32 q + 31
        const int row = 32 * q + lane;                // thread = row of the A tiles
        const int st = tid - 64;                      // 0..127 flat index for the B tiles
        for (int it = 0; it < n_it; ++it) {
            const int s = it % MM_STAGES;
            mbar_wait(bar_full + s, (it / MM_STAGES) & 1);
            if (tid == 64 && it < 8) PHI_TR(16 + it);
            // K* tile (already split by the kernel that produced K): this thread's row of hi + lo, 8 chunks of 16 B each
            // -- logical chunk c sits at position c ^ (row & 7) -- summed in j order into the row sum
            {
                const float4* hi = reinterpret_cast<const float4*>(stage_ptr(s, same_a ? 0 : 2)) + row * 8;
                const float4* lo = reinterpret_cast<const float4*>(stage_ptr(s, same_a ? 1 : 3)) + row * 8;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int pos = c ^ (row & 7);
                    const float4 vh = hi[pos], vl = lo[pos];
                    rs += ((vh.x + vl.x) + (vh.y + vl.y)) + ((vh.z + vl.z) + (vh.w + vl.w));
                }
            }
            // B tiles: 2 tiles x 512 chunks, elementwise (the swizzle is the same permutation in hi and lo)
            for (int b = 0; b < 2; ++b) {
                float4* hi = reinterpret_cast<float4*>(stage_ptr(s, 4 + 2 * b));
                float4* lo = reinterpret_cast<float4*>(stage_ptr(s, 5 + 2 * b));
#pragma unroll
                for (int c = 0; c < MM_B_BYTES / 16 / 128; ++c) {
                    const float4 v = hi[st + 128 * c];
                    float4 h4, l4;
                    split_tf32(v, h4, l4);
                    hi[st + 128 * c] = h4; lo[st + 128 * c] = l4;
                }
            }
            fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_split + s);
            if (tid == 64 && it < 8) PHI_TR(24 + it);
        }
        // ===== epilogue: TMEM -> registers -> partial plane of this j slice =====
        // the tile's own rows x_i are fetched while the last MMAs drain (registers are plentiful at one CTA per SM)
        float4 xpre[MM_ROWS * (MM_COLS / 4) / 128];
#pragma unroll
        for (int w = 0; w < MM_ROWS * (MM_COLS / 4) / 128; ++w) {
            const int idx = st + 128 * w;
            const int r = idx >> 4, c4 = (idx & 15) * 4;
            const int gi = i0 + r, gc = c0 + c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gi < p.n_rows) {
                const float* xr = p.x_all + (size_t)(p.row0 + gi) * p.ld + gc;
                if (gc + 3 < c_end && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0)) v = *reinterpret_cast<const float4*>(xr);
                else {
                    if (gc < c_end) v.x = xr[0];
                    if (gc + 1 < c_end) v.y = xr[1];
                    if (gc + 2 < c_end) v.z = xr[2];
                    if (gc + 3 < c_end) v.w = xr[3];
                }
            }
            xpre[w] = v;
        }
        mbar_wait(bar_accum, 0);
        if (tid == 64) PHI_TR(48);
        tc_fence_after();
        const float c2 = -2.0f / h;
        const size_t plane = (size_t)p.n_rows * D;
        float* part = p.phi_part + (size_t)blockIdx.z * plane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
        // Global traffic of the epilogue goes through two shared-memory tiles (the operand stages are free now: every
        // TMA load and every MMA has completed) so that rows are read and written 256 contiguous bytes at a time --
        // one row per thread straight to global memory costs 32 sectors per instruction (measured: ~20 us per CTA)
        constexpr int EP_LD = MM_COLS + 1;                        // odd stride: lane = row is conflict-free
        float* sX = reinterpret_cast<float*>(smem);               // [128][65] the tile's own rows x_i
        float* sO = sX + MM_ROWS * EP_LD;                         // [128][65] partial sums of this j slice
#pragma unroll
        for (int w = 0; w < MM_ROWS * (MM_COLS / 4) / 128; ++w) {
            const int idx = st + 128 * w;
            float* d = sX + (idx >> 4) * EP_LD + (idx & 15) * 4;
            d[0] = xpre[w].x; d[1] = xpre[w].y; d[2] = xpre[w].z; d[3] = xpre[w].w;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");           // the 128 epilogue threads
#pragma unroll 1
        for (int cb = 0; cb < MM_COLS; cb += 16) {
            float dr[16], kx[16];
            tc_ld16(lane_addr + cb, dr);
            tc_ld16(lane_addr + MM_COLS + cb, kx);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                // weighted_gradient_ascent + repulsion of this j slice (svgd.py:212-216), repulsion in GEMM form
                const float repulsion = kx[u] - rs * sX[row * EP_LD + cb + u];
                sO[row * EP_LD + cb + u] = fmaf(c2, repulsion, dr[u]);
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int idx = st; idx < MM_ROWS * (MM_COLS / 4); idx += 128) {
            const int r = idx >> 4, c4 = (idx & 15) * 4;
            const int gi = i0 + r, gc = c0 + c4;
            if (gi >= p.n_rows || gc >= c_end) continue;
            const float* sv = sO + r * EP_LD + c4;
            float* o = part + (size_t)gi * D + gc;
            if (gc + 3 < c_end && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) *reinterpret_cast<float4*>(o) = make_float4(sv[0], sv[1], sv[2], sv[3]);
            else {
                if (gc < c_end) o[0] = sv[0];
                if (gc + 1 < c_end) o[1] = sv[1];
                if (gc + 2 < c_end) o[2] = sv[2];
                if (gc + 3 < c_end) o[3] = sv[3];
            }
        }
        tc_fence_before();
        if (tid == 64) PHI_TR(49);
    }
    // ---- the tile's last j-slice CTA finishes: fixed-order sum of the slices -> phi, optimizer step, peer push
    __threadfence();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(MM_TMEM_COLS) : "memory");
    }
    if (tid == 0) {
        uint32_t* cnt = p.phi_cnt + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        const unsigned old = atomicAdd(cnt, 1u);
        s_last = (old == (unsigned)p.n_jsplit - 1u) ? 1 : 0;
        if (s_last) *cnt = 0u;
    }
    __syncthreads();
#ifdef DIBS_PHI_TRACE
    if (tr && tid == 0) { tr[51] = (unsigned long long)clock64(); tr[2] = global_timer_ns(); tr[3] = s_last; }
#endif
    if (!s_last) return;
    __threadfence();
    phi_finish_tile(p, i0, MM_ROWS, c0, c_end, tid, MM_THREADS);
    if (p.push_x.world) peer_signal(p.push_x, gridDim.x * gridDim.y);
#ifdef DIBS_PHI_TRACE
    if (tr && tid == 0) { tr[52] = (unsigned long long)clock64(); tr[4] = global_timer_ns(); }
#endif
}

}  // namespace dibs
