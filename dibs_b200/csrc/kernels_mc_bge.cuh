// BGe marginal likelihood scoring of sampled hard graphs, fused with Bernoulli sampling and the
// score-function aggregation (MarginalDiBS hot path).
//
// replaces: dibs/models/linearGaussian.py:63-170 (BGe._log_marginal_likelihood_single, log_marginal_likelihood,
// interventional_log_marginal_prob) and dibs/utils/func.py:128-145 (_slogdet_jax).
//
// The reference pads the parent sub-matrix of R with an identity and runs two d x d LU factorisations per node
// per graph, recomputing R (which depends only on the data) every time.  Here R_j is precomputed once per
// dibs_set_data in fp64 and a (graph, node) task needs  logdet R[P,P]  and the Schur complement of j on P
// (logdet R[P+j,P+j] = logdet R[P,P] + log schur).  Factorisations run in fp64: conditioning
// of R (~N var(x) / small_t) would otherwise leak ~1e-3 absolute error into log-probs that feed a softmax.
//
// The S samples of one (particle, node j) draw their parent sets from the same edge probabilities, and once
// alpha(t) has saturated most probabilities the sets differ in a few "uncertain" parents only.  Per (particle,
// node, chunk of samples) the kernel therefore splits the candidates EXACTLY (from the draws, no thresholds):
//   C = parents present in every sample,  U = parents present in some but not all,  the rest never occur;
// eliminates C once from R[C+U+j] (warp-cooperative right-looking LDL^T, one lane per row, rows in registers),
// which leaves  logdet R[C,C]  and the Schur complement  W = R[U+j,U+j] - R[U+j,C] R[C,C]^-1 R[C,U+j];  then
//   det R[C+Us, C+Us] = det R[C,C] det W[Us,Us]      for every sample's uncertain subset Us,
// so each sample only factorises a (|Us|+1)-sized principal block of W: one lane per sample, Cholesky in fp64.
// At t = 0 (all probabilities 0.5) C is empty and this degenerates to one full-size factorisation per task.
//
// Work decomposition: CTA = (particle, chunk of sample slots); a slot is the sample pair (s, s + S/2) -- the two
// lanes of one threefry block in JAX's legacy layout -- or a single sample.  Stage A: warps draw (slot, node)
// columns, lane = candidate parent, ballots give the parent bit-masks.  Stage B: a warp per node does the split,
// the elimination and the per-sample blocks.  Stage C: softmax over the chunk and the weighted mean graph.
#pragma once
#include "common.cuh"
#include "kernels_mc.cuh"

namespace dibs {

__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}

// 1/x for a positive, normal x: fp32 seed + Newton steps to full fp64 accuracy
__device__ __forceinline__ double fast_rcp_pos(double x) {
    double r = (double)__frcp_rn((float)x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// index (0..63) of the r-th set bit of a 64-bit mask; r < popc(mask)
__device__ __forceinline__ int nth_set_bit64(unsigned long long mask, int r) {
    const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
    const int nlo = __popc(lo);
    if (r < nlo) return (int)__fns(lo, 0u, r + 1);
    return 32 + (int)__fns(hi, 0u, r - nlo + 1);
}

// Warp-cooperative elimination.  Rows (lane r holds row r in aA, row r + 32 in aB) are the nodes
//   [ C ascending | U ascending | j ],   n_rows = nC + nU + 1;
// eliminates the first nC columns of R[rows, rows] (right-looking LDL^T), returns sum log pivots = logdet R[C,C]
// and writes the remaining lower triangle -- the Schur complement W, (nU+1) x (nU+1) -- packed row-major
// (W[a][b], a >= b, at a(a+1)/2 + b) into shared memory.
template <int DMAX>
__device__ __forceinline__ double bge_schur(const double* __restrict__ R, int d, unsigned long long Cm, unsigned long long Um,
                                            int j, double* __restrict__ colbuf, double* __restrict__ W, int lane) {
    constexpr int NA = DMAX < 32 ? DMAX : 32;          // row length of slot A (rows 0..31)
    constexpr bool HASB = DMAX > 32;
    constexpr int NB = HASB ? DMAX : 2;                // row length of slot B (rows 32..63)
    constexpr int LINE = HASB ? 64 : 32;               // doubles per published column
    const int nC = __popcll(Cm), nU = __popcll(Um);
    const int n = nC + nU + 1;
    const bool useB = HASB && n > 32;

    int idxA = j, idxB = j;
    if (lane < nC) idxA = nth_set_bit64(Cm, lane);
    else if (lane < nC + nU) idxA = nth_set_bit64(Um, lane - nC);
    if (HASB) {
        if (lane + 32 < nC) idxB = nth_set_bit64(Cm, lane + 32);
        else if (lane + 32 < nC + nU) idxB = nth_set_bit64(Um, lane + 32 - nC);
    }
    const double* rowA = R + (size_t)idxA * d;
    const double* rowB = R + (size_t)idxB * d;

    double aA[NA], aB[NB];
#pragma unroll
    for (int c = 0; c < NA; ++c) aA[c] = 0.0;
#pragma unroll
    for (int c = 0; c < NB; ++c) aB[c] = 0.0;
    // gather a[r][c] = R[idx_r][idx_c]  (symmetric; entries right of the diagonal are never used)
#pragma unroll
    for (int c = 0; c < NA; ++c) {
        if (c < n) {
            const int ic = __shfl_sync(0xffffffffu, idxA, c);
            aA[c] = rowA[ic];
            if (HASB) { if (useB) aB[c] = rowB[ic]; }
        }
    }
    if (HASB) {
#pragma unroll
        for (int c = 32; c < NB; ++c) {
            if (c < n) {
                const int ic = __shfl_sync(0xffffffffu, idxB, c - 32);
                aB[c] = rowB[ic];
            }
        }
    }

    double mypivA = 1.0, mypivB = 1.0;                  // pivot of the lane's own row (log taken once, at the end)
#pragma unroll
    for (int c = 0; c < NA; ++c) {
        if (c < nC) {                                   // warp-uniform
            const double piv = shfl_f64(aA[c], c);
            const double inv = fast_rcp_pos(piv);
            if (lane == c) mypivA = piv;
            double* line = colbuf + (c & 1) * LINE;
            const double wA = aA[c];
            line[lane] = wA;
            double wB = 0.0;
            if (HASB) { wB = aB[c]; if (useB) line[32 + lane] = wB; }
            __syncwarp();
            const double lA = -wA * inv, lB = -wB * inv;
#pragma unroll
            for (int c2 = (c + 1) & ~1; c2 < NA; c2 += 2) {
                if (c2 < n) {
                    const double2 cc = *reinterpret_cast<const double2*>(line + c2);
                    aA[c2] = fma(lA, cc.x, aA[c2]);
                    if (c2 + 1 < NA) aA[c2 + 1] = fma(lA, cc.y, aA[c2 + 1]);
                }
            }
            if (HASB) {
                if (useB) {
#pragma unroll
                    for (int c2 = (c + 1) & ~1; c2 < NB; c2 += 2) {
                        if (c2 < n) {
                            const double2 cc = *reinterpret_cast<const double2*>(line + c2);
                            aB[c2] = fma(lB, cc.x, aB[c2]);
                            if (c2 + 1 < NB) aB[c2 + 1] = fma(lB, cc.y, aB[c2 + 1]);
                        }
                    }
                }
            }
        }
    }
    if (HASB) {
        if (nC > 32) {
#pragma unroll
            for (int c = 32; c < NB; ++c) {
                if (c < nC) {
                    const double piv = shfl_f64(aB[c], c - 32);
                    const double inv = fast_rcp_pos(piv);
                    if (lane == c - 32) mypivB = piv;
                    double* line = colbuf + (c & 1) * LINE;
                    const double wB = aB[c];
                    line[32 + lane] = wB;
                    __syncwarp();
                    const double lB = -wB * inv;
#pragma unroll
                    for (int c2 = (c + 1) & ~1; c2 < NB; c2 += 2) {
                        if (c2 < n) {
                            const double2 cc = *reinterpret_cast<const double2*>(line + c2);
                            aB[c2] = fma(lB, cc.x, aB[c2]);
                            if (c2 + 1 < NB) aB[c2 + 1] = fma(lB, cc.y, aB[c2 + 1]);
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    // write the trailing lower triangle: W[a][b] = a[r][c], a = r - nC, b = c - nC
    {
        const int a = lane - nC;
        if (a >= 0 && lane < n) {
            double* wrow = W + (size_t)a * (a + 1) / 2 - nC;
#pragma unroll
            for (int c = 0; c < NA; ++c)
                if (c >= nC && c <= lane) wrow[c] = aA[c];
        }
    }
    if (HASB) {
        if (useB) {
            const int a = lane + 32 - nC;
            if (a >= 0 && lane + 32 < n) {
                double* wrow = W + (size_t)a * (a + 1) / 2 - nC;
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c >= nC && c <= lane + 32) wrow[c] = aB[c];
            }
        }
    }
    __syncwarp();
    // logdet R[C,C] = sum of log pivots: one log per lane, fixed-order butterfly
    double logdet = (lane < nC) ? log(mypivA) : 0.0;
    if (HASB) { if (lane + 32 < nC) logdet += log(mypivB); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) logdet += shfl_xor_f64(logdet, o);
    return logdet;
}

template <int DMAX, int MODE>
__global__ void __launch_bounds__(256, (DMAX > 32 ? 1 : 2)) k_mc_bge(const __grid_constant__ McParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LINE = DMAX > 32 ? 64 : 32;
    constexpr int ACCN = (DMAX * DMAX + 255) / 256;                    // accumulator entries per thread
    const int d = p.d, dd = d * d;
    const int WSZ = ((d + 1) * (d + 2) / 2 + 1) & ~1;  // doubles per packed W
    const int m = blockIdx.x, c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = p.st ? p.st->t : p.t_override;
    const int S = p.n_samples;
    const bool paired = p.paired != 0;
    const int per = paired ? 2 : 1;                     // samples per slot
    const int Qh = paired ? (S >> 1) : S;               // slots per particle
    const bool r_shared = p.bge_r_stride == 0;
    const int q_begin = c * p.s_per_chunk;
    const int q_end = min(Qh, q_begin + p.s_per_chunk);
    const int n_slots = q_end - q_begin;
    const int n_smp = n_slots * per;                    // local sample ls = per * slot + h
    const int cap = p.s_per_chunk * per;                // smem capacity in samples

    // ---- shared memory carve-up (8-byte types first)
    double* sR = reinterpret_cast<double*>(smem_raw);                 // [dd] (only when R is shared by all nodes)
    double* sWall = sR + (r_shared ? ((dd + 1) & ~1) : 0);           // [8][WSZ]
    double* sCol = sWall + 8 * WSZ;                                   // [8][2][LINE]
    unsigned long long* sPar = reinterpret_cast<unsigned long long*>(sCol + 8 * 2 * LINE);   // [cap][d]
    float* sA = reinterpret_cast<float*>(sPar + (size_t)cap * d);     // [dd] edge probabilities
    float* sNode = sA + dd;                                           // [cap][d]
    float* sWgt = sNode + (size_t)cap * d;                            // [cap] sample log-probs, then softmax weights
    float* sStat = sWgt + cap;                                        // [4]

    const bool use_ext = p.g_ext != nullptr;
    if (!use_ext) stage_scores(p, m, sA, true, t);
    if (r_shared) for (int e = tid; e < dd; e += blockDim.x) sR[e] = p.bge_r[e];
    const uint2 key = use_ext ? make_uint2(0, 0) : mc_key(p, m);
    __syncthreads();

    const uint32_t n_total = (uint32_t)S * dd;
    const uint32_t half = n_total >> 1;

    // ---- stage A: parent bit-masks of every (sample, node) of the chunk
    for (int w = warp; w < n_slots * d; w += 8) {
        const int sl = w / d, j = w - sl * d;
        const int q = q_begin + sl;
        uint32_t m0[2] = {0u, 0u}, m1[2] = {0u, 0u};
#pragma unroll
        for (int hb = 0; hb < (DMAX > 32 ? 2 : 1); ++hb) {
            const int i = lane + 32 * hb;
            bool g0 = false, g1 = false;
            if (i < d && i != j) {                       // zero_diagonal: the diagonal draw is discarded
                if (use_ext) {
                    g0 = p.g_ext[(((size_t)m * S + q) * d + i) * d + j] > 0.5f;
                } else if (paired) {
                    const uint32_t e0 = ((uint32_t)q * d + i) * d + j;
                    const uint2 bits = threefry2x32(key.x, key.y, e0, e0 + half);
                    const float pe = sA[i * d + j];
                    g0 = bits_to_unit(bits.x) < pe;
                    g1 = bits_to_unit(bits.y) < pe;
                } else {
                    const uint32_t e0 = ((uint32_t)q * d + i) * d + j;
                    g0 = bits_to_unit(jax_bits(key, e0, n_total, p.partitionable)) < sA[i * d + j];
                }
            }
            m0[hb] = __ballot_sync(0xffffffffu, g0);
            m1[hb] = __ballot_sync(0xffffffffu, g1);
        }
        if (lane == 0) {
            sPar[(size_t)(per * sl) * d + j] = ((unsigned long long)m0[1] << 32) | m0[0];
            if (paired) sPar[(size_t)(per * sl + 1) * d + j] = ((unsigned long long)m1[1] << 32) | m1[0];
        }
    }
    __syncthreads();

    // ---- stage B: a warp per node
    double* W = sWall + (size_t)warp * WSZ;
    double* colbuf = sCol + (size_t)warp * 2 * LINE;
    for (int j = warp; j < d; j += 8) {
        const bool node_ok = p.bge_coef[2 * j + 1] != 0.0f;
        if (!node_ok) {                                  // no observations for this node: score 0 (linearGaussian.py:138)
            for (int ls = lane; ls < n_smp; ls += 32) sNode[(size_t)ls * d + j] = 0.0f;
            continue;
        }
        unsigned long long am = ~0ull, om = 0ull;
        for (int ls = lane; ls < n_smp; ls += 32) { const unsigned long long pm = sPar[(size_t)ls * d + j]; am &= pm; om |= pm; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            am &= __shfl_xor_sync(0xffffffffu, am, o);
            om |= __shfl_xor_sync(0xffffffffu, om, o);
        }
        const unsigned long long Cm = am, Um = om & ~am;
        const int nC = __popcll(Cm), nU = __popcll(Um);
        const double* R = r_shared ? sR : p.bge_r + (size_t)j * p.bge_r_stride;
        const double logdetC = bge_schur<DMAX>(R, d, Cm, Um, j, colbuf, W, lane);
        const double coef_j = (double)p.bge_coef[2 * j];
        const float* table = p.bge_table + (size_t)j * (d + 1);
        // per-sample principal blocks of W: lane = sample
        for (int ls0 = 0; ls0 < n_smp; ls0 += 32) {
            const int ls = ls0 + lane;
            if (ls < n_smp) {
                const unsigned long long um = sPar[(size_t)ls * d + j] & Um;
                unsigned char idx[DMAX + 1];             // positions (within [U | j]) of the rows of this sample's block
                int lu = 0;
                {
                    unsigned long long rest = um;
                    while (rest) {
                        const int b = __ffsll((long long)rest) - 1;
                        rest &= rest - 1;
                        idx[lu++] = (unsigned char)__popcll(Um & ((1ull << b) - 1ull));
                    }
                }
                idx[lu] = (unsigned char)nU;
                double L[(DMAX + 1) * (DMAX + 2) / 2];
                double logdet = logdetC, schur = 1.0;
                for (int r = 0; r <= lu; ++r) {
                    const double* Wrow = W + (size_t)idx[r] * (idx[r] + 1) / 2;
                    const int ro = r * (r + 1) / 2;
                    for (int cc = 0; cc <= r; ++cc) {
                        const int co = cc * (cc + 1) / 2;
                        double sum = Wrow[idx[cc]];
                        for (int kk = 0; kk < cc; ++kk) sum -= L[ro + kk] * L[co + kk];
                        if (cc < r) L[ro + cc] = sum * L[co + cc];         // diagonal slots hold 1/L_cc
                        else if (r < lu) { logdet += log(sum); L[ro + r] = rsqrt(sum); }
                        else schur = sum;
                    }
                }
                // 0.5 A logdet R_PP - 0.5 (A+1) logdet R_(P+j)(P+j),  A = N_j + alpha_lambd - d + l  (linearGaussian.py:109-115)
                const int l = nC + lu;
                const double a1 = coef_j + (double)l + 1.0;
                sNode[(size_t)ls * d + j] = (float)((double)table[l] - 0.5 * logdet - 0.5 * a1 * log(schur));
            }
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- stage C: sample log-probs, softmax over the chunk, weighted mean graph
    for (int ls = tid; ls < n_smp; ls += blockDim.x) {
        float lp = 0.0f;
        for (int jj = 0; jj < d; ++jj) lp += sNode[(size_t)ls * d + jj];
        if (p.lp_out) {
            const int sl = ls / per, h = ls - sl * per;
            p.lp_out[(size_t)m * S + q_begin + sl + h * Qh] = lp;
        }
        sWgt[ls] = lp;
    }
    if (MODE == MC_LP_ONLY) return;
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int ls = lane; ls < n_smp; ls += 32) mx = fmaxf(mx, sWgt[ls]);
        mx = warp_max(mx);
        float se = 0.0f, sl_ = 0.0f;
        for (int ls0 = 0; ls0 < n_smp; ls0 += 32) {     // fixed order: blocks of 32, butterfly inside
            const int ls = ls0 + lane;
            float e = 0.0f, lpv = 0.0f;
            if (ls < n_smp) { lpv = sWgt[ls]; e = expf(lpv - mx); }
            se += warp_sum(e); sl_ += warp_sum(lpv);
            __syncwarp();
            if (ls < n_smp) sWgt[ls] = e;
        }
        if (lane == 0) { sStat[0] = mx; sStat[1] = se; sStat[2] = sl_; }
    }
    __syncthreads();
    float* out = p.part_acc + ((size_t)m * p.n_chunks + c) * p.acc_size;
#pragma unroll
    for (int a = 0; a < ACCN; ++a) {
        const int e = tid + 256 * a;
        if (e < dd) {
            const int i = e / d, jj = e - i * d;
            float add = 0.0f;
            for (int ls = 0; ls < n_smp; ++ls) {
                const unsigned long long pm = sPar[(size_t)ls * d + jj];
                add += ((pm >> i) & 1ull) ? sWgt[ls] : 0.0f;
            }
            out[e] = add;                                // weighted mean graph numerator (App. B-2)
        }
    }
    if (tid == 0) {
        float* stv = p.part_stats + ((size_t)m * p.n_chunks + c) * 4;
        stv[0] = sStat[0]; stv[1] = sStat[1]; stv[2] = sStat[2]; stv[3] = 0.0f;
    }
    fuse_arrive(p.fuse, m, reinterpret_cast<float*>(smem_raw));
}

inline size_t mc_bge_smem(int d, int dmax, int cap_samples, bool r_shared) {
    const size_t dd = (size_t)d * d;
    const int line = dmax > 32 ? 64 : 32;
    const size_t wsz = (((size_t)(d + 1) * (d + 2) / 2) + 1) & ~(size_t)1;
    size_t bytes = 0;
    if (r_shared) bytes += ((dd + 1) & ~(size_t)1) * sizeof(double);
    bytes += 8 * wsz * sizeof(double);
    bytes += (size_t)8 * 2 * line * sizeof(double);
    bytes += (size_t)cap_samples * d * sizeof(unsigned long long);
    bytes += (dd + (size_t)cap_samples * d + cap_samples + 4) * sizeof(float);
    return bytes + 16;
}

}  // namespace dibs
