#include "mc_inst.cuh"
#include "kernels_mc_lin_qr.cuh"
#ifndef DIBS_DMAX
#error "compile with -DDIBS_DMAX=<n>"
#endif
namespace dibs {
int DIBS_CAT(launch_mc_lingauss_, DIBS_DMAX)(int mode, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream) {
    switch (mode) {
        case MC_THETA_HARD: return mc_launch_one(k_mc_lingauss<DIBS_DMAX, MC_THETA_HARD>, q, grid, smem, stream);
        case MC_Z_SCORE: return mc_launch_one(k_mc_lingauss<DIBS_DMAX, MC_Z_SCORE>, q, grid, smem, stream);
        case MC_Z_REPARAM: return mc_launch_one(k_mc_lingauss<DIBS_DMAX, MC_Z_REPARAM>, q, grid, smem, stream);
        default: return mc_launch_one(k_mc_lingauss<DIBS_DMAX, MC_LP_ONLY>, q, grid, smem, stream);
    }
}
}  // namespace dibs


#if DIBS_DMAX <= 32
namespace dibs {
// QR-form fast path (observational data); r_packed_host: DMAX*(DMAX+1)/2 floats, see RTri
int DIBS_CAT(launch_mc_linqr_, DIBS_DMAX)(int mode, const McParams& q, dim3 grid, int threads, size_t smem,
                                         const float* r_packed_host, cudaStream_t stream) {
    RTri<DIBS_DMAX> R;
    for (int i = 0; i < DIBS_DMAX * (DIBS_DMAX + 1) / 2; ++i) R.v[i] = r_packed_host[i];
#define DIBS_GO(M)                                                                                          \
    {                                                                                                       \
        auto kern = (q.d == DIBS_DMAX) ? k_mc_lin_qr<DIBS_DMAX, M, true> : k_mc_lin_qr<DIBS_DMAX, M, false>; \
        if (smem > 48 * 1024) {                                                                             \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                            \
        }                                                                                                   \
        kern<<<grid, threads, smem, stream>>>(q, R);                                                        \
        return (int)cudaGetLastError();                                                                     \
    }
    switch (mode) {
        case MC_THETA_HARD: DIBS_GO(MC_THETA_HARD)
        case MC_Z_SCORE: DIBS_GO(MC_Z_SCORE)
        case MC_Z_REPARAM: DIBS_GO(MC_Z_REPARAM)
        default: DIBS_GO(MC_LP_ONLY)
    }
#undef DIBS_GO
}
}  // namespace dibs
#endif
