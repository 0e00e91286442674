#include "mc_inst.cuh"
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
