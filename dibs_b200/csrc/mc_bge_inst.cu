#include "mc_inst.cuh"
#include "kernels_mc_bge.cuh"
#include "kernels_mc_bge_soft.cuh"
#ifndef DIBS_DMAX
#error "compile with -DDIBS_DMAX=<n>"
#endif
namespace dibs {
int DIBS_CAT(launch_mc_bge_, DIBS_DMAX)(int mode, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream) {
    if (mode == MC_LP_ONLY) return mc_launch_one(k_mc_bge<DIBS_DMAX, MC_LP_ONLY>, q, grid, smem, stream);
    return mc_launch_one(k_mc_bge<DIBS_DMAX, MC_Z_SCORE>, q, grid, smem, stream);
}
}  // namespace dibs

#if DIBS_DMAX <= 32
namespace dibs {
// soft-graph BGe (MarginalDiBS + 'reparam'): log-probs of caller-supplied graphs, or the reparameterisation pass
int DIBS_CAT(launch_mc_bgesoft_, DIBS_DMAX)(int mode, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream) {
    if (mode == MC_LP_ONLY) return mc_launch_one(k_mc_bge_soft<DIBS_DMAX, MC_LP_ONLY>, q, grid, smem, stream);
    return mc_launch_one(k_mc_bge_soft<DIBS_DMAX, MC_Z_REPARAM>, q, grid, smem, stream);
}
}  // namespace dibs
#endif
