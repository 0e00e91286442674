#include "mc_inst.cuh"
#include "kernels_mc_nn.cuh"
#ifndef DIBS_DMAX
#error "compile with -DDIBS_DMAX=<n>"
#endif
namespace dibs {
template <typename K>
static inline int nn_launch_one(K kernel, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream) {
    const int threads = nn_threads(q.d, q.hidden);
    if (threads > nn_max_threads<DIBS_DMAX>()) return (int)cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kernel<<<grid, threads, smem, stream>>>(q);
    return (int)cudaGetLastError();
}
int DIBS_CAT(launch_mc_nn_, DIBS_DMAX)(int mode, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream) {
    // the reference's default width (hidden_layers=(5,), target.py:271) gets a specialised instance
#if DIBS_DMAX <= 32
    if (q.hidden == 5) {
        switch (mode) {
            case MC_THETA_HARD: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_THETA_HARD, 5>, q, grid, smem, stream);
            case MC_Z_SCORE: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_Z_SCORE, 5>, q, grid, smem, stream);
            case MC_Z_REPARAM: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_Z_REPARAM, 5>, q, grid, smem, stream);
            default: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_LP_ONLY, 5>, q, grid, smem, stream);
        }
    }
#endif
    switch (mode) {
        case MC_THETA_HARD: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_THETA_HARD, 0>, q, grid, smem, stream);
        case MC_Z_SCORE: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_Z_SCORE, 0>, q, grid, smem, stream);
        case MC_Z_REPARAM: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_Z_REPARAM, 0>, q, grid, smem, stream);
        default: return nn_launch_one(k_mc_nn<DIBS_DMAX, MC_LP_ONLY, 0>, q, grid, smem, stream);
    }
}
}  // namespace dibs
