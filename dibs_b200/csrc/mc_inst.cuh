// Per-(family, DMAX) instantiation unit: compiled once per value of DIBS_DMAX / DIBS_FAMILY so that the
// register-tiled kernels build in parallel (see dibs_b200/build.py).
#pragma once
#include "kernels_mc.cuh"

namespace dibs {
// returns a cudaError_t (as int); 0 on success
typedef int (*mc_launch_fn)(int mode, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream);

template <typename K>
static inline int mc_launch_one(K kernel, const McParams& q, dim3 grid, size_t smem, cudaStream_t stream) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kernel<<<grid, 256, smem, stream>>>(q);
    return (int)cudaGetLastError();
}
}  // namespace dibs

#define DIBS_CAT_(a, b) a##b
#define DIBS_CAT(a, b) DIBS_CAT_(a, b)
