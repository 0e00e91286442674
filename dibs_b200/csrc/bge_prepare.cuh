// One-off BGe data statistics (per dibs_set_data): R_j in fp64, log-gamma table.
// replaces the per-call recomputation in dibs/models/linearGaussian.py:78-107.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "../../include/dibs_b200.h"
#include "common.cuh"

namespace dibs {

// R_j = T + S_N + (N alpha_mu / (N + alpha_mu)) (xbar - mu0)(xbar - mu0)^T over rows where j is not intervened
__global__ void __launch_bounds__(256) k_bge_stats(const float* x, const int32_t* mask, int n_obs, int d,
                                                   const float* mean_obs, double small_t, double alpha_mu,
                                                   double* r_out, float* n_out) {
    extern __shared__ __align__(16) double sd[];
    double* xbar = sd;  // [d]
    const int j = blockIdx.x, tid = threadIdx.x;
    __shared__ double s_n;
    if (tid == 0) {
        double n = 0.0;
        for (int r = 0; r < n_obs; ++r) n += (mask && mask[r * d + j]) ? 0.0 : 1.0;
        s_n = n;
    }
    __syncthreads();
    const double nj = s_n;
    for (int a = tid; a < d; a += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < n_obs; ++r)
            if (!(mask && mask[r * d + j])) s += (double)x[r * d + a];
        xbar[a] = nj > 0.0 ? s / nj : 0.0;
    }
    __syncthreads();
    const double cf = (nj * alpha_mu) / (nj + alpha_mu);
    for (int e = tid; e < d * d; e += blockDim.x) {
        int a = e / d, b = e % d;
        double s = 0.0;
        for (int r = 0; r < n_obs; ++r)
            if (!(mask && mask[r * d + j])) s += ((double)x[r * d + a] - xbar[a]) * ((double)x[r * d + b] - xbar[b]);
        double ma = mean_obs ? (double)mean_obs[a] : 0.0, mb = mean_obs ? (double)mean_obs[b] : 0.0;
        r_out[(size_t)j * d * d + e] = (a == b ? small_t : 0.0) + s + cf * (xbar[a] - ma) * (xbar[b] - mb);
    }
    if (tid == 0) n_out[j] = (float)nj;
}

static inline int bge_prepare(const dibs_config& cfg, int d, int n_obs, const float* x, const int32_t* mask,
                              const float* mean_obs_host, double** r_dev, float** table_dev, float** coef_dev,
                              int* r_stride, cudaStream_t stream, std::string& err) {
    auto bad = [&](cudaError_t e, const char* what) { err = std::string(what) + ": " + cudaGetErrorString(e); return (int)DIBS_ERR_CUDA; };
    const double alpha_mu = cfg.bge_alpha_mu, alpha_lambd = cfg.bge_alpha_lambd;
    if (!(alpha_lambd > d + 1)) { err = "BGe requires alpha_lambd > n_vars + 1"; return DIBS_ERR_INVALID_ARG; }
    const double small_t = (alpha_mu * (alpha_lambd - d - 1)) / (alpha_mu + 1);
    const int n_r = mask ? d : 1;
    cudaError_t e;
    if (*r_dev) cudaFree(*r_dev);
    if (*table_dev) cudaFree(*table_dev);
    if (*coef_dev) cudaFree(*coef_dev);
    float* mean_dev = nullptr;
    float* n_dev = nullptr;
    if ((e = cudaMalloc((void**)r_dev, (size_t)n_r * d * d * sizeof(double))) != cudaSuccess) return bad(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)table_dev, (size_t)d * (d + 1) * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)coef_dev, (size_t)d * 2 * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc");
    if ((e = cudaMalloc((void**)&n_dev, d * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc");
    if (mean_obs_host) {
        if ((e = cudaMalloc((void**)&mean_dev, d * sizeof(float))) != cudaSuccess) return bad(e, "cudaMalloc");
        cudaMemcpyAsync(mean_dev, mean_obs_host, d * sizeof(float), cudaMemcpyHostToDevice, stream);
    }
    k_bge_stats<<<n_r, 256, d * sizeof(double), stream>>>(x, mask, n_obs, d, mean_dev, small_t, alpha_mu, *r_dev, n_dev);
    if ((e = cudaGetLastError()) != cudaSuccess) return bad(e, "k_bge_stats");
    std::vector<float> nj(d);
    if ((e = cudaMemcpyAsync(nj.data(), n_dev, n_r * sizeof(float), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return bad(e, "memcpy");
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return bad(e, "sync");
    for (int j = n_r; j < d; ++j) nj[j] = nj[0];
    // log-gamma terms depend on (N_j, n_parents) only (linearGaussian.py:99-107): tabulate in double
    std::vector<float> table((size_t)d * (d + 1)), coef((size_t)d * 2);
    for (int j = 0; j < d; ++j) {
        const double n = nj[j];
        coef[2 * j] = (float)(n + alpha_lambd - d);
        coef[2 * j + 1] = n > 0.0 ? 1.0f : 0.0f;   // jnp.where(isclose(N, 0), 0.0, .)
        for (int l = 0; l <= d; ++l) {
            double v = 0.5 * (std::log(alpha_mu) - std::log(n + alpha_mu)) + std::lgamma(0.5 * (n + alpha_lambd - d + l + 1)) -
                       std::lgamma(0.5 * (alpha_lambd - d + l + 1)) - 0.5 * n * std::log(M_PI) +
                       0.5 * (alpha_lambd - d + 2 * l + 1) * std::log(small_t);
            table[(size_t)j * (d + 1) + l] = (float)v;
        }
    }
    cudaMemcpyAsync(*table_dev, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(*coef_dev, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return bad(e, "sync");
    cudaFree(n_dev);
    if (mean_dev) cudaFree(mean_dev);
    *r_stride = mask ? d * d : 0;
    return DIBS_OK;
}

}  // namespace dibs
