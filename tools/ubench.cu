// Pipe-throughput microbenchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry
// (fp32 / fp64 SIMT FMA, integer ALU for threefry, MUFU for the logistic/sigmoid transforms).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct CParams { float c[512]; };

// independent chains per thread so latency is hidden; ILP = 8
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(float* out, int iters, float a, float b, CParams cp) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-3f + i;
    if (MODE == 0) {          // FFMA reg,reg,reg
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {   // FFMA with a constant-bank (kernel-parameter) operand
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], cp.c[r * 8 + i], b);
    } else if (MODE == 2) {   // FFMA with two distinct register multiplicands (R*u + acc, u varying)
        float u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) u[i] = b + i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(u[(i + r) & 7], cp.c[r * 8 + i], x[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed fp32x2 FMA (Blackwell FFMA2): 2 FMAs per lane per instruction
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters, float a, float b) {
    unsigned long long x[8];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&av), bb = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(a + threadIdx.x * 1e-3f + i, a - i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = ffma2(x[i], aa, bb);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// threefry2x32-20 blocks, 4 independent blocks per thread per iteration
__global__ void __launch_bounds__(256) k_threefry(uint32_t* out, int iters, uint32_t k0, uint32_t k1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    uint32_t acc = 0;
    uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t x0 = base + q + it * 1024u + k0, x1 = base + q + 77u + k1;
#define R(r) { x0 += x1; x1 = rotl(x1, r); x1 ^= x0; }
            R(13) R(15) R(26) R(6) x0 += k1; x1 += k2 + 1u;
            R(17) R(29) R(16) R(24) x0 += k2; x1 += k0 + 2u;
            R(13) R(15) R(26) R(6) x0 += k0; x1 += k1 + 3u;
            R(17) R(29) R(16) R(24) x0 += k1; x1 += k2 + 4u;
            R(13) R(15) R(26) R(6) x0 += k2; x1 += k0 + 5u;
#undef R
            acc ^= x0 ^ x1;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// logistic + sigmoid transform: logf, log1pf, expf, division (what a soft-graph entry costs)
template <int FAST>
__global__ void __launch_bounds__(256) k_softg(float* out, int iters, float a) {
    float acc = 0;
    float f = 0.001f + (threadIdx.x & 255) * 0.0039f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float u = f + q * 1e-4f + it * 1e-7f;
            float l, g;
            if (FAST) { l = __logf(u) - __logf(1.0f - u); g = __fdividef(1.0f, 1.0f + __expf(-(l + a))); }
            else { l = logf(u) - log1pf(-u); g = 1.0f / (1.0f + expf(-(l + a))); }
            acc += g;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d", prop.name, sms);
    const int blocks = sms * 8, threads = 256, iters = 2000;
    float* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    CParams cp; for (int i = 0; i < 512; ++i) cp.c[i] = 1.0f + 1e-6f * i;
    const double fma_per_thread = (double)iters * 16 * 8;
    const double total = fma_per_thread * blocks * threads;
    float ms;
    ms = time_ms([&] { k_pipe<0><<<blocks, threads>>>(out, iters, 1.0001f, 1e-6f, cp); });
    printf(", \"ffma_rrr_tflops\": %.2f", 2 * total / ms / 1e9);
    ms = time_ms([&] { k_pipe<1><<<blocks, threads>>>(out, iters, 1.0001f, 1e-6f, cp); });
    printf(", \"ffma_rcr_tflops\": %.2f", 2 * total / ms / 1e9);
    ms = time_ms([&] { k_pipe<2><<<blocks, threads>>>(out, iters, 1.0001f, 1e-6f, cp); });
    printf(", \"ffma_rc_acc_tflops\": %.2f", 2 * total / ms / 1e9);
    ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, iters, 1.0001f, 1e-6f); });
    printf(", \"ffma2_tflops\": %.2f", 4 * total / ms / 1e9);
    ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)out, iters, 1.0000001, 1e-9); });
    printf(", \"dfma_tflops\": %.2f", 2 * total / ms / 1e9);
    const int tf_iters = 500;
    ms = time_ms([&] { k_threefry<<<blocks, threads>>>((uint32_t*)out, tf_iters, 0x12345u, 0x6789u); });
    printf(", \"threefry_gblocks_per_s\": %.2f", (double)tf_iters * 4 * blocks * threads / ms / 1e6);
    ms = time_ms([&] { k_softg<0><<<blocks, threads>>>(out, tf_iters, 0.3f); });
    printf(", \"softgraph_gentries_per_s\": %.2f", (double)tf_iters * 4 * blocks * threads / ms / 1e6);
    ms = time_ms([&] { k_softg<1><<<blocks, threads>>>(out, tf_iters, 0.3f); });
    printf(", \"softgraph_fast_gentries_per_s\": %.2f", (double)tf_iters * 4 * blocks * threads / ms / 1e6);
    printf("}\n");
    return 0;
}
