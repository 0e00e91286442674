// Particle / gradient exchange between the GPUs of one node over NVLink peer memory (CUDA IPC), replacing the two
// NCCL all-gathers of the sharded step.  The reference is single-device; this is the exchange step of
// the particle sharding: every rank PUSHES the rows it produced straight into the same rows of every peer's buffer
// (remote stores through NVSwitch), then raises a per-(kind, source-rank) epoch flag in each peer; consumers (the
// pairwise kernels) spin on their LOCAL flags before touching remote-produced rows.  No rendezvous, no staging
// buffer: the cost on the critical path is one small kernel (~M_loc x D x 4 bytes x (world-1) of NVLink stores)
// plus a flag round trip, instead of a collective launch.
//
// Buffer reuse is safe without further handshakes because a rank can run at most one step ahead of a peer (its phi
// of step t needs the peer's gradient flag of step t): particles ping-pong by step parity (already), gradients are
// double-buffered by step parity, and a call-level "done" flag keeps a new dibs_svgd_steps call from packing new
// particles into rows a slower peer is still reading.
#pragma once
#include "common.cuh"

namespace dibs {

constexpr int PEER_MAX = 16;
enum { PEER_KIND_GRAD = 0, PEER_KIND_X = 1, PEER_KIND_DONE = 2, PEER_KIND_SYNC = 3, PEER_KINDS = 4 };

// what a pushing kernel needs: destination base pointers (same buffer on every peer), the peers' flag arrays
struct PeerPush {
    int world, rank, kind;
    float* dst[PEER_MAX];            // peer q's copy of the buffer being pushed (dst[rank] unused)
    uint32_t* flags[PEER_MAX];       // peer q's flag array [PEER_KINDS][PEER_MAX] (flags[rank] = own, local)
    uint32_t* epoch;                 // [PEER_KINDS] this rank's epoch counters (device memory)
    uint32_t* counter;               // [PEER_KINDS] CTA arrival counters
};

// what a consuming kernel needs
struct PeerWait {
    const uint32_t* flags;           // local flag array [PEER_KINDS][PEER_MAX], or null (single GPU / NCCL path)
    const uint32_t* epoch;           // [PEER_KINDS]
    int world, kind;
    uint32_t* error;                 // plan-level error word (mapped host memory): set when a wait times out
    unsigned long long timeout_ns;   // bound of one wait (0: wait forever)
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// every CTA of a consumer calls this first: wait until all ranks' rows of the current epoch have landed
// The spin is BOUNDED: a rank that died or fell out of sequence must not hang every GPU of the node.  After
// timeout_ns the waiter records (kind, missing rank) in the plan's error word and carries on with whatever rows it
// has; the host turns a non-zero error word into DIBS_ERR_STATE at the next call / dibs_plan_status().
__device__ __forceinline__ void peer_wait(const PeerWait& w) {
    if (w.flags == nullptr) return;
    if ((int)threadIdx.x < w.world) {
        const uint32_t want = w.epoch[w.kind];
        const uint32_t* f = w.flags + w.kind * PEER_MAX + threadIdx.x;
        if ((int32_t)(ld_acquire_sys(f) - want) < 0) {
            const unsigned long long t0 = global_timer_ns();
            unsigned spins = 0;
            while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
                __nanosleep(20);
                if (w.timeout_ns && (++spins & 1023u) == 0 && global_timer_ns() - t0 > w.timeout_ns) {
                    if (w.error) *(volatile uint32_t*)w.error = 0x80000000u | ((uint32_t)w.kind << 8) | threadIdx.x;   // benign race: any writer will do
                    break;
                }
            }
        }
    }
    __syncthreads();
}

// End of a producing kernel: every thread has issued its remote stores; the last CTA to arrive bumps the epoch and
// raises the flag of `kind` in every peer (and locally).  Must be called by ALL threads of ALL CTAs of the grid.
__device__ __forceinline__ void peer_signal(const PeerPush& p, unsigned n_ctas) {
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(p.counter + p.kind, 1u) == n_ctas - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x == 0) p.counter[p.kind] = 0u;
    const uint32_t e = p.epoch[p.kind] + 1u;
    __syncthreads();
    if ((int)threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.kind * PEER_MAX + p.rank, e);
    if (threadIdx.x == 0) p.epoch[p.kind] = e;
}

// store one value at the same offset of every peer's copy of a buffer
__device__ __forceinline__ void peer_store(const PeerPush& p, size_t off, float v) {
#pragma unroll 1
    for (int q = 0; q < p.world; ++q)
        if (q != p.rank) p.dst[q][off] = v;
}

// copy `n4` float4 (the rank's own rows, already at their final offset in the local buffer) into every peer, then
// the last CTA to finish bumps the epoch and raises the flags.  n4 == 0: signal only.
static __global__ void __launch_bounds__(256) k_peer_push(PeerPush p, const float4* __restrict__ src, size_t off4, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
        const float4 v = src[off4 + e];
#pragma unroll 1
        for (int q = 0; q < p.world; ++q)
            if (q != p.rank) reinterpret_cast<float4*>(p.dst[q])[off4 + e] = v;
    }
    peer_signal(p, gridDim.x);
}

// stand-alone wait (call-level "done" barrier at the start of dibs_svgd_steps)
static __global__ void k_peer_wait(PeerWait w) { peer_wait(w); }

}  // namespace dibs
