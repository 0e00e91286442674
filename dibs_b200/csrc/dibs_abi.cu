// C ABI of dibs_b200 (see include/dibs_b200.h): plan, workspace, launch logic, CUDA-graph replay, NCCL.
#include "../../include/dibs_b200.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_mc.cuh"
#include "kernels_mc_lin_qr.cuh"
#include "kernels_mc_bge.cuh"
#include "kernels_mc_bge_soft.cuh"
#include "bge_prepare.cuh"
#include "kernels_mc_nn.cuh"
#include "kernels_prior.cuh"
#include "kernels_acyclic.cuh"
#include "kernels_dense.cuh"
#include "kernels_pair.cuh"
#include "kernels_phi_mma.cuh"
#include "kernels_init.cuh"

using namespace dibs;

// ------------------------------------------------------------------------------------------
// errors, launch accounting
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(DIBS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e) + " (" + \
                                           __FILE__ + ":" + std::to_string(__LINE__) + ")");       \
    } while (0)

#define LAUNCHED()                                                                                 \
    do {                                                                                           \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
        CU(cudaGetLastError());                                                                    \
    } while (0)

#define TRY(expr)                \
    do {                         \
        int _r = (expr);         \
        if (_r != DIBS_OK) return _r; \
    } while (0)

extern "C" const char* dibs_last_error(void) { return g_last_error.c_str(); }
extern "C" int dibs_abi_version(void) { return 1; }
extern "C" int64_t dibs_launch_count(void) { return (int64_t)g_launches.load(); }

// ------------------------------------------------------------------------------------------
// NCCL through dlopen (the process normally has torch's libnccl.so.2 loaded already)
// ------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*CommAbort)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return DIBS_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    if (const char* env = getenv("DIBS_B200_NCCL_LIB")) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    for (int i = 0; i < 2 && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(DIBS_ERR_NCCL, std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
    g_nccl.CommAbort = (int (*)(NcclComm))dlsym(lib, "ncclCommAbort");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy)
        return fail(DIBS_ERR_NCCL, "libnccl is missing a required symbol");
    g_nccl.lib = lib;
    return DIBS_OK;
}
#define NC(call)                                                                                     \
    do {                                                                                             \
        int _r = (call);                                                                             \
        if (_r != 0)                                                                                 \
            return fail(DIBS_ERR_NCCL, std::string(#call) + ": " +                                   \
                                           (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?")); \
    } while (0)

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
struct dibs_plan {
    dibs_config cfg;
    int d, k, M, M_loc, row0, Dz, Dth, D, ld;
    int N = 0;
    bool has_data = false, has_mask = false;
    int dmax = 0;
    // derived fp32 constants, rounded like the reference
    float s2, log2pis2, sig2_edge, lognorm_edge, sig2_param, lognorm_param, er_coef, sigma_z2;
    // data
    float* x = nullptr; int32_t* mask = nullptr;
    double* bge_r = nullptr; float* bge_table = nullptr; float* bge_coef = nullptr; int bge_r_stride = 0;
    // particle state: two packed buffers [M][ld], row = [Z | Theta | dZ | dTheta]
    float* pk[2] = {nullptr, nullptr};   // particles [M][D], row = [Z | Theta], ping-pong; a rank updates its own rows
    float* gk[2] = {nullptr, nullptr};   // log-prob gradients [M][D], row = [dZ | dTheta], by step parity
    // peer-memory exchange (kernels_peer.cuh): flags / epochs / counters, and the peers' buffers opened through CUDA IPC
    uint32_t* peer_flags_local = nullptr;   // [PEER_KINDS][PEER_MAX]
    uint32_t* peer_epoch = nullptr;         // [PEER_KINDS]
    uint32_t* peer_counter = nullptr;       // [PEER_KINDS]
    bool p2p = false;
    uint32_t* peer_error = nullptr;         // mapped host word: a bounded peer wait that timed out leaves (kind, rank) here
    uint32_t* peer_error_dev = nullptr;     // its device alias
    unsigned long long peer_timeout_ns = 0;
    float* summary_ws = nullptr; size_t summary_cap = 0;   // dibs_particle_summary scratch
    bool step_ws = false;                   // the step workspace below is allocated on first use (hook-only plans never need it)
    float* peer_pk[2][PEER_MAX] = {};
    float* peer_gk[2][PEER_MAX] = {};
    uint32_t* peer_flags[PEER_MAX] = {};
    std::vector<void*> ipc_opened;
    float* v = nullptr;            // [M_loc][D]
    float* base = nullptr;         // [M_loc]
    StepState* st = nullptr;
    uint32_t* step_keys = nullptr;  // [n_splits][M_loc][2] sub-keys of the current step (k_prologue)
    float* scores = nullptr;        // [M_loc][d*d] raw U V^T of the current step (k_prologue)
    // LinearGaussian, observational data, d <= 32: packed upper-triangular QR factor of x (kernels_mc_lin_qr.cuh)
    bool use_qr = false;
    std::vector<float> lin_r;
    // LinearGaussian, observational data, n_vars > 32: dense Rx and Rx^T on the device (kernels_dense.cuh)
    bool use_dense = false;
    float* rx_dense = nullptr;
    // MC workspace
    int max_chunks = 1, th_acc_size = 0;
    float *th_acc = nullptr, *th_stats = nullptr, *z_acc = nullptr, *z_stats = nullptr, *acyc = nullptr;
    // pairwise workspace
    int n_split = 1, n_split_z = 1, split_len_z = 0, split_len_t = 0;
    float *dist_part = nullptr, *kz = nullptr, *kt = nullptr, *kfull = nullptr;
    int n_jsplit = 1, j_len = 0;
    float* phi_part = nullptr;     // [n_jsplit][M_loc][D]
    // arrival counters of the in-kernel reductions (zero between launches): gradient CTAs per particle (fused
    // assemble), feature splits per K tile, j slices per phi tile
    uint32_t *arrive = nullptr, *phi_cnt = nullptr;
    // tensor-core phi (kernels_phi_mma.cuh): TMA tensor maps of the K planes and, per buffer parity, of the particle
    // and gradient buffers; chosen by the GLOBAL particle count only, so every world size takes the same arithmetic
    bool phi_mma = false;
    PhiMmaMaps mma_maps[2];
    float* k_split = nullptr;      // [6][M_loc][M] TF32 hi / lo parts of K, K_z, K_theta (k_pair_finish)
    // CUDA graphs of one step, per buffer parity
    cudaGraphExec_t gexec[2] = {nullptr, nullptr};
    int kernels_per_step = 0;
    bool use_graph = true;
    cudaStream_t cap_stream = nullptr;
    // fork/join inside a step: the two MC passes, the acyclicity pass and (single GPU) the kernel-matrix pass are
    // mutually independent, each too small to fill 148 SMs alone -> they run on sibling streams (graph branches)
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork[2] = {nullptr, nullptr};
    cudaEvent_t ev_join[3] = {nullptr, nullptr, nullptr};
    bool concurrent = true;
    // NCCL
    NcclComm comm = nullptr;        // gradient rows (critical path of the step)
    NcclComm comm_x = nullptr;      // particle rows (side branch, hidden behind the gradient phase)
    // optional per-kernel event timing (dibs_svgd_steps_timed): events recorded after each launch of an eager step
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_phase;     // phase id of the work that ENDS at ev_pool[i]; -1 = start marker
    size_t ev_used = 0;
    // timeline mode of dibs_svgd_steps_timed (per_kernel = 2): the step graph with an event behind every kernel
    int tl_capture = -1;
    cudaGraphExec_t tl_exec[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> tl_ev[2];
    std::vector<int> tl_phase[2];
};

// mark the end of one kernel (phase) on `stream` when the plan is in timing mode
static void mark(dibs_plan* p, cudaStream_t stream, int phase) {
    if (p->tl_capture >= 0) {
        // timeline mode: an external event-record node behind the kernel, on the branch the kernel runs on
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        cudaEventRecordWithFlags(e, stream, cudaEventRecordExternal);
        p->tl_ev[p->tl_capture].push_back(e);
        p->tl_phase[p->tl_capture].push_back(phase);
        return;
    }
    if (!p->timing) return;
    if (p->ev_used == p->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        p->ev_pool.push_back(e);
        p->ev_phase.push_back(0);
    }
    p->ev_phase[p->ev_used] = phase;
    cudaEventRecord(p->ev_pool[p->ev_used++], stream);
}

static int pick_dmax(int d) {
    const int opts[] = {8, 16, 20, 32, 64, 128};
    for (int o : opts) if (d <= o) return o;
    return 0;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// CTAs per particle of the acyclicity pass (a function of (A, d, PRNG layout) only)
constexpr int ACYC_WPC = ACYC_SLOTS / 2;      // sample pairs per CTA of the row-per-thread kernel

static int mc_target_ctas(int per_sm) { return per_sm * 148; }
static bool acyc_rows_path(const dibs_plan* p) {
    return p->d <= 32 && (p->cfg.n_acyclicity_mc_samples % 2) == 0 && !p->cfg.prng_partitionable;
}
static int acyc_chunks(const dibs_plan* p) {
    if (p->d > 32 && p->d <= 64) return acyc_dense4_shape(p->d, p->cfg.n_acyclicity_mc_samples).chunks;
    if (p->d > 64) return acyc_dense_shape(p->d, p->cfg.n_acyclicity_mc_samples).chunks;
    return acyc_rows_path(p) ? ceil_div(p->cfg.n_acyclicity_mc_samples / 2, ACYC_WPC) : 1;
}

extern "C" int dibs_theta_dim(const dibs_plan* plan) { return plan ? plan->Dth : 0; }

// How one Monte-Carlo pass over `n_local` particles x S samples is cut into CTAs.
struct McShape {
    bool qr;        // QR-form LinearGaussian kernel: the units below are sample PAIRS (slots)
    int gpb;        // graphs (or slots) per block-round
    int chunks;     // CTAs per particle (blockIdx.y)
    int spc;        // samples (or slots) per chunk
    int threads;    // CTA size
    bool paired;    // a slot is the sample pair (s, s + S/2)
    bool dense;     // LinearGaussian n_vars > 32: whole-graph matrix-product kernel, units are samples
};

static bool qr_eligible(int likelihood, int dmax) { return likelihood == DIBS_LIK_LINEAR_GAUSSIAN && dmax <= 32; }

static McShape mc_shape_for(bool qr, int likelihood, int d, int hidden, int n_local, int S, bool pair_ok = false) {
    McShape sh;
    sh.qr = qr;
    sh.paired = false;
    sh.dense = false;
    if (n_local < 1) n_local = 1;
    if (qr) {
        const int Q = (S + 1) / 2;
        int best = 1; long best_cost = -1;
        for (int g = 1; g * d <= 192 && g <= Q && g <= 16; ++g) {       // <= 16 slots: 32 samples per warp-wide reduction
            long cost = (long)ceil_div(Q, g) * (((g * d) + 31) / 32 * 32);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = g; }
        }
        sh.gpb = best;
        sh.threads = ((best * d) + 31) / 32 * 32;
        const int rounds = ceil_div(Q, best);
        // up to 4 chunks per particle whatever the number of particles a rank owns: the chunking fixes the order in which
        // the online-softmax partials are merged, so it must not depend on the rank count (bit-identical results).
        // 4 x 128 particles x 2 passes + the acyclicity CTAs still oversubscribe 148 SMs on 8 GPUs; 8 chunks cost 4 %
        // of the step on one GPU (measured) for nothing
        int chunks = 4;
        (void)n_local;
        if (chunks > rounds) chunks = rounds;
        if (chunks < 1) chunks = 1;
        const int rpc = ceil_div(rounds, chunks);
        sh.chunks = ceil_div(rounds, rpc);
        sh.spc = rpc * best;
        return sh;
    }
    if (likelihood == DIBS_LIK_BGE) {
        // whole chunk of slots (sample pairs when the PRNG layout allows) per CTA, as many as shared memory holds
        sh.paired = pair_ok && (S % 2) == 0;
        const int per = sh.paired ? 2 : 1;
        const int Q = sh.paired ? S / 2 : S;
        int want = ceil_div(mc_target_ctas(2), n_local);
        if (want > Q) want = Q;
        if (want < 1) want = 1;
        int spc = ceil_div(Q, want);
        const int dmax = d <= 8 ? 8 : d <= 16 ? 16 : d <= 20 ? 20 : d <= 32 ? 32 : 64;
        while (spc > 1 && (spc * per > 256 || mc_bge_smem(d, dmax, spc * per, true) > 200 * 1024)) spc = (spc + 1) / 2;
        sh.spc = spc;
        sh.chunks = ceil_div(Q, spc);
        sh.gpb = spc;
        sh.threads = 256;
        return sh;
    }
    if (likelihood == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN) {
        // one slot (sample pair when the PRNG layout allows, else one sample) at a time per CTA; chunks of slots
        sh.paired = pair_ok && (S % 2) == 0;
        const int Q = sh.paired ? S / 2 : S;
        int want = ceil_div(mc_target_ctas(4), n_local);
        if (want > Q) want = Q;
        if (want < 1) want = 1;
        sh.spc = ceil_div(Q, want);
        sh.chunks = ceil_div(Q, sh.spc);
        sh.gpb = 1;
        sh.threads = nn_threads(d, hidden);
        return sh;
    }
    int per_item = d;
    int gpb = 256 / per_item; if (gpb < 1) gpb = 1; if (gpb > S) gpb = S;
    sh.gpb = gpb;
    sh.threads = 256;
    int max_chunks = ceil_div(S, gpb);
    int want = ceil_div(mc_target_ctas(2), n_local);
    if (want < 1) want = 1;
    if (want > max_chunks) want = max_chunks;
    sh.spc = ceil_div(ceil_div(S, want), gpb) * gpb;
    sh.chunks = ceil_div(S, sh.spc);
    return sh;
}

static McShape mc_shape_dense(int d, int n_local, int S, bool pair_ok = false) {
    McShape sh;
    sh.qr = false; sh.dense = true;
    sh.paired = false;
    (void)pair_ok;
    if (n_local < 1) n_local = 1;
    const int U = S;                                     // units per particle: samples
    int want = ceil_div(mc_target_ctas(4), n_local);
    if (want > U) want = U;
    if (want < 1) want = 1;
    sh.spc = ceil_div(U, want);
    sh.chunks = ceil_div(U, sh.spc);
    sh.gpb = 1;
    sh.threads = ((dense_nt(d) + 31) / 32) * 32;
    return sh;
}

static bool phi_mma_eligible(int n_all, int ld, int min_particles);

extern "C" int dibs_plan_create(const dibs_config* cfg, dibs_plan** out) {
    if (!cfg || !out) return fail(DIBS_ERR_INVALID_ARG, "null argument");
    const dibs_config& c = *cfg;
    if (c.n_vars < 2 || c.n_dim < 1 || c.n_particles < 1) return fail(DIBS_ERR_INVALID_ARG, "n_vars>=2, n_dim>=1, n_particles>=1 required");
    if (c.n_grad_mc_samples < 1 || c.n_acyclicity_mc_samples < 1) return fail(DIBS_ERR_INVALID_ARG, "MC sample counts must be positive");
    if (c.grad_estimator_z != DIBS_ESTIMATOR_SCORE && c.grad_estimator_z != DIBS_ESTIMATOR_REPARAM)
        return fail(DIBS_ERR_INVALID_ARG, "Unknown gradient estimator");
    if (c.optimizer != DIBS_OPT_GD && c.optimizer != DIBS_OPT_RMSPROP) return fail(DIBS_ERR_INVALID_ARG, "unknown optimizer");
    if (c.graph_prior < 0 || c.graph_prior > 2) return fail(DIBS_ERR_INVALID_ARG, "unknown graph prior");
    if (c.world_size < 1 || c.rank < 0 || c.rank >= c.world_size || c.n_particles % c.world_size)
        return fail(DIBS_ERR_INVALID_ARG, "n_particles must be divisible by world_size and 0 <= rank < world_size");
    if (c.joint && c.likelihood == DIBS_LIK_BGE) return fail(DIBS_ERR_INVALID_ARG, "BGe is a marginal likelihood: use MarginalDiBS");
    if (!c.joint && c.likelihood != DIBS_LIK_BGE) return fail(DIBS_ERR_UNSUPPORTED, "MarginalDiBS is implemented for BGe only");
    if (c.likelihood == DIBS_LIK_BGE && c.grad_estimator_z != DIBS_ESTIMATOR_SCORE && c.n_vars > 32)
        return fail(DIBS_ERR_UNSUPPORTED, "BGe + reparam estimator (soft-graph BGe) is implemented for n_vars <= 32");
    if (c.likelihood < 0 || c.likelihood > 2) return fail(DIBS_ERR_UNSUPPORTED, "unknown likelihood model");
    if (c.likelihood == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN && (c.activation < 0 || c.activation > 3))
        return fail(DIBS_ERR_INVALID_ARG, "Invalid activation function");      /* KeyError in the reference (nonlinearGaussian.py:61) */
    if (c.likelihood == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN && (c.hidden < 1 || c.hidden > 32))
        return fail(DIBS_ERR_UNSUPPORTED, "DenseNonlinearGaussian: one hidden layer of 1..32 units is implemented");
    if (pick_dmax(c.n_vars) == 0) return fail(DIBS_ERR_UNSUPPORTED, "n_vars > 128 is not implemented");
    if (c.likelihood == DIBS_LIK_BGE && c.n_vars > 64) return fail(DIBS_ERR_UNSUPPORTED, "BGe: n_vars > 64 is not implemented");

    dibs_plan* p = new dibs_plan();
    p->cfg = c;
    p->d = c.n_vars; p->k = c.n_dim; p->M = c.n_particles;
    p->M_loc = c.n_particles / c.world_size; p->row0 = c.rank * p->M_loc;
    p->Dz = 2 * p->d * p->k;
    p->Dth = 0;
    if (c.likelihood == DIBS_LIK_LINEAR_GAUSSIAN) p->Dth = p->d * p->d;
    if (c.likelihood == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN) p->Dth = p->d * (p->d * c.hidden + 2 * c.hidden + 1);
    p->D = p->Dz + p->Dth;
    p->ld = (p->D + 3) & ~3;                 // row stride: a multiple of 4 floats (128-bit pushes / loads)
    p->dmax = pick_dmax(p->d);
    // fp32 constants with the reference's rounding: scale = sqrt(obs_noise); scale^2; log(2 pi scale^2)
    float scale = sqrtf(c.obs_noise);
    p->s2 = scale * scale;
    p->log2pis2 = logf(6.283185307179586f * p->s2);
    p->sig2_edge = c.sig_edge * c.sig_edge;
    p->lognorm_edge = logf(6.283185307179586f * p->sig2_edge);
    p->sig2_param = c.sig_param * c.sig_param;
    p->lognorm_param = logf(6.283185307179586f * p->sig2_param);
    p->er_coef = logf(c.er_p) - logf(1.0f - c.er_p);       // NaN / inf for p >= 1 like graph.py:108
    p->sigma_z2 = powf(c.latent_prior_std, 2.0f);           // latent_prior_std ** 2.0 (dibs.py:657)
    if (const char* e = getenv("DIBS_B200_NO_GRAPH")) p->use_graph = !(e[0] == '1');      // debugging: eager launches
    if (const char* e = getenv("DIBS_B200_SERIAL")) p->concurrent = !(e[0] == '1');       // debugging: no sibling branches

    const int d = p->d, S = c.n_grad_mc_samples;
    {
        // workspace for the larger of the two possible pass shapes (the QR path is chosen in dibs_set_data)
        McShape a = mc_shape_for(false, c.likelihood, d, c.hidden, p->M_loc, S, false);
        p->max_chunks = a.chunks;
        {
            McShape b = mc_shape_for(false, c.likelihood, d, c.hidden, p->M_loc, S, true);
            if (b.chunks > p->max_chunks) p->max_chunks = b.chunks;
            if (c.likelihood == DIBS_LIK_LINEAR_GAUSSIAN && d > 32) {
                b = mc_shape_dense(d, p->M_loc, S);
                if (b.chunks > p->max_chunks) p->max_chunks = b.chunks;
            }
        }

        if (qr_eligible(c.likelihood, p->dmax)) {
            McShape b = mc_shape_for(true, c.likelihood, d, c.hidden, p->M_loc, S);
            if (b.chunks > p->max_chunks) p->max_chunks = b.chunks;
        }
    }
    p->th_acc_size = p->Dth;

    // pairwise: split the feature axis until the distance pass has >= ~2 x 148 CTAs; Z and Theta features are cut
    // separately so that a split never straddles the Z | Theta boundary
    {
        // the split is a function of (M, D) only -- sized for the 8-way sharded case -- so K, and everything downstream,
        // is bit-identical for any number of ranks
        int tiles = ceil_div(p->M, KT) * ceil_div(p->M, KT);
        int ns = ceil_div(8 * 2 * 148, tiles);
        int max_ns = ceil_div(p->D, 2 * KF);
        if (ns > max_ns) ns = max_ns;
        if (ns < 1) ns = 1;
        int len = ceil_div(ceil_div(p->D, ns), KF) * KF;
        p->split_len_z = len < p->Dz ? len : ceil_div(p->Dz, KF) * KF;
        p->n_split_z = ceil_div(p->Dz, p->split_len_z);
        p->split_len_t = p->Dth ? (len < p->Dth ? len : ceil_div(p->Dth, KF) * KF) : KF;
        p->n_split = p->n_split_z + (p->Dth ? ceil_div(p->Dth, p->split_len_t) : 0);
    }
    // phi: slices of the j axis so that the pass fills the GPU even when a rank owns few rows; the slice length is
    // a function of (M, D) only -- never of the rank count -- so results are bit-identical for any world size
    {
        const int col_tiles = ceil_div(p->Dz, PT_C) + ceil_div(p->Dth, PT_C);
        const int row_scale = p->M / 256 > 1 ? p->M / 256 : 1;
        const int phi_ctas = 592;                       // ~4 CTAs of 128 threads per SM
        int ns = ceil_div(phi_ctas, col_tiles * row_scale);
        const int max_ns = p->M / 64 > 1 ? p->M / 64 : 1;
        if (ns > max_ns) ns = max_ns;
        if (ns < 1) ns = 1;
        p->j_len = ceil_div(ceil_div(p->M, ns), PT_J) * PT_J;
        p->n_jsplit = ceil_div(p->M, p->j_len);
        // tensor-core phi (>= MMA_MIN_PARTICLES particles, see phi_mma_eligible): a CTA pays ~5 us of fixed cost
        // (pipeline fill, drain, epilogue) and 1.23 us per 32 particles (tools/phi_trace.py: MMA issue 0.77 us of it on a
        // 2-stage ring), so slices are longer than the SIMT kernel's -- 256 particles, again a function of M only:
        // 8 stages per CTA, and still 80 CTAs when 8 ranks own 128 rows each
        const int ld_rows = (p->D + 3) & ~3;
        if (phi_mma_eligible(p->M, ld_rows, 512)) {
            p->j_len = 256;
            p->n_jsplit = ceil_div(p->M, p->j_len);
        }
    }
    *out = p;
    return DIBS_OK;
}

// ---- TMA tensor maps (driver entry point resolved at run time: the library does not link libcuda) -----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)sym;
    }();
    return fn;
}
// fp32 row-major matrix [rows][cols] with row stride `ld` floats; box = 32 columns (one 128-byte swizzle row) x box_rows
static int make_map(CUtensorMap* m, const float* base, size_t rows, size_t cols, size_t ld, unsigned box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(DIBS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DIBS_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return DIBS_OK;
}
// The tensor-core kernel is taken from MMA_MIN_PARTICLES particles on (a function of the GLOBAL count, so every world
// size runs the same arithmetic): below that its 128-row x 64-column tiles are too few to fill 148 SMs and the SIMT
// kernel wins (measured at 256 particles: 67 vs 39 us).
constexpr int MMA_MIN_PARTICLES = 512;
static bool phi_mma_eligible(int n_all, int ld, int min_particles = MMA_MIN_PARTICLES) {
    static const bool off = getenv("DIBS_B200_PHI_SIMT") && getenv("DIBS_B200_PHI_SIMT")[0] == '1';      // debugging: SIMT phi everywhere
    return !off && n_all >= min_particles && (n_all % 4) == 0 && (ld % 4) == 0;
}
static int fill_mma_maps(PhiMmaMaps& mm, const float* k_split, bool has_theta, int n_rows, int n_all,
                         const float* x_all, const float* g_all, int ld) {
    const size_t plane = (size_t)n_rows * n_all;
    for (int i = 0; i < 6; ++i) TRY(make_map(&mm.a[i], k_split + (size_t)((i < 4 || has_theta) ? i : i - 2) * plane, n_rows, n_all, n_all, MM_ROWS));
    // MN-major 32-bit operands: 32-byte swizzle atoms (see kernels_phi_mma.cuh)
    TRY(make_map(&mm.b_x, x_all, n_all, ld, ld, MM_KS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    TRY(make_map(&mm.b_g, g_all, n_all, ld, ld, MM_KS, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    return DIBS_OK;
}

// Workspace of the step loop (particle / gradient buffers, Monte-Carlo partials, pairwise planes): allocated by the
// first dibs_svgd_steps / dibs_plan_ipc_export call, so plans that only serve the per-function hooks stay a few KB.
static int ensure_step_ws(dibs_plan* p) {
    if (p->step_ws) return DIBS_OK;
    const int d = p->d;
    auto alloc = [&](void** ptr, size_t bytes) -> int {
        CU(cudaMalloc(ptr, bytes ? bytes : 4));
        CU(cudaMemset(*ptr, 0, bytes ? bytes : 4));
        return DIBS_OK;
    };
    int r = DIBS_OK;
    if (!p->peer_error) {
        CU(cudaHostAlloc((void**)&p->peer_error, sizeof(uint32_t), cudaHostAllocMapped));
        *p->peer_error = 0u;
        CU(cudaHostGetDevicePointer((void**)&p->peer_error_dev, p->peer_error, 0));
        const char* e = getenv("DIBS_B200_PEER_TIMEOUT_MS");
        p->peer_timeout_ns = (unsigned long long)(e && *e ? atoll(e) : 10000) * 1000000ull;
    }
    // buffers that may be shared through CUDA IPC get whole 2 MiB blocks of their own
    size_t pk_bytes = (((size_t)p->M * p->ld * sizeof(float)) + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    if ((r = alloc((void**)&p->pk[0], pk_bytes)) || (r = alloc((void**)&p->pk[1], pk_bytes)) ||
        (r = alloc((void**)&p->gk[0], pk_bytes)) || (r = alloc((void**)&p->gk[1], pk_bytes)) ||
        (r = alloc((void**)&p->peer_flags_local, 2u << 20)) ||
        (r = alloc((void**)&p->peer_epoch, PEER_KINDS * sizeof(uint32_t))) ||
        (r = alloc((void**)&p->peer_counter, PEER_KINDS * sizeof(uint32_t))) ||
        (r = alloc((void**)&p->v, (size_t)p->M_loc * p->D * sizeof(float))) ||
        (r = alloc((void**)&p->base, (size_t)p->M_loc * sizeof(float))) ||
        (r = alloc((void**)&p->st, 2 * sizeof(StepState))) ||
        (r = alloc((void**)&p->step_keys, (size_t)3 * p->M_loc * 2 * sizeof(uint32_t))) ||
        (r = alloc((void**)&p->scores, (size_t)p->M_loc * p->d * p->d * sizeof(float))) ||
        (r = alloc((void**)&p->z_acc, (size_t)p->M_loc * p->max_chunks * d * d * sizeof(float))) ||
        (r = alloc((void**)&p->z_stats, (size_t)p->M_loc * p->max_chunks * 4 * sizeof(float))) ||
        (r = alloc((void**)&p->th_acc, (size_t)p->M_loc * p->max_chunks * p->th_acc_size * sizeof(float))) ||
        (r = alloc((void**)&p->th_stats, (size_t)p->M_loc * p->max_chunks * 4 * sizeof(float))) ||
        (r = alloc((void**)&p->acyc, (size_t)p->M_loc * acyc_chunks(p) * d * d * sizeof(float))))
        return r;
    size_t plane = (size_t)p->M_loc * p->M * sizeof(float);
    if ((r = alloc((void**)&p->arrive, (size_t)p->M_loc * sizeof(uint32_t))) ||
        (r = alloc((void**)&p->phi_cnt, (size_t)(ceil_div(p->Dz, PT_C) + ceil_div(p->Dth, PT_C)) * ceil_div(p->M_loc, 32) * sizeof(uint32_t))) ||
        (r = alloc((void**)&p->dist_part, plane * p->n_split)) || (r = alloc((void**)&p->kz, plane)) ||
        (r = alloc((void**)&p->kt, plane)) || (r = alloc((void**)&p->kfull, plane)) ||
        (r = alloc((void**)&p->phi_part, (size_t)p->n_jsplit * p->M_loc * p->D * sizeof(float))))
        return r;
    p->phi_mma = phi_mma_eligible(p->M, p->ld);
    if (p->phi_mma) {
        if ((r = alloc((void**)&p->k_split, 6 * plane))) return r;
        for (int par = 0; par < 2; ++par)
            TRY(fill_mma_maps(p->mma_maps[par], p->k_split, p->Dth > 0, p->M_loc, p->M, p->pk[par], p->gk[par], p->ld));
    }
    p->step_ws = true;
    return DIBS_OK;
}

extern "C" int dibs_plan_status(dibs_plan* p) {
    if (!p) return fail(DIBS_ERR_INVALID_ARG, "null plan");
    if (p->peer_error && *(volatile uint32_t*)p->peer_error) {
        const uint32_t w = *(volatile uint32_t*)p->peer_error;
        static const char* kinds[] = {"gradient rows", "particle rows", "call-done flag", "sync flag"};
        return fail(DIBS_ERR_STATE, std::string("peer-memory exchange timed out waiting for the ") + kinds[(w >> 8) & 3] +
                                        " of rank " + std::to_string(w & 0xff) + " (a peer died or the ranks ran different step counts)");
    }
    return DIBS_OK;
}

extern "C" int dibs_plan_destroy(dibs_plan* p) {
    if (!p) return DIBS_OK;
    for (int i = 0; i < 2; ++i) if (p->gexec[i]) cudaGraphExecDestroy(p->gexec[i]);
    for (int i = 0; i < 2; ++i) { if (p->tl_exec[i]) cudaGraphExecDestroy(p->tl_exec[i]); for (cudaEvent_t e : p->tl_ev[i]) cudaEventDestroy(e); }
    for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
    if (!p->ipc_opened.empty()) {
        cudaDeviceSynchronize();
        for (void* q : p->ipc_opened) cudaIpcCloseMemHandle(q);
    }
    if (p->comm_x) {
        cudaDeviceSynchronize();
        if (g_nccl.CommAbort) g_nccl.CommAbort(p->comm_x);
        else if (g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm_x);
    }
    if (p->comm) {
        // plans are torn down whenever the host garbage-collects them, not at a point all ranks agree on: abort is
        // the non-collective teardown (ncclCommDestroy may wait for the peers); nothing is in flight after the sync
        cudaDeviceSynchronize();
        if (g_nccl.CommAbort) g_nccl.CommAbort(p->comm);
        else if (g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm);
    }
    if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
    for (int i = 0; i < 3; ++i) { if (p->aux[i]) cudaStreamDestroy(p->aux[i]); if (p->ev_join[i]) cudaEventDestroy(p->ev_join[i]); }
    for (int i = 0; i < 2; ++i) if (p->ev_fork[i]) cudaEventDestroy(p->ev_fork[i]);
    if (p->peer_error) cudaFreeHost(p->peer_error);
    void* ptrs[] = {p->summary_ws, p->arrive, p->phi_cnt, p->k_split, p->rx_dense, p->x, p->mask, p->bge_r, p->bge_table, p->bge_coef, p->pk[0], p->pk[1], p->gk[0], p->gk[1], p->peer_flags_local, p->peer_epoch, p->peer_counter, p->v, p->base, p->st, p->step_keys, p->scores,
                    p->th_acc, p->th_stats, p->z_acc, p->z_stats, p->acyc, p->dist_part, p->kz, p->kt, p->kfull, p->phi_part};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
    return DIBS_OK;
}

// Thin QR of x [n_obs, d] by Householder reflections in fp64; returns the upper-triangular factor (rows beyond
// min(n_obs, d) are zero) packed like RTri<dmax> and rounded to fp32.  Only R^T R = x^T x matters downstream, so the
// sign convention of the reflections is irrelevant.
static void qr_upper_packed(const std::vector<float>& x, int n_obs, int d, int dmax, std::vector<float>& out) {
    std::vector<double> a((size_t)n_obs * d);
    for (size_t i = 0; i < a.size(); ++i) a[i] = (double)x[i];
    const int r = n_obs < d ? n_obs : d;
    std::vector<double> v(n_obs);
    for (int c = 0; c < r; ++c) {
        double norm = 0.0;
        for (int i = c; i < n_obs; ++i) norm += a[(size_t)i * d + c] * a[(size_t)i * d + c];
        norm = std::sqrt(norm);
        if (norm == 0.0) continue;
        const double a0 = a[(size_t)c * d + c];
        const double alpha = a0 > 0.0 ? -norm : norm;
        double vnorm2 = 0.0;
        for (int i = c; i < n_obs; ++i) { v[i] = a[(size_t)i * d + c] - (i == c ? alpha : 0.0); vnorm2 += v[i] * v[i]; }
        if (vnorm2 == 0.0) continue;
        for (int k = c; k < d; ++k) {
            double dot = 0.0;
            for (int i = c; i < n_obs; ++i) dot += v[i] * a[(size_t)i * d + k];
            const double f = 2.0 * dot / vnorm2;
            for (int i = c; i < n_obs; ++i) a[(size_t)i * d + k] -= f * v[i];
        }
    }
    out.assign((size_t)dmax * (dmax + 1) / 2, 0.0f);
    for (int i = 0; i < r; ++i)
        for (int k = i; k < d; ++k) out[(size_t)i * dmax - (size_t)i * (i - 1) / 2 + (k - i)] = (float)a[(size_t)i * d + k];
}

extern "C" int dibs_set_data(dibs_plan* p, const float* x, const int32_t* mask, int32_t n_obs,
                             const float* bge_mean_obs_host, void* stream_) {
    if (!p || !x || n_obs < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_set_data: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int d = p->d;
    for (int i = 0; i < 2; ++i) if (p->gexec[i]) { cudaGraphExecDestroy(p->gexec[i]); p->gexec[i] = nullptr; }
    if (p->x) { cudaFree(p->x); p->x = nullptr; }
    if (p->mask) { cudaFree(p->mask); p->mask = nullptr; }
    p->N = n_obs;
    CU(cudaMalloc((void**)&p->x, (size_t)n_obs * d * sizeof(float)));
    CU(cudaMemcpyAsync(p->x, x, (size_t)n_obs * d * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    p->has_mask = false;
    if (mask) {
        // an all-zero mask is the observational case (svgd.py:86-87): detect it once so kernels skip masking
        std::vector<int32_t> h((size_t)n_obs * d);
        CU(cudaMemcpyAsync(h.data(), mask, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        for (int32_t v : h) if (v) { p->has_mask = true; break; }
        if (p->has_mask) {
            CU(cudaMalloc((void**)&p->mask, h.size() * sizeof(int32_t)));
            CU(cudaMemcpyAsync(p->mask, mask, h.size() * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
        }
    }
    p->use_qr = false;
    if (qr_eligible(p->cfg.likelihood, p->dmax) && !p->has_mask) {
        std::vector<float> hx((size_t)n_obs * d);
        CU(cudaMemcpyAsync(hx.data(), p->x, hx.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        qr_upper_packed(hx, n_obs, d, p->dmax, p->lin_r);
        p->use_qr = true;
    }
    p->use_dense = false;
    if (p->cfg.likelihood == DIBS_LIK_LINEAR_GAUSSIAN && d > 32 && !p->has_mask && lin_dense_smem(d) <= 227 * 1024) {
        std::vector<float> hx((size_t)n_obs * d), packed;
        CU(cudaMemcpyAsync(hx.data(), p->x, hx.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        qr_upper_packed(hx, n_obs, d, d, packed);      // packed with stride d: row i starts at i*d - i(i-1)/2
        const int ld = dense_ld(d);
        std::vector<float> dense((size_t)2 * ld * ld, 0.0f);
        for (int i = 0; i < d; ++i)
            for (int k = i; k < d; ++k) {
                const float v = packed[(size_t)i * d - (size_t)i * (i - 1) / 2 + (k - i)];
                dense[(size_t)i * ld + k] = v;                          // Rx[i][k]
                dense[(size_t)ld * ld + (size_t)k * ld + i] = v;        // Rx^T[k][i]
            }
        if (p->rx_dense) { cudaFree(p->rx_dense); p->rx_dense = nullptr; }
        CU(cudaMalloc((void**)&p->rx_dense, dense.size() * sizeof(float)));
        CU(cudaMemcpyAsync(p->rx_dense, dense.data(), dense.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));
        p->use_dense = true;
    }
    if (p->cfg.likelihood == DIBS_LIK_BGE) TRY(bge_prepare(p->cfg, d, n_obs, p->x, p->mask, bge_mean_obs_host, &p->bge_r,
                                                       &p->bge_table, &p->bge_coef, &p->bge_r_stride, stream, g_last_error));
    p->has_data = true;
    return DIBS_OK;
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
// pass shape for this plan: the QR kernel needs observational data and -- unless the graphs are supplied by the
// caller (lp_only hook) -- the legacy threefry layout with an even number of samples (two draws per block)
static McShape mc_shape(const dibs_plan* p, int n_local, int S, bool lp_only) {
    if (p->use_dense) return mc_shape_dense(p->d, n_local, S, !lp_only && !p->cfg.prng_partitionable);
    const bool qr = p->use_qr && (lp_only || (!p->cfg.prng_partitionable && (S % 2) == 0));
    const bool bge_soft = p->cfg.likelihood == DIBS_LIK_BGE && p->cfg.grad_estimator_z == DIBS_ESTIMATOR_REPARAM;
    return mc_shape_for(qr, p->cfg.likelihood, p->d, p->cfg.hidden, n_local, S, !lp_only && !p->cfg.prng_partitionable && !bge_soft);
}
static void apply_shape(McParams& q, const McShape& sh) {
    q.n_chunks = sh.chunks; q.s_per_chunk = sh.spc; q.gpb = sh.gpb; q.paired = sh.paired ? 1 : 0;
}

struct Src {                 // where the particles of a launch live
    const float* z; int z_ld;
    const float* theta; int th_ld;
    int n; int m_offset;
    const StepState* st;
    const uint32_t* keys; int t;         // per-particle sub-keys [n][2] (hooks) ...
    const uint32_t* step_keys;           // ... or [n_splits][n][2] written by k_prologue (step loop)
    const float* scores;                 // [n][d*d] raw U V^T written by k_prologue
};

static const uint32_t* pass_keys(const Src& s, int which) {
    return s.step_keys ? s.step_keys + (size_t)which * s.n * 2 : s.keys;
}

static void fill_mc(const dibs_plan* p, const Src& s, McParams& q) {
    memset(&q, 0, sizeof(q));
    q.z = s.z; q.z_ld = s.z_ld; q.theta = s.theta; q.th_ld = s.th_ld; q.scores = s.scores;
    q.n_local = s.n; q.m_offset = s.m_offset; q.n_particles = p->M;
    q.d = p->d; q.k = p->k; q.n_obs = p->N; q.n_samples = p->cfg.n_grad_mc_samples;
    q.x = p->x; q.mask = p->has_mask ? p->mask : nullptr;
    q.st = s.st; q.partitionable = p->cfg.prng_partitionable;
    q.keys_override = s.keys; q.t_override = s.t;
    q.alpha_linear = p->cfg.alpha_linear; q.tau = p->cfg.tau;
    q.s2 = p->s2; q.log2pis2 = p->log2pis2;
    q.mean_edge = p->cfg.mean_edge; q.sig2_edge = p->sig2_edge; q.lognorm_edge = p->lognorm_edge;
    if (p->cfg.likelihood == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN) {
        q.mean_edge = 0.0f; q.sig2_edge = p->sig2_param; q.lognorm_edge = p->lognorm_param;
    }
    q.hidden = p->cfg.hidden; q.hp = nn_hp(p->cfg.hidden); q.activation = p->cfg.activation;
    q.bge_r = p->bge_r; q.bge_r_stride = p->bge_r_stride; q.bge_table = p->bge_table; q.bge_coef = p->bge_coef;
    q.bge_alpha_mu = p->cfg.bge_alpha_mu; q.bge_alpha_lambd = p->cfg.bge_alpha_lambd;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) return fail(DIBS_ERR_UNSUPPORTED, "problem size needs more than 227 KB of shared memory per CTA");
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DIBS_OK;
}

#define DECL_MC(FAM, DM) namespace dibs { int launch_mc_##FAM##_##DM(int, const McParams&, dim3, size_t, cudaStream_t); }
DECL_MC(lingauss, 8) DECL_MC(lingauss, 16) DECL_MC(lingauss, 20) DECL_MC(lingauss, 32) DECL_MC(lingauss, 64) DECL_MC(lingauss, 128)
DECL_MC(nn, 8) DECL_MC(nn, 16) DECL_MC(nn, 20) DECL_MC(nn, 32) DECL_MC(nn, 64) DECL_MC(nn, 128)
DECL_MC(bge, 8) DECL_MC(bge, 16) DECL_MC(bge, 20) DECL_MC(bge, 32) DECL_MC(bge, 64)
#undef DECL_MC
namespace dibs {
int launch_mc_bgesoft_8(int, const McParams&, dim3, size_t, cudaStream_t);
int launch_mc_bgesoft_16(int, const McParams&, dim3, size_t, cudaStream_t);
int launch_mc_bgesoft_20(int, const McParams&, dim3, size_t, cudaStream_t);
int launch_mc_bgesoft_32(int, const McParams&, dim3, size_t, cudaStream_t);
int launch_mc_linqr_8(int, const McParams&, dim3, int, size_t, const float*, cudaStream_t);
int launch_mc_linqr_16(int, const McParams&, dim3, int, size_t, const float*, cudaStream_t);
int launch_mc_linqr_20(int, const McParams&, dim3, int, size_t, const float*, cudaStream_t);
int launch_mc_linqr_32(int, const McParams&, dim3, int, size_t, const float*, cudaStream_t);
}

// the register-tiled MC kernels are instantiated per DMAX in mc_*_inst.cu (parallel build)
template <int MODE>
static int launch_mc(const dibs_plan* p, McParams q, const McShape& sh, cudaStream_t stream) {
    apply_shape(q, sh);
    dim3 grid(q.n_local, q.n_chunks);
    const int lik = p->cfg.likelihood;
    size_t smem = 0;
    int e = 0;
    // a fused launch (step loop) also needs room for the assemble step its last CTA per particle runs
    const size_t fuse_smem = q.fuse.arrive ? assemble_smem(p->d, p->k, q.fuse.a.z_chunks, q.fuse.a.th_chunks) : 0;
    if (sh.dense) {
        q.paired = 0;
        smem = std::max(lin_dense_smem(p->d), fuse_smem);
        auto kern = k_mc_lin_dense<MODE>;
        TRY(set_smem(kern, smem));
        kern<<<grid, sh.threads, smem, stream>>>(q, p->rx_dense, dense_ld(p->d), dense_nt(p->d));
        LAUNCHED();
        return DIBS_OK;
    }
    if (sh.qr) {
        smem = std::max(mc_lin_qr_smem(p->dmax, q.gpb), fuse_smem);
        if (smem > 227 * 1024) return fail(DIBS_ERR_UNSUPPORTED, "problem size needs more than 227 KB of shared memory per CTA");
        switch (p->dmax) {
            case 8: e = launch_mc_linqr_8(MODE, q, grid, sh.threads, smem, p->lin_r.data(), stream); break;
            case 16: e = launch_mc_linqr_16(MODE, q, grid, sh.threads, smem, p->lin_r.data(), stream); break;
            case 20: e = launch_mc_linqr_20(MODE, q, grid, sh.threads, smem, p->lin_r.data(), stream); break;
            default: e = launch_mc_linqr_32(MODE, q, grid, sh.threads, smem, p->lin_r.data(), stream); break;
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != 0) return fail(DIBS_ERR_CUDA, std::string("MC kernel launch: ") + cudaGetErrorString((cudaError_t)e));
        return DIBS_OK;
    }
    if (lik == DIBS_LIK_BGE && p->cfg.grad_estimator_z == DIBS_ESTIMATOR_REPARAM) {
        // soft-graph BGe: one kernel for the reparameterisation pass and for log-probs of supplied (soft or hard) graphs
        if (MODE != MC_Z_REPARAM && MODE != MC_LP_ONLY) return fail(DIBS_ERR_STATE, "BGe + reparam plan asked for a hard-graph pass");
        q.paired = 0;
        smem = std::max(mc_bge_soft_smem(p->d, p->dmax, p->bge_r_stride == 0), fuse_smem);
        if (smem > 227 * 1024) return fail(DIBS_ERR_UNSUPPORTED, "problem size needs more than 227 KB of shared memory per CTA");
        switch (p->dmax) {
            case 8: e = launch_mc_bgesoft_8(MODE, q, grid, smem, stream); break;
            case 16: e = launch_mc_bgesoft_16(MODE, q, grid, smem, stream); break;
            case 20: e = launch_mc_bgesoft_20(MODE, q, grid, smem, stream); break;
            default: e = launch_mc_bgesoft_32(MODE, q, grid, smem, stream); break;
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e != 0) return fail(DIBS_ERR_CUDA, std::string("MC kernel launch: ") + cudaGetErrorString((cudaError_t)e));
        return DIBS_OK;
    }
    if (lik == DIBS_LIK_LINEAR_GAUSSIAN) {
        smem = mc_lingauss_smem(p->d, p->k, p->N, q.gpb, p->dmax, q.mask != nullptr);
    } else if (lik == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN) {
        smem = mc_nn_smem(p->d, p->N, p->dmax, q.hidden, q.mask != nullptr);
    } else {
        if (MODE != MC_Z_SCORE && MODE != MC_LP_ONLY) return fail(DIBS_ERR_UNSUPPORTED, "BGe supports the score estimator only");
        smem = mc_bge_smem(p->d, p->dmax, q.s_per_chunk * (q.paired ? 2 : 1), p->bge_r_stride == 0);
    }
    smem = std::max(smem, fuse_smem);
    if (smem > 227 * 1024) return fail(DIBS_ERR_UNSUPPORTED, "problem size needs more than 227 KB of shared memory per CTA");
#define GO(FAM) switch (p->dmax) { case 8: e = launch_mc_##FAM##_8(MODE, q, grid, smem, stream); break; \
        case 16: e = launch_mc_##FAM##_16(MODE, q, grid, smem, stream); break; \
        case 20: e = launch_mc_##FAM##_20(MODE, q, grid, smem, stream); break; \
        case 32: e = launch_mc_##FAM##_32(MODE, q, grid, smem, stream); break; \
        case 64: e = launch_mc_##FAM##_64(MODE, q, grid, smem, stream); break; \
        default: e = launch_mc_##FAM##_LAST(MODE, q, grid, smem, stream); break; }
#define launch_mc_lingauss_LAST launch_mc_lingauss_128
#define launch_mc_nn_LAST launch_mc_nn_128
#define launch_mc_bge_LAST launch_mc_bge_64
    if (lik == DIBS_LIK_LINEAR_GAUSSIAN) GO(lingauss)
    else if (lik == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN) GO(nn)
    else GO(bge)
#undef GO
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != 0) return fail(DIBS_ERR_CUDA, std::string("MC kernel launch: ") + cudaGetErrorString((cudaError_t)e));
    return DIBS_OK;
}

static int launch_acyc(const dibs_plan* p, const Src& s, int which_split, float* ds_out, cudaStream_t stream,
                       const FuseAsm* fuse = nullptr) {
    AcycParams a;
    memset(&a, 0, sizeof(a));
    a.z = s.z; a.z_ld = s.z_ld; a.scores = s.scores; a.n_local = s.n; a.m_offset = s.m_offset; a.n_particles = p->M;
    a.d = p->d; a.k = p->k; a.n_samples = p->cfg.n_acyclicity_mc_samples;
    a.st = s.st; a.which_split = which_split; a.partitionable = p->cfg.prng_partitionable;
    a.keys_override = pass_keys(s, which_split); a.t_override = s.t;
    a.alpha_linear = p->cfg.alpha_linear; a.tau = p->cfg.tau; a.ds_out = ds_out;
    if (fuse) a.fuse = *fuse;
    const size_t fuse_smem = fuse ? assemble_smem(p->d, p->k, fuse->a.z_chunks, fuse->a.th_chunks) : 0;
    const int d = p->d;
    if (acyc_rows_path(p)) {
        // row-per-thread kernel: 4 sample pairs per CTA (both lanes of each threefry block are used)
        const size_t smem = std::max(acyclic_rows_smem(d, p->dmax), fuse_smem);
        dim3 grid(s.n, acyc_chunks(p));
#define ACYC_GO(DM) { TRY(set_smem(k_acyclic_rows<DM>, smem)); k_acyclic_rows<DM><<<grid, acyclic_rows_threads(d), smem, stream>>>(a); }
        switch (p->dmax) {
            case 8: ACYC_GO(8) break;
            case 16: ACYC_GO(16) break;
            case 20: ACYC_GO(20) break;
            default: ACYC_GO(32) break;
        }
#undef ACYC_GO
    } else if (d <= 32) {
        int warps = a.n_samples < 8 ? a.n_samples : 8;
        size_t smem = acyclic_smem(d, p->k, warps);
        while (smem > 200 * 1024 && warps > 1) { warps /= 2; smem = acyclic_smem(d, p->k, warps); }
        smem = std::max(smem, fuse_smem);
        TRY(set_smem(k_acyclic_grad<true>, smem));
        k_acyclic_grad<true><<<s.n, warps * 32, smem, stream>>>(a);
    } else if (d <= 64) {
        // 4 x 4 register tiles: one sample per CTA at a time, several CTAs per SM
        const AcycDenseShape sh = acyc_dense4_shape(d, a.n_samples);
        const size_t smem = std::max(sh.smem, fuse_smem);
        TRY(set_smem(k_acyclic_dense4, smem));
        k_acyclic_dense4<<<dim3(s.n, sh.chunks), sh.threads, smem, stream>>>(a, sh.ld, sh.nt, sh.rounds);
    } else {
        // register-tiled matrix powers on shared-memory operands (kernels_dense.cuh)
        const AcycDenseShape sh = acyc_dense_shape(d, a.n_samples);
        const size_t smem = std::max(sh.smem, fuse_smem);
        TRY(set_smem(k_acyclic_dense, smem));
        k_acyclic_dense<<<dim3(s.n, sh.chunks), sh.threads, smem, stream>>>(a, sh.ld, sh.nt, sh.ng, sh.rounds);
    }
    LAUNCHED();
    return DIBS_OK;
}

static void fill_asm(const dibs_plan* p, const Src& s, AsmParams& a) {
    memset(&a, 0, sizeof(a));
    a.z = s.z; a.z_ld = s.z_ld; a.scores = s.scores; a.n_local = s.n; a.d = p->d; a.k = p->k;
    a.st = s.st; a.t_override = s.t;
    a.alpha_linear = p->cfg.alpha_linear; a.beta_linear = p->cfg.beta_linear;
    a.n_samples = p->cfg.n_grad_mc_samples; a.sf_coef = p->cfg.score_function_baseline;
    a.n_acyc = p->cfg.n_acyclicity_mc_samples;
    a.prior_kind = p->cfg.graph_prior; a.er_coef = p->er_coef; a.sigma_z2 = p->sigma_z2;
    a.z_mode = p->cfg.grad_estimator_z == DIBS_ESTIMATOR_SCORE ? MC_Z_SCORE : MC_Z_REPARAM;
}

static int launch_asm(const dibs_plan* p, const AsmParams& a, cudaStream_t stream) {
    size_t smem = assemble_smem(p->d, p->k, a.z_chunks, a.th_chunks);
    TRY(set_smem(k_assemble_grad, smem));
    k_assemble_grad<<<a.n_local, 256, smem, stream>>>(a);
    LAUNCHED();
    return DIBS_OK;
}

// scores + per-pass sub-keys for `n` particles (see k_prologue)
static int launch_prologue(dibs_plan* p, const float* z, int z_ld, int n, int m_offset, const StepState* st,
                           const uint32_t* keys_in, int n_splits, uint32_t pre_split_mask, float* scores,
                           uint32_t* keys_out, cudaStream_t stream) {
    PrologueParams q;
    memset(&q, 0, sizeof(q));
    q.z = z; q.z_ld = z_ld; q.n_local = n; q.m_offset = m_offset; q.n_particles = p->M; q.d = p->d; q.k = p->k;
    q.st = st; q.keys_in = keys_in; q.n_splits = n_splits; q.pre_split_mask = pre_split_mask;
    q.partitionable = p->cfg.prng_partitionable; q.scores = scores; q.keys_out = keys_out;
    size_t smem = (size_t)2 * p->d * p->k * sizeof(float);
    TRY(set_smem(k_prologue, smem));
    k_prologue<<<n, 128, smem, stream>>>(q);
    LAUNCHED();
    return DIBS_OK;
}

// make `to` wait for everything enqueued so far on `from` (a graph edge under stream capture)
static int stream_edge(cudaStream_t from, cudaStream_t to, cudaEvent_t ev) {
    CU(cudaEventRecord(ev, from));
    CU(cudaStreamWaitEvent(to, ev, 0));
    return DIBS_OK;
}

static int ensure_aux(dibs_plan* p) {
    for (int i = 0; i < 3; ++i) {
        if (!p->aux[i]) CU(cudaStreamCreateWithFlags(&p->aux[i], cudaStreamNonBlocking));
        if (!p->ev_join[i]) CU(cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) if (!p->ev_fork[i]) CU(cudaEventCreateWithFlags(&p->ev_fork[i], cudaEventDisableTiming));
    return DIBS_OK;
}

// gradient phase for `s.n` particles: MC passes | acyclicity (independent: sibling streams when `conc`); the assemble
// step of a particle runs inside whichever of its gradient CTAs finishes last (fuse_arrive, kernels_prior.cuh)
static int enqueue_grads(dibs_plan* p, const Src& s, float* th_acc, float* th_stats, float* z_acc, float* z_stats,
                         float* acyc, const float* base_in, float* base_out, float* grad_z, int gz_ld, float* grad_th,
                         int gth_ld, cudaStream_t stream, bool conc, uint32_t* next_keys, StepState* st_next,
                         const PeerPush* push = nullptr) {
    const bool joint = p->cfg.joint;
    const McShape sh = mc_shape(p, s.n, p->cfg.n_grad_mc_samples, false);
    FuseAsm fuse;
    memset(&fuse, 0, sizeof(fuse));
    AsmParams& a = fuse.a;
    fill_asm(p, s, a);
    a.zacc = z_acc; a.zstats = z_stats; a.z_chunks = sh.chunks;
    a.baselines_in = base_in; a.baselines_out = base_out;
    if (joint) { a.thacc = th_acc; a.thstats = th_stats; a.th_chunks = sh.chunks; a.th_dim = p->Dth; }
    a.acyc = acyc; a.acyc_chunks = acyc_chunks(p);
    a.grad_z = grad_z; a.gz_ld = gz_ld; a.grad_th = grad_th; a.gth_ld = gth_ld;
    a.next_keys = next_keys; a.st_next = st_next;
    a.n_step_splits = joint ? 3 : 2; a.n_particles = p->M; a.partitionable = p->cfg.prng_partitionable;
    a.m_offset = s.m_offset; a.pre_split_mask = joint ? 2u : 1u;
    if (push) a.push = *push;
    fuse.arrive = p->arrive;
    fuse.total = (joint ? sh.chunks : 0) + sh.chunks + a.acyc_chunks;

    McParams q;
    cudaStream_t s_th = conc ? p->aux[0] : stream, s_ac = conc ? p->aux[1] : stream;
    if (conc) {
        CU(cudaEventRecord(p->ev_fork[1], stream));
        if (joint && s_th != stream) CU(cudaStreamWaitEvent(s_th, p->ev_fork[1], 0));
        if (s_ac != stream) CU(cudaStreamWaitEvent(s_ac, p->ev_fork[1], 0));
    }
    if (joint) {
        fill_mc(p, s, q);
        q.which_split = 0; q.keys_override = pass_keys(s, 0);
        q.part_acc = th_acc; q.acc_size = p->th_acc_size; q.part_stats = th_stats;
        q.fuse = fuse;
        TRY(launch_mc<MC_THETA_HARD>(p, q, sh, s_th));
        mark(p, s_th, DIBS_PHASE_MC_THETA);
    }
    fill_mc(p, s, q);
    q.which_split = joint ? 1 : 0; q.keys_override = pass_keys(s, q.which_split);
    q.part_acc = z_acc; q.acc_size = p->d * p->d; q.part_stats = z_stats;
    q.fuse = fuse;
    if (p->cfg.grad_estimator_z == DIBS_ESTIMATOR_SCORE) TRY(launch_mc<MC_Z_SCORE>(p, q, sh, stream));
    else TRY(launch_mc<MC_Z_REPARAM>(p, q, sh, stream));
    mark(p, stream, DIBS_PHASE_MC_Z);
    TRY(launch_acyc(p, s, joint ? 2 : 1, acyc, s_ac, &fuse));
    mark(p, s_ac, DIBS_PHASE_ACYCLIC);
    if (conc) {
        if (joint && s_th != stream) TRY(stream_edge(s_th, stream, p->ev_join[0]));
        if (s_ac != stream) TRY(stream_edge(s_ac, stream, p->ev_join[1]));
    }
    return DIBS_OK;
}

static void fill_pair(const dibs_plan* p, PairParams& q) {
    memset(&q, 0, sizeof(q));
    q.n_all = p->M; q.row0 = p->row0; q.n_rows = p->M_loc; q.dz = p->Dz; q.dth = p->Dth;
    q.n_split = p->n_split; q.n_split_z = p->n_split_z; q.split_len_z = p->split_len_z; q.split_len_t = p->split_len_t;
    q.dist_part = p->dist_part;
    q.kz = p->kz; q.kt = p->Dth ? p->kt : nullptr; q.kfull = p->kfull;
    q.k_split = p->phi_mma ? p->k_split : nullptr;
    q.h_z = p->cfg.h_latent; q.h_t = p->cfg.h_theta; q.scale_z = p->cfg.scale_latent; q.scale_t = p->cfg.scale_theta;
    q.n_jsplit = p->n_jsplit; q.j_len = p->j_len; q.phi_part = p->phi_part; q.phi_cnt = p->phi_cnt;
    q.optimizer = p->cfg.optimizer; q.stepsize = p->cfg.stepsize;
}

// kernel matrix of the rank's rows against all particles: squared distances per feature split, then exp -> K
static int launch_kmat(dibs_plan* p, const PairParams& q, cudaStream_t stream) {
    dim3 g1(ceil_div(q.n_all, KT), ceil_div(q.n_rows, KT), q.n_split);
    k_pair_dist<<<g1, 256, 0, stream>>>(q);
    LAUNCHED();
    mark(p, stream, DIBS_PHASE_PAIR_DIST);
    size_t plane = (size_t)q.n_rows * q.n_all;
    int blocks = (int)((plane + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    k_pair_finish<<<blocks, 256, 0, stream>>>(q);
    LAUNCHED();
    mark(p, stream, DIBS_PHASE_PAIR_KERNEL);
    return DIBS_OK;
}

// phi partial sums per j slice, finished (mean, optimizer step, peer push) by each tile's last slice CTA
static int launch_phi(dibs_plan* p, const PairParams& q, cudaStream_t stream, const PhiMmaMaps* mma = nullptr) {
    if (mma) {
        // tensor-core path: 128-row x 64-column tiles per j slice, operands by TMA, 3 x TF32 into TMEM
        static bool attr_set = false;
        if (!attr_set) { CU(cudaFuncSetAttribute(k_phi_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM_BYTES)); attr_set = true; }
        dim3 gm(ceil_div(q.dz, MM_COLS) + ceil_div(q.dth, MM_COLS), ceil_div(q.n_rows, MM_ROWS), q.n_jsplit);
        k_phi_mma<<<gm, MM_THREADS, MM_SMEM_BYTES, stream>>>(q, *mma);
        LAUNCHED();
        mark(p, stream, DIBS_PHASE_PHI_UPDATE);
        return DIBS_OK;
    }
    // 64-row tiles once a rank owns enough rows to fill the GPU with them, 32-row tiles below (twice the CTAs)
    const int cols = ceil_div(q.dz, PT_C) + ceil_div(q.dth, PT_C);
    if (q.n_rows >= 512) k_phi<8><<<dim3(cols, ceil_div(q.n_rows, 64), q.n_jsplit), 128, 0, stream>>>(q);
    else k_phi<4><<<dim3(cols, ceil_div(q.n_rows, 32), q.n_jsplit), 128, 0, stream>>>(q);
    LAUNCHED();
    mark(p, stream, DIBS_PHASE_PHI_UPDATE);
    return DIBS_OK;
}

static int launch_edge_probs(dibs_plan* p, const float* z, int z_ld, int n, float alpha, float* p_out, int32_t* g_out, int raw,
                             cudaStream_t stream) {
    const int warps = edge_probs_warps(p->d, p->k);
    const size_t smem = warps * edge_probs_smem_per_warp(p->d, p->k);
    TRY(set_smem(k_edge_probs, smem));
    int blocks = ceil_div(n, warps);
    const int cap = 148 * 6;                      // grid-stride beyond ~6 CTAs per SM
    if (blocks > cap) blocks = cap;
    k_edge_probs<<<blocks, warps * 32, smem, stream>>>(z, z_ld, n, p->d, p->k, alpha, p_out, g_out, raw);
    LAUNCHED();
    return DIBS_OK;
}

// one full _svgd_step; reads the particles pk[cur], writes the updated local rows into pk[cur^1].
// The raw scores / sub-keys of the step were produced by the previous step's k_opt_update (or by k_prologue
// before the first step of a call); the loop state is read from st[cur] and carried into st[cur^1].
// `conc`: independent passes go to sibling streams (branches of the captured graph); the kernel matrix depends
// on the particles only, so on a single GPU it overlaps the whole gradient phase.
static void fill_push(const dibs_plan* p, PeerPush& pp, int kind, float* const* dst) {
    memset(&pp, 0, sizeof(pp));
    pp.world = p->cfg.world_size; pp.rank = p->cfg.rank; pp.kind = kind;
    for (int q = 0; q < pp.world; ++q) { pp.dst[q] = dst ? dst[q] : nullptr; pp.flags[q] = p->peer_flags[q]; }
    pp.epoch = p->peer_epoch; pp.counter = p->peer_counter;
}

// push this rank's rows of `local` (and/or just raise the flag of `kind`) to every peer
static int launch_push(dibs_plan* p, int kind, const float* local, float* const* peer_bufs, cudaStream_t stream) {
    PeerPush pp;
    fill_push(p, pp, kind, peer_bufs);
    const size_t off4 = local ? (size_t)p->row0 * p->ld / 4 : 0, n4 = local ? (size_t)p->M_loc * p->ld / 4 : 0;
    int blocks = (int)((n4 + 1023) / 1024);
    if (blocks < 1) blocks = 1;
    if (blocks > 2 * 148) blocks = 2 * 148;
    k_peer_push<<<blocks, 256, 0, stream>>>(pp, reinterpret_cast<const float4*>(local), off4, n4);
    LAUNCHED();
    return DIBS_OK;
}

static PeerWait make_wait(const dibs_plan* p, int kind) {
    PeerWait w;
    w.flags = p->p2p ? p->peer_flags_local : nullptr; w.epoch = p->peer_epoch; w.world = p->cfg.world_size; w.kind = kind;
    w.error = p->peer_error_dev; w.timeout_ns = p->peer_timeout_ns;
    return w;
}

static int enqueue_step(dibs_plan* p, int cur, cudaStream_t stream, bool conc) {
    float* P = p->pk[cur];
    float* Pn = p->pk[cur ^ 1];
    float* G = p->gk[cur];
    float* loc = P + (size_t)p->row0 * p->ld;
    float* gloc = G + (size_t)p->row0 * p->ld;
    StepState* st = p->st + cur;
    const bool multi = p->cfg.world_size > 1;
    if (multi && !p->p2p && (!p->comm || !p->comm_x))
        return fail(DIBS_ERR_STATE, "world_size > 1 but neither peer memory nor NCCL communicators are attached");
    if (conc) TRY(ensure_aux(p));
    PairParams q;
    fill_pair(p, q);
    q.x_all = P; q.ld = p->ld; q.g_all = G; q.g_ld = p->ld;
    q.wait_x = make_wait(p, PEER_KIND_X); q.wait_g = make_wait(p, PEER_KIND_GRAD);
    // the kernel matrix needs the particles only: it runs on a side branch under the gradient phase.  On several GPUs
    // the rows every rank updated at the end of the previous step were stored into this rank's buffer by the peers'
    // phi kernels (the distance kernel waits on their flags) or, on the NCCL path, arrive by an all-gather here
    cudaStream_t s_k = conc ? p->aux[2] : stream;
    if (conc && s_k != stream) TRY(stream_edge(stream, s_k, p->ev_fork[0]));
    if (multi && !p->p2p) {
        NC(g_nccl.AllGather(loc, P, (size_t)p->M_loc * p->ld, /*ncclFloat32*/ 7, p->comm_x, s_k));
        mark(p, s_k, DIBS_PHASE_ALLGATHER);
    }
    TRY(launch_kmat(p, q, s_k));
    Src s{loc, p->ld, p->Dth ? loc + p->Dz : nullptr, p->ld, p->M_loc, p->row0, st, nullptr, 0, p->step_keys, p->scores};
    // peer-memory path: the producing kernels store their rows into every peer themselves (fused exchange)
    const bool fuse = multi && p->p2p;
    PeerPush push_g;
    if (fuse) fill_push(p, push_g, PEER_KIND_GRAD, p->peer_gk[cur]);
    TRY(enqueue_grads(p, s, p->th_acc, p->th_stats, p->z_acc, p->z_stats, p->acyc, p->base, p->base,
                      gloc, p->ld, p->Dth ? gloc + p->Dz : nullptr, p->ld, stream, conc, p->step_keys,
                      p->st + (cur ^ 1), fuse ? &push_g : nullptr));
    if (multi && !fuse) {
        // the one exchange on the critical path: every rank contributes its gradient rows [dZ | dTheta]
        NC(g_nccl.AllGather(gloc, G, (size_t)p->M_loc * p->ld, /*ncclFloat32*/ 7, p->comm, stream));
        mark(p, stream, DIBS_PHASE_ALLGATHER);
    }
    if (conc && s_k != stream) TRY(stream_edge(s_k, stream, p->ev_join[2]));
    q.x_next = Pn + (size_t)p->row0 * p->ld; q.next_ld = p->ld;
    q.v = p->v; q.v_ld = p->D;
    if (fuse) fill_push(p, q.push_x, PEER_KIND_X, p->peer_pk[cur ^ 1]);
    if (p->tl_capture >= 0) mark(p, stream, DIBS_PHASE_ASSEMBLE);     // timeline diagnostics: the join point before phi
    TRY(launch_phi(p, q, stream, p->phi_mma ? &p->mma_maps[cur] : nullptr));
    // raw scores U V^T of the NEXT step from the updated latent rows (edge-probability pass, dibs.py:179-181)
    TRY(launch_edge_probs(p, q.x_next, p->ld, p->M_loc, 0.0f, p->scores, nullptr, 1, stream));
    mark(p, stream, DIBS_PHASE_STEP_KEYS);
    return DIBS_OK;
}

__global__ void k_set_state(StepState* st, const uint32_t* key, int t) {
    st->key[0] = key[0]; st->key[1] = key[1]; st->t = t; st->pad = 0;
}
__global__ void k_get_key(const StepState* st, uint32_t* key) { key[0] = st->key[0]; key[1] = st->key[1]; }
// per-kernel timing mode: hold the stream for `ns` so that the host has enqueued the whole eager step before its first
// kernel starts -- the event-to-event times are then kernel durations, not host launch gaps
__global__ void k_delay(unsigned long long ns) {
    const unsigned long long t0 = global_timer_ns();
    while (global_timer_ns() - t0 < ns) __nanosleep(200);
}

static int svgd_steps_impl(dibs_plan* p, int32_t t_start, int32_t n_steps, float* z, float* theta, float* v_z,
                           float* v_theta, uint32_t* key, float* sf_baseline, void* stream_, bool timed, int per_kernel,
                           void* flush_buf, size_t flush_bytes, float* step_ms, float* phase_ms) {
    if (!p || !z || !key || !sf_baseline) return fail(DIBS_ERR_INVALID_ARG, "dibs_svgd_steps: null argument");
    if (!p->has_data) return fail(DIBS_ERR_STATE, "dibs_set_data has not been called");
    if (p->Dth && !theta) return fail(DIBS_ERR_INVALID_ARG, "theta is required for joint inference");
    const bool rms = p->cfg.optimizer == DIBS_OPT_RMSPROP;
    if (rms && (!v_z || (p->Dth && !v_theta))) return fail(DIBS_ERR_INVALID_ARG, "rmsprop needs v_z / v_theta");
    if (n_steps <= 0) return DIBS_OK;
    TRY(dibs_plan_status(p));
    TRY(ensure_step_ws(p));
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t fz = sizeof(float) * p->Dz, ft = sizeof(float) * p->Dth, fl = sizeof(float) * p->ld, fd = sizeof(float) * p->D;
    float* loc0 = p->pk[0] + (size_t)p->row0 * p->ld;
    // pack caller-owned arrays into the plan's row layout
    CU(cudaMemcpy2DAsync(loc0, fl, z, fz, fz, p->M_loc, cudaMemcpyDeviceToDevice, stream));
    if (p->Dth) CU(cudaMemcpy2DAsync(loc0 + p->Dz, fl, theta, ft, ft, p->M_loc, cudaMemcpyDeviceToDevice, stream));
    if (rms) {
        CU(cudaMemcpy2DAsync(p->v, fd, v_z, fz, fz, p->M_loc, cudaMemcpyDeviceToDevice, stream));
        if (p->Dth) CU(cudaMemcpy2DAsync(p->v + p->Dz, fd, v_theta, ft, ft, p->M_loc, cudaMemcpyDeviceToDevice, stream));
    }
    CU(cudaMemcpyAsync(p->base, sf_baseline, sizeof(float) * p->M_loc, cudaMemcpyDeviceToDevice, stream));
    if (p->p2p) {
        // call-level barrier: no peer may still be reading the rows this call is about to overwrite
        k_peer_wait<<<1, 32, 0, stream>>>(make_wait(p, PEER_KIND_DONE));
        LAUNCHED();
    }
    k_set_state<<<1, 1, 0, stream>>>(p->st, key, t_start);
    LAUNCHED();
    // scores and sub-keys of the first step; later steps get theirs from the previous step's k_opt_update
    TRY(launch_prologue(p, loc0, p->ld, p->M_loc, p->row0, p->st, nullptr, p->cfg.joint ? 3 : 2, p->cfg.joint ? 2u : 1u,
                        p->scores, p->step_keys, stream));

    // peer-memory path: publish the rows this call just packed (later steps: the phi kernel pushes its updated rows)
    if (p->p2p) TRY(launch_push(p, PEER_KIND_X, p->pk[0], p->peer_pk[0], stream));
    const bool timeline = timed && per_kernel == 2;
    if (timeline) per_kernel = 0;
    const bool graph = p->use_graph && !(timed && per_kernel);
    if (timeline && !p->tl_exec[0]) {
        for (int par = 0; par < 2; ++par) {
            cudaGraph_t g = nullptr;
            if (!p->cap_stream) CU(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
            long long before = g_launches.load();
            CU(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
            p->tl_capture = par;
            mark(p, p->cap_stream, -1);
            int r = enqueue_step(p, par, p->cap_stream, p->concurrent);
            p->tl_capture = -1;
            cudaError_t e = cudaStreamEndCapture(p->cap_stream, &g);
            g_launches.store(before);
            if (r != DIBS_OK) { if (g) cudaGraphDestroy(g); return r; }
            if (e != cudaSuccess) return fail(DIBS_ERR_CUDA, std::string("cudaStreamEndCapture (timeline): ") + cudaGetErrorString(e));
            CU(cudaGraphInstantiate(&p->tl_exec[par], g, 0));
            cudaGraphDestroy(g);
        }
    }
    if (timeline && phase_ms) for (int i = 0; i < DIBS_N_PHASES; ++i) phase_ms[i] = 0.0f;
    p->ev_used = 0;
    if (graph && !p->gexec[0]) {
        if (p->cfg.world_size > 1 && !p->p2p) {
            // warm NCCL up outside the capture (first-use allocations and connection set-up are not capturable);
            // pk[1] holds nothing yet, so an in-place all-gather on it is harmless
            if (!p->comm || !p->comm_x) return fail(DIBS_ERR_STATE, "world_size > 1 but no NCCL communicators attached");
            NC(g_nccl.AllGather(p->pk[1] + (size_t)p->row0 * p->ld, p->pk[1], (size_t)p->M_loc * p->ld, 7, p->comm_x, stream));
            NC(g_nccl.AllGather(p->gk[1] + (size_t)p->row0 * p->ld, p->gk[1], (size_t)p->M_loc * p->ld, 7, p->comm, stream));
            CU(cudaStreamSynchronize(stream));
        }
        for (int par = 0; par < 2; ++par) {
            long long before = g_launches.load();
            cudaGraph_t g = nullptr;
            // capture on a plan-owned stream (the caller's may be the legacy default stream, which cannot capture)
            if (!p->cap_stream) CU(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
            CU(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeThreadLocal));
            int r = enqueue_step(p, par, p->cap_stream, p->concurrent);
            cudaError_t e = cudaStreamEndCapture(p->cap_stream, &g);
            if (r != DIBS_OK) { if (g) cudaGraphDestroy(g); return r; }
            if (e != cudaSuccess) return fail(DIBS_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            CU(cudaGraphInstantiate(&p->gexec[par], g, 0));
            cudaGraphDestroy(g);
            p->kernels_per_step = (int)(g_launches.load() - before);
            g_launches.store(before);   // captured, not executed
        }
    }
    for (int i = 0; i < n_steps; ++i) {
        int cur = i & 1;
        if (timed) {
            // cold-L2 protocol: overwrite a buffer larger than L2 (untimed), then bracket the step with events
            if (flush_buf) CU(cudaMemsetAsync(flush_buf, i & 0xff, flush_bytes, stream));
            if (flush_buf && p->cfg.world_size > 1) {
                // the flush takes a different time on every rank: line the ranks up again before the bracket opens,
                // otherwise a step is charged the peers' flush while it waits for their rows
                if (p->p2p) {
                    TRY(launch_push(p, PEER_KIND_SYNC, nullptr, nullptr, stream));
                    k_peer_wait<<<1, 32, 0, stream>>>(make_wait(p, PEER_KIND_SYNC));
                    LAUNCHED();
                } else {
                    NC(g_nccl.AllGather(p->peer_epoch, p->peer_flags_local, 1, 7, p->comm, stream));
                }
            }
            if (per_kernel) { k_delay<<<1, 1, 0, stream>>>(300000ull); LAUNCHED(); }
            p->timing = true;
            mark(p, stream, -1);
            p->timing = per_kernel != 0;
        }
        if (timeline) {
            // one step at a time: the events of a parity graph are re-recorded by its next launch
            CU(cudaGraphLaunch(p->tl_exec[cur], stream));
            g_launches.fetch_add(p->kernels_per_step, std::memory_order_relaxed);
            CU(cudaStreamSynchronize(stream));
            float last = 0.0f;
            for (size_t e = 1; e < p->tl_ev[cur].size(); ++e) {
                float ms = 0.0f;
                CU(cudaEventElapsedTime(&ms, p->tl_ev[cur][0], p->tl_ev[cur][e]));
                const int ph = p->tl_phase[cur][e];
                if (phase_ms && ph >= 0 && ph < DIBS_N_PHASES) phase_ms[ph] += ms;      // END time of the phase's kernel
                if (ms > last) last = ms;
            }
            if (step_ms) step_ms[i] = last;
            continue;
        } else if (graph) {
            CU(cudaGraphLaunch(p->gexec[cur], stream));
            g_launches.fetch_add(p->kernels_per_step, std::memory_order_relaxed);
        } else {
            TRY(enqueue_step(p, cur, stream, p->concurrent && !(timed && per_kernel)));
        }
        if (timed && !per_kernel) { p->timing = true; mark(p, stream, DIBS_N_PHASES); }
        p->timing = false;
    }
    if (timed && !timeline) {
        // device time between consecutive events on the launching stream; a -1 marker starts a step
        if (phase_ms) for (int i = 0; i < DIBS_N_PHASES; ++i) phase_ms[i] = 0.0f;
        CU(cudaStreamSynchronize(stream));
        int step = -1;
        for (size_t i = 0; i < p->ev_used; ++i) {
            if (p->ev_phase[i] < 0) { ++step; if (step_ms && step < n_steps) step_ms[step] = 0.0f; continue; }
            float ms = 0.0f;
            CU(cudaEventElapsedTime(&ms, p->ev_pool[i - 1], p->ev_pool[i]));
            if (phase_ms && p->ev_phase[i] < DIBS_N_PHASES) phase_ms[p->ev_phase[i]] += ms;
            if (step_ms && step >= 0 && step < n_steps) step_ms[step] += ms;
        }
    }
    float* locf = p->pk[n_steps & 1] + (size_t)p->row0 * p->ld;
    CU(cudaMemcpy2DAsync(z, fz, locf, fl, fz, p->M_loc, cudaMemcpyDeviceToDevice, stream));
    if (p->Dth) CU(cudaMemcpy2DAsync(theta, ft, locf + p->Dz, fl, ft, p->M_loc, cudaMemcpyDeviceToDevice, stream));
    if (rms) {
        CU(cudaMemcpy2DAsync(v_z, fz, p->v, fd, fz, p->M_loc, cudaMemcpyDeviceToDevice, stream));
        if (p->Dth) CU(cudaMemcpy2DAsync(v_theta, ft, p->v + p->Dz, fd, ft, p->M_loc, cudaMemcpyDeviceToDevice, stream));
    }
    CU(cudaMemcpyAsync(sf_baseline, p->base, sizeof(float) * p->M_loc, cudaMemcpyDeviceToDevice, stream));
    if (p->p2p) TRY(launch_push(p, PEER_KIND_DONE, nullptr, nullptr, stream));
    k_get_key<<<1, 1, 0, stream>>>(p->st + (n_steps & 1), key);
    LAUNCHED();
    return DIBS_OK;
}

extern "C" int dibs_svgd_steps(dibs_plan* p, int32_t t_start, int32_t n_steps, float* z, float* theta, float* v_z,
                               float* v_theta, uint32_t* key, float* sf_baseline, void* stream) {
    return svgd_steps_impl(p, t_start, n_steps, z, theta, v_z, v_theta, key, sf_baseline, stream, false, 0, nullptr, 0,
                           nullptr, nullptr);
}

extern "C" int dibs_svgd_steps_timed(dibs_plan* p, int32_t t_start, int32_t n_steps, float* z, float* theta, float* v_z,
                                     float* v_theta, uint32_t* key, float* sf_baseline, void* stream,
                                     int32_t per_kernel, void* l2_flush_buf, int64_t l2_flush_bytes,
                                     float* step_ms_host, float* phase_ms_host) {
    if (!step_ms_host && !phase_ms_host) return fail(DIBS_ERR_INVALID_ARG, "dibs_svgd_steps_timed: no output buffer");
    if (phase_ms_host && !per_kernel) return fail(DIBS_ERR_INVALID_ARG, "phase_ms_host needs per_kernel = 1 or 2");
    if (l2_flush_buf && l2_flush_bytes <= 0) return fail(DIBS_ERR_INVALID_ARG, "l2_flush_bytes must be positive");
    return svgd_steps_impl(p, t_start, n_steps, z, theta, v_z, v_theta, key, sf_baseline, stream, true, per_kernel,
                           l2_flush_buf, (size_t)(l2_flush_buf ? l2_flush_bytes : 0), step_ms_host, phase_ms_host);
}

// ------------------------------------------------------------------------------------------
// NCCL attach
// ------------------------------------------------------------------------------------------
extern "C" int dibs_nccl_unique_id(uint8_t* id128) {
    if (!id128) return fail(DIBS_ERR_INVALID_ARG, "null id buffer");
    TRY(nccl_load());
    NcclId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return DIBS_OK;
}

extern "C" int dibs_plan_attach_nccl(dibs_plan* p, const uint8_t* id128) {
    if (!p || !id128) return fail(DIBS_ERR_INVALID_ARG, "null argument");
    TRY(nccl_load());
    NcclId id;
    memcpy(id.internal, id128, 128);
    // first call: communicator of the gradient exchange; second call (another unique id): the particle exchange
    NC(g_nccl.CommInitRank(p->comm ? &p->comm_x : &p->comm, p->cfg.world_size, id, p->cfg.rank));
    return DIBS_OK;
}

// ------------------------------------------------------------------------------------------
// peer-memory exchange set-up (CUDA IPC between the ranks of one node)
// ------------------------------------------------------------------------------------------
static const int IPC_BUFS = 5;   // pk[0], pk[1], gk[0], gk[1], flags

extern "C" int dibs_plan_ipc_export(dibs_plan* p, uint8_t* handles_out) {
    if (!p || !handles_out) return fail(DIBS_ERR_INVALID_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    TRY(ensure_step_ws(p));
    void* bufs[IPC_BUFS] = {p->pk[0], p->pk[1], p->gk[0], p->gk[1], p->peer_flags_local};
    for (int i = 0; i < IPC_BUFS; ++i) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, bufs[i]));
        memcpy(handles_out + 64 * i, &h, 64);
    }
    return DIBS_OK;
}

extern "C" int dibs_plan_ipc_detach(dibs_plan* p);

extern "C" int dibs_plan_ipc_attach(dibs_plan* p, const uint8_t* all_handles) {
    if (!p || !all_handles) return fail(DIBS_ERR_INVALID_ARG, "null argument");
    const int W = p->cfg.world_size;
    if (W < 2 || W > PEER_MAX) return fail(DIBS_ERR_UNSUPPORTED, "peer-memory exchange needs 2..16 ranks");
    for (int q = 0; q < W; ++q) {
        void* opened[IPC_BUFS];
        if (q == p->cfg.rank) {
            void* mine[IPC_BUFS] = {p->pk[0], p->pk[1], p->gk[0], p->gk[1], p->peer_flags_local};
            memcpy(opened, mine, sizeof(mine));
        } else {
            for (int i = 0; i < IPC_BUFS; ++i) {
                cudaIpcMemHandle_t h;
                memcpy(&h, all_handles + ((size_t)q * IPC_BUFS + i) * 64, 64);
                cudaError_t e = cudaIpcOpenMemHandle(&opened[i], h, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    dibs_plan_ipc_detach(p);
                    return fail(DIBS_ERR_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(q) + "): " + cudaGetErrorString(e));
                }
                p->ipc_opened.push_back(opened[i]);
            }
        }
        p->peer_pk[0][q] = (float*)opened[0]; p->peer_pk[1][q] = (float*)opened[1];
        p->peer_gk[0][q] = (float*)opened[2]; p->peer_gk[1][q] = (float*)opened[3];
        p->peer_flags[q] = (uint32_t*)opened[4];
    }
    for (int i = 0; i < 2; ++i) if (p->gexec[i]) { cudaGraphExecDestroy(p->gexec[i]); p->gexec[i] = nullptr; }
    p->p2p = true;
    return DIBS_OK;
}

extern "C" int dibs_plan_ipc_detach(dibs_plan* p) {
    if (!p) return fail(DIBS_ERR_INVALID_ARG, "null argument");
    if (!p->ipc_opened.empty()) {
        cudaDeviceSynchronize();
        for (void* q : p->ipc_opened) cudaIpcCloseMemHandle(q);
        p->ipc_opened.clear();
    }
    for (int q = 0; q < PEER_MAX; ++q) {
        p->peer_pk[0][q] = p->peer_pk[1][q] = p->peer_gk[0][q] = p->peer_gk[1][q] = nullptr;
        p->peer_flags[q] = nullptr;
    }
    for (int i = 0; i < 2; ++i) if (p->gexec[i]) { cudaGraphExecDestroy(p->gexec[i]); p->gexec[i] = nullptr; }
    p->p2p = false;
    return DIBS_OK;
}

// ------------------------------------------------------------------------------------------
// init
// ------------------------------------------------------------------------------------------
extern "C" int dibs_init_particles(dibs_plan* p, const uint32_t* key, float* z_all, float* theta_all, void* stream_) {
    if (!p || !key || !z_all) return fail(DIBS_ERR_INVALID_ARG, "dibs_init_particles: null argument");
    if (p->Dth && !theta_all) return fail(DIBS_ERR_INVALID_ARG, "theta buffer required");
    cudaStream_t stream = (cudaStream_t)stream_;
    InitParams q;
    q.key = key; q.M = p->M; q.d = p->d; q.k = p->k; q.std_z = p->cfg.latent_prior_std; q.z = z_all; q.theta = theta_all;
    q.lik = p->cfg.likelihood; q.hidden = p->cfg.hidden; q.partitionable = p->cfg.prng_partitionable;
    q.mean_edge = p->cfg.mean_edge; q.sig_edge = p->cfg.sig_edge; q.min_edge = p->cfg.min_edge; q.sig_param = p->cfg.sig_param;
    q.dth = p->Dth;
    size_t nz = (size_t)p->M * p->Dz;
    k_init_z<<<(int)((nz + 255) / 256), 256, 0, stream>>>(q);
    LAUNCHED();
    if (p->cfg.likelihood == DIBS_LIK_LINEAR_GAUSSIAN) {
        size_t nt = (size_t)p->M * p->Dth;
        k_init_theta_lin<<<(int)((nt + 255) / 256), 256, 0, stream>>>(q);
        LAUNCHED();
    } else if (p->cfg.likelihood == DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN) {
        k_init_theta_nn<<<ceil_div(p->M * p->d, 4), 128, 0, stream>>>(q);
        LAUNCHED();
    }
    return DIBS_OK;
}

// ------------------------------------------------------------------------------------------
// hooks
// ------------------------------------------------------------------------------------------
struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch() { for (void* q : ptrs) cudaFree(q); }
    template <typename T> int get(T** out, size_t count) {
        void* q = nullptr;
        CU(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
        ptrs.push_back(q);
        *out = (T*)q;
        return DIBS_OK;
    }
};

extern "C" int dibs_edge_probs(dibs_plan* p, const float* z, int32_t n, int32_t t, float* p_out, void* stream_) {
    if (!p || !z || !p_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_edge_probs: bad arguments");
    return launch_edge_probs(p, z, p->Dz, n, p->cfg.alpha_linear * (float)t, p_out, nullptr, 0, (cudaStream_t)stream_);
}

extern "C" int dibs_particle_to_g_lim(dibs_plan* p, const float* z, int32_t n, int32_t* g_out, void* stream_) {
    if (!p || !z || !g_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_particle_to_g_lim: bad arguments");
    return launch_edge_probs(p, z, p->Dz, n, 0.0f, nullptr, g_out, 0, (cudaStream_t)stream_);
}

extern "C" int dibs_sample_graphs(dibs_plan* p, const float* probs, const uint32_t* keys, int32_t n, int32_t n_samples,
                                  int32_t* g_out, void* stream_) {
    if (!p || !probs || !keys || !g_out || n < 1 || n_samples < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_sample_graphs: bad arguments");
    size_t smem = ((size_t)p->d * p->d + 2 * p->d * p->k) * sizeof(float);
    TRY(set_smem(k_sample_graphs, smem));
    k_sample_graphs<<<n, 256, smem, (cudaStream_t)stream_>>>(probs, p->d * p->d, keys, p->d, p->k, n_samples, 1, 0.0f, 0.0f,
                                                            p->cfg.prng_partitionable, g_out, nullptr);
    LAUNCHED();
    return DIBS_OK;
}

extern "C" int dibs_soft_graphs(dibs_plan* p, const float* z, const uint32_t* keys, int32_t n, int32_t n_samples,
                                int32_t t, float* g_out, void* stream_) {
    if (!p || !z || !keys || !g_out || n < 1 || n_samples < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_soft_graphs: bad arguments");
    size_t smem = ((size_t)p->d * p->d + 2 * p->d * p->k) * sizeof(float);
    TRY(set_smem(k_sample_graphs, smem));
    k_sample_graphs<<<n, 256, smem, (cudaStream_t)stream_>>>(z, p->Dz, keys, p->d, p->k, n_samples, 0,
                                                            p->cfg.alpha_linear * (float)t, p->cfg.tau,
                                                            p->cfg.prng_partitionable, nullptr, g_out);
    LAUNCHED();
    return DIBS_OK;
}

extern "C" int dibs_log_joint_prob(dibs_plan* p, const float* g, const float* theta, int32_t n, int32_t n_samples,
                                   float* lp_out, void* stream_) {
    if (!p || !g || !lp_out || n < 1 || n_samples < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_log_joint_prob: bad arguments");
    if (!p->has_data) return fail(DIBS_ERR_STATE, "dibs_set_data has not been called");
    if (p->Dth && !theta) return fail(DIBS_ERR_INVALID_ARG, "theta required");
    cudaStream_t stream = (cudaStream_t)stream_;
    Scratch sc;
    float* zdummy;
    TRY(sc.get(&zdummy, (size_t)n * p->Dz));
    CU(cudaMemsetAsync(zdummy, 0, (size_t)n * p->Dz * sizeof(float), stream));
    Src s{zdummy, p->Dz, theta, p->Dth, n, 0, nullptr, nullptr, 0, nullptr, nullptr};
    McParams q;
    fill_mc(p, s, q);
    q.n_samples = n_samples; q.g_ext = g; q.lp_out = lp_out;
    McShape sh = mc_shape(p, n, n_samples, true);
    sh.chunks = 1; sh.spc = sh.qr ? (n_samples + 1) / 2 : n_samples;     // one CTA per particle, all samples
    sh.paired = false;
    TRY(launch_mc<MC_LP_ONLY>(p, q, sh, stream));
    CU(cudaStreamSynchronize(stream));
    return DIBS_OK;
}

static int hook_grads(dibs_plan* p, const float* z, const float* theta, const float* baselines, int t,
                      const uint32_t* keys, int n, int what /*0 z-lik, 1 theta, 2 prior, 3 constraint only*/,
                      float* grad_out, float* baselines_out, cudaStream_t stream) {
    if (!p->has_data && what < 2) return fail(DIBS_ERR_STATE, "dibs_set_data has not been called");
    Scratch sc;
    float* scores = nullptr; uint32_t* pkeys = nullptr;
    TRY(sc.get(&scores, (size_t)n * p->d * p->d));
    TRY(sc.get(&pkeys, (size_t)n * 2));
    TRY(launch_prologue(p, z, p->Dz, n, 0, nullptr, keys, 1, what == 0 ? 1u : 0u, scores, pkeys, stream));
    Src s{z, p->Dz, theta, p->Dth, n, 0, nullptr, pkeys, t, nullptr, scores};
    const McShape sh = mc_shape(p, n, p->cfg.n_grad_mc_samples, false);
    const int chunks = sh.chunks;
    float *acc = nullptr, *stats = nullptr, *acyc = nullptr, *gz_tmp = nullptr;
    AsmParams a;
    fill_asm(p, s, a);
    McParams q;
    fill_mc(p, s, q);
    if (what == 0) {
        TRY(sc.get(&acc, (size_t)n * chunks * p->d * p->d));
        TRY(sc.get(&stats, (size_t)n * chunks * 4));
        q.part_acc = acc; q.acc_size = p->d * p->d; q.part_stats = stats;
        if (p->cfg.grad_estimator_z == DIBS_ESTIMATOR_SCORE) TRY(launch_mc<MC_Z_SCORE>(p, q, sh, stream));
        else TRY(launch_mc<MC_Z_REPARAM>(p, q, sh, stream));
        a.zacc = acc; a.zstats = stats; a.z_chunks = chunks; a.baselines_in = baselines; a.baselines_out = baselines_out;
        a.grad_z = grad_out; a.gz_ld = p->Dz;
    } else if (what == 1) {
        if (!p->Dth) return fail(DIBS_ERR_INVALID_ARG, "no theta in marginal inference");
        TRY(sc.get(&acc, (size_t)n * chunks * p->Dth));
        TRY(sc.get(&stats, (size_t)n * chunks * 4));
        TRY(sc.get(&gz_tmp, (size_t)n * p->Dz));
        q.part_acc = acc; q.acc_size = p->Dth; q.part_stats = stats;
        TRY(launch_mc<MC_THETA_HARD>(p, q, sh, stream));
        a.thacc = acc; a.thstats = stats; a.th_chunks = chunks; a.th_dim = p->Dth;
        a.grad_z = gz_tmp; a.gz_ld = p->Dz; a.grad_th = grad_out; a.gth_ld = p->Dth;
    } else {
        TRY(sc.get(&acyc, (size_t)n * acyc_chunks(p) * p->d * p->d));
        TRY(launch_acyc(p, s, 0, acyc, stream));
        a.acyc = acyc; a.acyc_chunks = acyc_chunks(p); a.constraint_only = (what == 3);
        a.grad_z = grad_out; a.gz_ld = p->Dz;
    }
    TRY(launch_asm(p, a, stream));
    CU(cudaStreamSynchronize(stream));
    return DIBS_OK;
}

extern "C" int dibs_grad_z_likelihood(dibs_plan* p, const float* z, const float* theta, const float* baselines, int32_t t,
                                      const uint32_t* keys, int32_t n, float* grad_out, float* baselines_out, void* stream) {
    if (!p || !z || !keys || !grad_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_grad_z_likelihood: bad arguments");
    if (p->Dth && !theta) return fail(DIBS_ERR_INVALID_ARG, "theta required");
    return hook_grads(p, z, theta, baselines, t, keys, n, 0, grad_out, baselines_out, (cudaStream_t)stream);
}

extern "C" int dibs_grad_theta_likelihood(dibs_plan* p, const float* z, const float* theta, int32_t t, const uint32_t* keys,
                                          int32_t n, float* grad_out, void* stream) {
    if (!p || !z || !theta || !keys || !grad_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_grad_theta_likelihood: bad arguments");
    return hook_grads(p, z, theta, nullptr, t, keys, n, 1, grad_out, nullptr, (cudaStream_t)stream);
}

extern "C" int dibs_grad_latent_prior(dibs_plan* p, const float* z, int32_t t, const uint32_t* keys, int32_t n,
                                      int32_t constraint_only, float* grad_out, void* stream) {
    if (!p || !z || !keys || !grad_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_grad_latent_prior: bad arguments");
    return hook_grads(p, z, nullptr, nullptr, t, keys, n, constraint_only ? 3 : 2, grad_out, nullptr, (cudaStream_t)stream);
}

extern "C" int dibs_acyclic_constr(dibs_plan* p, const float* g, int32_t n, float* h_out, void* stream_) {
    if (!p || !g || !h_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_acyclic_constr: bad arguments");
    size_t smem = (size_t)4 * p->d * (p->d | 1) * sizeof(float);
    TRY(set_smem(k_acyclic_value, smem));
    k_acyclic_value<<<n, 256, smem, (cudaStream_t)stream_>>>(g, n, p->d, h_out);
    LAUNCHED();
    return DIBS_OK;
}

extern "C" int dibs_particle_summary(dibs_plan* p, const float* z, int32_t n, int32_t t, float* summary_host, void* stream_) {
    if (!p || !z || !summary_host || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_particle_summary: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int d = p->d, dd = d * d;
    const size_t need = (size_t)n * dd + n + 2 + dd;
    if (p->summary_cap < need) {
        // growing the scratch is the only synchronising path (first call / larger particle set)
        if (p->summary_ws) { CU(cudaStreamSynchronize(stream)); CU(cudaFree(p->summary_ws)); p->summary_ws = nullptr; }
        CU(cudaMalloc((void**)&p->summary_ws, need * sizeof(float)));
        p->summary_cap = need;
    }
    float* p_all = p->summary_ws; float* h_all = p_all + (size_t)n * dd; float* rec = h_all + n;
    const size_t smem = ((size_t)4 * d * (d | 1) + 2 * (size_t)d * p->k) * sizeof(float);
    TRY(set_smem(k_particle_summary, smem));
    k_particle_summary<<<n, 256, smem, stream>>>(z, p->Dz, n, d, p->k, p->cfg.alpha_linear * (float)t, p_all, h_all);
    LAUNCHED();
    k_summary_reduce<<<ceil_div(dd, 256), 256, 0, stream>>>(p_all, h_all, n, dd, rec);
    LAUNCHED();
    CU(cudaMemcpyAsync(summary_host, rec, (size_t)(2 + dd) * sizeof(float), cudaMemcpyDeviceToHost, stream));
    return DIBS_OK;
}

static int hook_pair(dibs_plan* p, const float* z, const float* theta, const float* gz, const float* gth, int n,
                     float* k_out, float* phi_z, float* phi_th, cudaStream_t stream) {
    Scratch sc;
    const int D = p->D;
    float *xs, *gs, *dist, *kz, *kt, *kf, *phi = nullptr;
    uint32_t* cnt = nullptr;
    TRY(sc.get(&xs, (size_t)n * D));
    TRY(sc.get(&gs, (size_t)n * D));
    const size_t fz = sizeof(float) * p->Dz, ft = sizeof(float) * p->Dth, fd = sizeof(float) * D;
    CU(cudaMemsetAsync(gs, 0, (size_t)n * fd, stream));
    CU(cudaMemcpy2DAsync(xs, fd, z, fz, fz, n, cudaMemcpyDeviceToDevice, stream));
    if (p->Dth) CU(cudaMemcpy2DAsync(xs + p->Dz, fd, theta, ft, ft, n, cudaMemcpyDeviceToDevice, stream));
    if (gz) CU(cudaMemcpy2DAsync(gs, fd, gz, fz, fz, n, cudaMemcpyDeviceToDevice, stream));
    if (gth && p->Dth) CU(cudaMemcpy2DAsync(gs + p->Dz, fd, gth, ft, ft, n, cudaMemcpyDeviceToDevice, stream));
    PairParams q;
    fill_pair(p, q);
    q.n_all = n; q.row0 = 0; q.n_rows = n;
    size_t plane = (size_t)n * n;
    const size_t n_cnt = (size_t)(ceil_div(p->Dz, PT_C) + ceil_div(p->Dth, PT_C)) * ceil_div(n, 32);
    TRY(sc.get(&cnt, n_cnt));
    CU(cudaMemsetAsync(cnt, 0, n_cnt * sizeof(uint32_t), stream));
    q.phi_cnt = cnt;
    TRY(sc.get(&dist, plane * p->n_split));
    TRY(sc.get(&kz, plane)); TRY(sc.get(&kt, plane));
    if (!k_out) TRY(sc.get(&kf, plane)); else kf = k_out;
    q.dist_part = dist; q.kz = kz; q.kt = p->Dth ? kt : nullptr; q.kfull = kf;
    q.x_all = xs; q.ld = D; q.g_all = gs; q.g_ld = D;
    const bool mma = phi_z && phi_mma_eligible(n, D, 128);     // hooks: from 128 particles, so the parity tests reach the kernel cheaply
    float* ksp = nullptr;
    if (mma) TRY(sc.get(&ksp, 6 * plane));
    q.k_split = ksp;
    TRY(launch_kmat(p, q, stream));
    if (phi_z) {
        TRY(sc.get(&phi, (size_t)n * D));
        float* part;
        q.j_len = ceil_div(n, PT_J) * PT_J; q.n_jsplit = 1;
        TRY(sc.get(&part, (size_t)n * D));
        q.phi_part = part;
        q.phi_out = phi; q.phi_ld = D;            // phi only: no optimizer step (x_next == null)
        PhiMmaMaps mm;
        if (mma) TRY(fill_mma_maps(mm, ksp, p->Dth > 0, n, n, xs, gs, D));
        TRY(launch_phi(p, q, stream, mma ? &mm : nullptr));
        CU(cudaMemcpy2DAsync(phi_z, fz, phi, fd, fz, n, cudaMemcpyDeviceToDevice, stream));
        if (phi_th && p->Dth) CU(cudaMemcpy2DAsync(phi_th, ft, phi + p->Dz, fd, ft, n, cudaMemcpyDeviceToDevice, stream));
    }
    CU(cudaStreamSynchronize(stream));
    return DIBS_OK;
}

extern "C" int dibs_kernel_matrix(dibs_plan* p, const float* z, const float* theta, int32_t n, float* k_out, void* stream) {
    if (!p || !z || !k_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_kernel_matrix: bad arguments");
    if (p->Dth && !theta) return fail(DIBS_ERR_INVALID_ARG, "theta required");
    return hook_pair(p, z, theta, nullptr, nullptr, n, k_out, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int dibs_svgd_phi(dibs_plan* p, const float* z, const float* theta, const float* grad_z, const float* grad_theta,
                             int32_t n, float* phi_z_out, float* phi_theta_out, void* stream) {
    if (!p || !z || !grad_z || !phi_z_out || n < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_svgd_phi: bad arguments");
    if (p->Dth && (!theta || !grad_theta || !phi_theta_out)) return fail(DIBS_ERR_INVALID_ARG, "theta arguments required");
    return hook_pair(p, z, theta, grad_z, grad_theta, n, nullptr, phi_z_out, phi_theta_out, (cudaStream_t)stream);
}

// host-side key arithmetic for the Python layer (random.split before the jit boundary, svgd.py:294,751)
extern "C" int dibs_prng_split(const uint32_t* key_host, int32_t num, int32_t partitionable, uint32_t* out_host) {
    if (!key_host || !out_host || num < 1) return fail(DIBS_ERR_INVALID_ARG, "dibs_prng_split: bad arguments");
    const uint2 key = make_uint2(key_host[0], key_host[1]);
    for (int r = 0; r < num; ++r) {
        uint2 row = jax_split_row(key, (uint32_t)r, (uint32_t)num, partitionable != 0);
        out_host[2 * r] = row.x; out_host[2 * r + 1] = row.y;
    }
    return DIBS_OK;
}

#ifdef DIBS_PHI_TRACE
// debug builds only (python -m dibs_b200.build is never run with this macro): copy out the phi pipeline trace
extern "C" int dibs_debug_phi_trace(unsigned long long* out_host, int n_words) {
    cudaDeviceSynchronize();
    CU(cudaMemcpyFromSymbol(out_host, g_phi_trace, (size_t)n_words * sizeof(unsigned long long)));
    return DIBS_OK;
}
#endif
