/*
 * dibs_b200 -- C ABI of the B200-native DiBS SVGD particle-update hot path.
 *
 * Plain C, plain pointers and sizes; no torch / CUDA runtime types in the signatures
 * (a CUDA stream is passed as an opaque `void*` holding a `cudaStream_t`; every data
 * pointer is a DEVICE pointer unless the name ends in `_host`).
 *
 * Every entry point replaces one piece of the reference (larslorch/dibs @ 5350d1a; paths
 * relative to the reference repository).  The reference has no FFI of its own -- it is pure
 * Python/JAX -- so the "interface it replaces" is the Python method cited at each entry.
 * INTEGRATION.md shows the ctypes binding (dibs_b200/_native.py) a maintainer would add.
 *
 * All functions return 0 on success or a negative `dibs_status`; they never throw and never
 * fall back to a CPU path.  `dibs_last_error()` returns a human-readable message for the
 * calling thread's last failure.
 *
 * Particle memory layout (same as the reference, dibs/inference/svgd.py:146,323):
 *   Z      float32 [n_particles, n_vars, n_dim, 2]   (U, V interleaved innermost)
 *   Theta  float32 [n_particles, theta_dim]          flattened per particle:
 *            LinearGaussian:         theta[i, j]                       (d*d)
 *            DenseNonlinearGaussian: W1[j,i,h] | b1[j,h] | W2[j,h] | b2[j]   (hidden_layers=(H,), any of the
 *                                    reference's four activations)
 *   keys   uint32  [.., 2]                           JAX threefry keys
 *   G      int32   [.., n_vars, n_vars]
 */
#ifndef DIBS_B200_H
#define DIBS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DIBS_OK = 0,
    DIBS_ERR_INVALID_ARG = -1,   /* ValueError in the Python layer                         */
    DIBS_ERR_UNSUPPORTED = -2,   /* NotImplementedError: combination outside the native table */
    DIBS_ERR_CUDA = -3,          /* RuntimeError: CUDA failure                              */
    DIBS_ERR_NCCL = -4,          /* RuntimeError: NCCL failure                              */
    DIBS_ERR_STATE = -5          /* RuntimeError: call order (e.g. no data set)             */
} dibs_status;

enum { DIBS_LIK_BGE = 0, DIBS_LIK_LINEAR_GAUSSIAN = 1, DIBS_LIK_DENSE_NONLINEAR_GAUSSIAN = 2 };
enum { DIBS_PRIOR_ERDOS_RENYI = 0, DIBS_PRIOR_SCALE_FREE = 1, DIBS_PRIOR_UNIFORM = 2 };
enum { DIBS_ESTIMATOR_SCORE = 0, DIBS_ESTIMATOR_REPARAM = 1 };
enum { DIBS_OPT_GD = 0, DIBS_OPT_RMSPROP = 1 };
enum { DIBS_ACT_RELU = 0, DIBS_ACT_TANH = 1, DIBS_ACT_SIGMOID = 2, DIBS_ACT_LEAKYRELU = 3 };

/* POD mirror of the constructor arguments of MarginalDiBS / JointDiBS
 * (dibs/inference/svgd.py:60-77, 425-442) and of the plugin objects they receive. */
typedef struct dibs_config {
    int32_t n_vars;                    /* d                                                   */
    int32_t n_dim;                     /* k (n_dim_particles; default d, svgd.py:138-139)     */
    int32_t n_particles;               /* M, global                                            */
    int32_t joint;                     /* 0 = MarginalDiBS, 1 = JointDiBS                      */
    int32_t likelihood;                /* DIBS_LIK_*                                           */
    int32_t graph_prior;               /* DIBS_PRIOR_*                                         */
    int32_t grad_estimator_z;          /* DIBS_ESTIMATOR_* (dibs.py:311-318)                   */
    int32_t optimizer;                 /* DIBS_OPT_* (svgd.py:117-122)                         */
    int32_t n_grad_mc_samples;         /* S                                                    */
    int32_t n_acyclicity_mc_samples;   /* A                                                    */
    int32_t hidden;                    /* H of DenseNonlinearGaussian(hidden_layers=(H,))      */
    int32_t prng_partitionable;        /* 0 = legacy threefry layout (JAX < 0.5), 1 = partitionable */
    float alpha_linear, beta_linear, tau;          /* dibs.py:70-72                           */
    float score_function_baseline;                 /* dibs.py:363-367,388-389                 */
    float latent_prior_std;                        /* sigma_z; caller passes 1/sqrt(k) if None */
    float h_latent, h_theta, scale_latent, scale_theta;   /* kernel.py:16,45-49               */
    float stepsize;                                /* optimizer_param['stepsize']             */
    float er_p;                                    /* ErdosReniDAGDistribution.p (graph.py:27-30) */
    float obs_noise, mean_edge, sig_edge, min_edge; /* LinearGaussian (linearGaussian.py:190-195)  */
    float sig_param;                               /* DenseNonlinearGaussian (nonlinearGaussian.py:105-111) */
    float bge_alpha_mu, bge_alpha_lambd;           /* BGe (linearGaussian.py:43-45)           */
    /* particle sharding (one process per GPU) */
    int32_t world_size;                /* number of ranks (1 = single GPU)                     */
    int32_t rank;                      /* this rank owns particles [rank*M/W, (rank+1)*M/W)    */
    int32_t activation;                /* DenseNonlinearGaussian(activation=): DIBS_ACT_* (nonlinearGaussian.py:52-61) */
} dibs_config;

typedef struct dibs_plan dibs_plan;

const char* dibs_last_error(void);
int dibs_abi_version(void);

/* ---- plan lifecycle -------------------------------------------------------------------
 * replaces: MarginalDiBS.__init__ / JointDiBS.__init__ (svgd.py:60-122, 425-487) binding
 * graph_model / likelihood_model / kernel / optimizer into the inference object. */
int dibs_plan_create(const dibs_config* cfg, dibs_plan** out);
int dibs_plan_destroy(dibs_plan* plan);
int dibs_theta_dim(const dibs_plan* plan);
/* Health of the plan's multi-GPU exchange: DIBS_ERR_STATE once a bounded wait on a peer's rows timed out (a rank
 * died or ran a different number of steps; DIBS_B200_PEER_TIMEOUT_MS, default 10 s), DIBS_OK otherwise.  The
 * reference is single-process and has no counterpart; the closest is the exception a failed jax collective raises. */
int dibs_plan_status(dibs_plan* plan);

/* replaces: the `x=` / `interv_mask=` constructor arguments (svgd.py:86-92, 451-457).
 * x float32 [n_obs, d]; interv_mask int32 [n_obs, d] or NULL (= zeros); bge_mean_obs_host
 * float32 [d] HOST pointer or NULL (= zeros).  Copies into plan-owned device memory and
 * precomputes the data-only BGe statistics (linearGaussian.py:78-94). */
int dibs_set_data(dibs_plan* plan, const float* x, const int32_t* interv_mask, int32_t n_obs,
                  const float* bge_mean_obs_host, void* stream);

/* ---- multi-GPU ----------------------------------------------------------------------------
 * The reference is single-device; sharding by particle is new.  NCCL is resolved
 * at run time with dlopen("libnccl.so.2") (the copy torch already loaded). */
int dibs_nccl_unique_id(uint8_t* id128_host);
int dibs_plan_attach_nccl(dibs_plan* plan, const uint8_t* id128_host);
/* Peer-memory exchange (preferred on one NVSwitch node): each rank exports CUDA-IPC handles of its particle /
 * gradient / flag buffers (5 x 64 bytes), the host all-gathers them, every rank opens its peers' buffers; the
 * step then pushes rows straight into peer memory instead of calling NCCL.  handles_out: 320 bytes;
 * all_handles: world_size x 320 bytes in rank order. */
int dibs_plan_ipc_export(dibs_plan* plan, uint8_t* handles_out_host);
int dibs_plan_ipc_attach(dibs_plan* plan, const uint8_t* all_handles_host);
/* Undo dibs_plan_ipc_attach (also after a partial failure): close every opened peer buffer and return the plan to
 * the state in which dibs_plan_attach_nccl can be used instead. */
int dibs_plan_ipc_detach(dibs_plan* plan);

/* ---- the hot loop ---------------------------------------------------------------------------
 * replaces: _svgd_loop -> lax.fori_loop over _svgd_step (svgd.py:226-272, 673-727).
 * Runs steps t_start .. t_start+n_steps-1 in place on the LOCAL shard of particles:
 *   z [M_local, d, k, 2], theta [M_local, theta_dim] (NULL for marginal),
 *   v_z / v_theta: RMSprop second-moment state, same shapes (ignored for GD, may be NULL),
 *   key uint32[2]: loop key, replaced by the key after the last step,
 *   sf_baseline [M_local].
 * All work is enqueued on `stream`; the call does not synchronise. */
int dibs_svgd_steps(dibs_plan* plan, int32_t t_start, int32_t n_steps, float* z, float* theta,
                    float* v_z, float* v_theta, uint32_t* key, float* sf_baseline, void* stream);

/* Measurement variant of the same loop for bench.py (the reference has no counterpart: its only timing is a
 * `%%time` notebook cell, examples/dibs_marginal.ipynb:88).  Each step is bracketed by CUDA events on `stream`;
 * if l2_flush_buf is non-NULL it is overwritten (cudaMemsetAsync of l2_flush_bytes > L2 size) before every step,
 * outside the bracket, so every step starts with a cold L2.  per_kernel = 0: steps run exactly like
 * dibs_svgd_steps (CUDA-graph replay), step_ms_host[n_steps] receives each step's device time.
 * per_kernel = 1: steps are launched eagerly with an event after every kernel; phase_ms_host[DIBS_N_PHASES]
 * receives the summed device time of each kernel over the n_steps.  Synchronises `stream`. */
enum {
    DIBS_PHASE_MC_THETA = 0,    /* grad_theta estimator pass  (dibs.py:488-551)              */
    DIBS_PHASE_MC_Z = 1,        /* grad_z likelihood pass     (dibs.py:325-459)              */
    DIBS_PHASE_ACYCLIC = 2,     /* grad_constraint_gumbel     (dibs.py:576-601); in the step loop a particle's last
                                   gradient CTA also runs the assemble step below (chain rule + latent prior + the
                                   next step's key splits, dibs.py:626-658; svgd.py:245,251,695,699,703), so in the
                                   serial per-kernel mode this phase carries it */
    DIBS_PHASE_ASSEMBLE = 3,    /* (hooks only) stand-alone assemble kernel                  */
    DIBS_PHASE_ALLGATHER = 4,   /* NCCL all-gather (multi-GPU fallback path only)            */
    DIBS_PHASE_PAIR_DIST = 5,   /* pairwise squared distances (kernel.py:30,66-71)           */
    DIBS_PHASE_PAIR_KERNEL = 6, /* sum of the feature splits, exp -> K (kernel.py:30,66-71)  */
    DIBS_PHASE_PHI_UPDATE = 7,  /* phi + optimizer step + peer push (svgd.py:194-224,591-670,265,718-719) */
    DIBS_PHASE_STEP_KEYS = 8,   /* next step's raw scores U V^T = the edge-probability pass (dibs.py:179-181) */
    DIBS_N_PHASES = 9
};
int dibs_svgd_steps_timed(dibs_plan* plan, int32_t t_start, int32_t n_steps, float* z, float* theta,
                          float* v_z, float* v_theta, uint32_t* key, float* sf_baseline, void* stream,
                          int32_t per_kernel, void* l2_flush_buf, int64_t l2_flush_bytes,
                          float* step_ms_host, float* phase_ms_host);

/* replaces: _sample_initial_random_particles (svgd.py:125-148, 489-515) incl.
 * likelihood_model.sample_parameters (linearGaussian.py:212-227, nonlinearGaussian.py:155-186).
 * `key` is the sub-key handed to that function; fills the full [M, ...] arrays (every rank
 * draws all particles and keeps its shard). */
int dibs_init_particles(dibs_plan* plan, const uint32_t* key, float* z_all, float* theta_all, void* stream);

/* ---- per-function hooks (same kernels as the loop; used by the parity tests) ----------------- */

/* DiBS.edge_probs (dibs.py:168-184): z [n, d, k, 2], t -> p [n, d, d] */
int dibs_edge_probs(dibs_plan* plan, const float* z, int32_t n, int32_t t, float* p_out, void* stream);
/* DiBS.particle_to_g_lim (dibs.py:84-99): z [n, d, k, 2] -> g int32 [n, d, d] */
int dibs_particle_to_g_lim(dibs_plan* plan, const float* z, int32_t n, int32_t* g_out, void* stream);
/* DiBS.sample_g (dibs.py:102-119): p [n, d, d], keys [n, 2] -> g int32 [n, n_samples, d, d] */
int dibs_sample_graphs(dibs_plan* plan, const float* p, const uint32_t* keys, int32_t n, int32_t n_samples,
                       int32_t* g_out, void* stream);
/* random.logistic + DiBS.particle_to_soft_graph (dibs.py:121-140, 431, 595):
 * z [n, d, k, 2], keys [n, 2], t -> soft g float32 [n, n_samples, d, d] */
int dibs_soft_graphs(dibs_plan* plan, const float* z, const uint32_t* keys, int32_t n, int32_t n_samples,
                     int32_t t, float* g_out, void* stream);
/* DiBS.eltwise_log_joint_prob (dibs.py:255-269) -> plugin interventional_log_{joint,marginal}_prob:
 * g float32 [n, n_samples, d, d] (hard 0/1 or soft), theta [n, theta_dim] or NULL -> lp [n, n_samples] */
int dibs_log_joint_prob(dibs_plan* plan, const float* g, const float* theta, int32_t n, int32_t n_samples,
                        float* lp_out, void* stream);
/* DiBS.eltwise_grad_z_likelihood (dibs.py:295-459): per-particle subkeys [n, 2] ->
 * grad [n, d, k, 2], baselines_out [n] */
int dibs_grad_z_likelihood(dibs_plan* plan, const float* z, const float* theta, const float* baselines,
                           int32_t t, const uint32_t* keys, int32_t n, float* grad_out, float* baselines_out,
                           void* stream);
/* DiBS.eltwise_grad_theta_likelihood (dibs.py:467-551) -> grad [n, theta_dim] */
int dibs_grad_theta_likelihood(dibs_plan* plan, const float* z, const float* theta, int32_t t,
                               const uint32_t* keys, int32_t n, float* grad_out, void* stream);
/* DiBS.eltwise_grad_latent_prior (dibs.py:626-658) -> grad [n, d, k, 2];
 * constraint_only != 0 returns grad_constraint_gumbel alone (dibs.py:576-601) */
int dibs_grad_latent_prior(dibs_plan* plan, const float* z, int32_t t, const uint32_t* keys, int32_t n,
                           int32_t constraint_only, float* grad_out, void* stream);
/* acyclic_constr_nograd (graph_utils.py:8-28), vmapped: g float32 [n, d, d] -> h [n] */
int dibs_acyclic_constr(dibs_plan* plan, const float* g, int32_t n, float* h_out, void* stream);
/* Progress summary for the callback path, replacing the host-side body of DiBS.visualize_callback (dibs.py:661-692:
 * particle_to_g_lim + edge_probs + `(elwise_acyclic_constr_nograd(gs) > 0).sum()`): z [n, d, k, 2], t ->
 * summary_host float32 [2 + d*d] in PINNED host memory = { #particles whose G_lim is cyclic, n, mean edge
 * probabilities }.  Enqueues two kernels and an asynchronous copy on `stream`; never synchronises (except to grow its
 * scratch), so a callback does not stall the step loop -- read the record after an event recorded behind this call. */
int dibs_particle_summary(dibs_plan* plan, const float* z, int32_t n, int32_t t, float* summary_host, void* stream);
/* _f_kernel_mat (svgd.py:165-176, 537-551; kernel.py:20-30, 52-71):
 * z [n, d*k*2], theta [n, theta_dim] or NULL -> K [n, n] */
int dibs_kernel_matrix(dibs_plan* plan, const float* z, const float* theta, int32_t n, float* k_out, void* stream);
/* _parallel_update_z / _parallel_update_theta (svgd.py:194-224, 591-670): particles + their
 * log-prob gradients -> phi_z [n, d*k*2], phi_theta [n, theta_dim] (NULL for marginal) */
int dibs_svgd_phi(dibs_plan* plan, const float* z, const float* theta, const float* grad_z, const float* grad_theta,
                  int32_t n, float* phi_z_out, float* phi_theta_out, void* stream);

/* random.split(key, num) on HOST memory (key handling before the jit boundary, svgd.py:294,751):
 * key_host uint32[2] -> out_host uint32[num, 2] */
int dibs_prng_split(const uint32_t* key_host, int32_t num, int32_t partitionable, uint32_t* out_host);

/* number of kernels this library launched since load (bench.py's `gpu_launches`) */
int64_t dibs_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DIBS_B200_H */
