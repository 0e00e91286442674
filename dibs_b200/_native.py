"""ctypes binding of the C ABI in include/dibs_b200.h (the drop-in boundary).

There is no fallback: if ``libdibs_b200.so`` is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdibs_b200.so")

LIK = {"bge": 0, "lingauss": 1, "densenn": 2}
PRIOR = {"er": 0, "sf": 1, "uniform": 2}
ESTIMATOR = {"score": 0, "reparam": 1}
OPTIMIZER = {"gd": 0, "rmsprop": 1}
ACTIVATION = {"relu": 0, "tanh": 1, "sigmoid": 2, "leakyrelu": 3}
PEER_MAX = 16          # ranks the peer-memory flag table holds (kernels_peer.cuh)


class DibsConfig(ctypes.Structure):
    """Mirror of ``struct dibs_config``."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "n_vars", "n_dim", "n_particles", "joint", "likelihood", "graph_prior", "grad_estimator_z", "optimizer",
        "n_grad_mc_samples", "n_acyclicity_mc_samples", "hidden", "prng_partitionable")] + [
        (n, ctypes.c_float) for n in (
            "alpha_linear", "beta_linear", "tau", "score_function_baseline", "latent_prior_std",
            "h_latent", "h_theta", "scale_latent", "scale_theta", "stepsize", "er_p",
            "obs_noise", "mean_edge", "sig_edge", "min_edge", "sig_param", "bge_alpha_mu", "bge_alpha_lambd")] + [
        ("world_size", ctypes.c_int32), ("rank", ctypes.c_int32), ("activation", ctypes.c_int32)]


_P = ctypes.c_void_p
_I = ctypes.c_int32

# name -> (restype, argtypes); must list every symbol declared in include/dibs_b200.h
PROTOTYPES = {
    "dibs_last_error": (ctypes.c_char_p, []),
    "dibs_abi_version": (ctypes.c_int, []),
    "dibs_plan_create": (ctypes.c_int, [ctypes.POINTER(DibsConfig), ctypes.POINTER(_P)]),
    "dibs_plan_destroy": (ctypes.c_int, [_P]),
    "dibs_theta_dim": (ctypes.c_int, [_P]),
    "dibs_plan_status": (ctypes.c_int, [_P]),
    "dibs_set_data": (ctypes.c_int, [_P, _P, _P, _I, _P, _P]),
    "dibs_nccl_unique_id": (ctypes.c_int, [_P]),
    "dibs_plan_attach_nccl": (ctypes.c_int, [_P, _P]),
    "dibs_plan_ipc_export": (ctypes.c_int, [_P, _P]),
    "dibs_plan_ipc_attach": (ctypes.c_int, [_P, _P]),
    "dibs_plan_ipc_detach": (ctypes.c_int, [_P]),
    "dibs_svgd_steps": (ctypes.c_int, [_P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "dibs_svgd_steps_timed": (ctypes.c_int, [_P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, ctypes.c_int64, _P, _P]),
    "dibs_init_particles": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "dibs_edge_probs": (ctypes.c_int, [_P, _P, _I, _I, _P, _P]),
    "dibs_particle_to_g_lim": (ctypes.c_int, [_P, _P, _I, _P, _P]),
    "dibs_sample_graphs": (ctypes.c_int, [_P, _P, _P, _I, _I, _P, _P]),
    "dibs_soft_graphs": (ctypes.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "dibs_log_joint_prob": (ctypes.c_int, [_P, _P, _P, _I, _I, _P, _P]),
    "dibs_grad_z_likelihood": (ctypes.c_int, [_P, _P, _P, _P, _I, _P, _I, _P, _P, _P]),
    "dibs_grad_theta_likelihood": (ctypes.c_int, [_P, _P, _P, _I, _P, _I, _P, _P]),
    "dibs_grad_latent_prior": (ctypes.c_int, [_P, _P, _I, _P, _I, _I, _P, _P]),
    "dibs_acyclic_constr": (ctypes.c_int, [_P, _P, _I, _P, _P]),
    "dibs_particle_summary": (ctypes.c_int, [_P, _P, _I, _I, _P, _P]),
    "dibs_kernel_matrix": (ctypes.c_int, [_P, _P, _P, _I, _P, _P]),
    "dibs_svgd_phi": (ctypes.c_int, [_P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "dibs_prng_split": (ctypes.c_int, [_P, _I, _I, _P]),
    "dibs_launch_count": (ctypes.c_int64, []),
}

# kernels of one step in the order of DIBS_PHASE_*; "assemble" is hook-only: in the step loop the assemble step runs
# inside the gradient kernels (the serial per-kernel mode charges it to "acyclic")
PHASES = ("mc_theta", "mc_z", "acyclic", "assemble", "allgather", "pair_dist", "pair_kernel", "phi_update", "scores")

_lib = None


def lib():
    """Load (once) and return the native library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m dibs_b200.build` "
                               "(dibs_b200 has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


_ERRORS = {-1: ValueError, -2: NotImplementedError, -3: RuntimeError, -4: RuntimeError, -5: RuntimeError}


def check(status):
    """Map a ``dibs_status`` to the exception type the reference would raise."""
    if status != 0:
        msg = lib().dibs_last_error().decode("utf-8", "replace")
        raise _ERRORS.get(status, RuntimeError)(msg or f"dibs_b200 native error {status}")


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous(), "native calls need contiguous tensors"
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)
