"""Host-side mirror of the reference's ``DiBS`` backbone (dibs/inference/dibs.py:12-658).

Every method keeps the reference's name, argument meaning and batching (``eltwise_*`` = batched over
particles) but dispatches to the CUDA kernels through the C ABI (dibs_b200/_native.py); there is no
Python/PyTorch implementation of the math and no fallback.  Arrays are torch CUDA tensors:

  z       float32 [..., d, k, 2]      theta   the likelihood model's pytree (or its flat [M, Dtheta] form)
  keys    uint32 bit patterns [.., 2] (any integer dtype / numpy / torch accepted)
"""
import collections
import ctypes
import hashlib
import os

import numpy as np
import torch

from .. import _native as nat


def as_key(key):
    """-> numpy uint32[2] (JAX ``PRNGKey`` bit pattern)."""
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    key = np.asarray(key)
    if key.dtype.kind == "i":
        key = key.astype(np.int64) & 0xFFFFFFFF
    return np.ascontiguousarray(key.astype(np.uint32))


def PRNGKey(seed):
    """``jax.random.PRNGKey``: uint32[2] = [seed >> 32, seed & 0xffffffff]."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num=2, partitionable=False):
    """``jax.random.split`` on the host (key handling outside the step loop, svgd.py:294,751)."""
    key = as_key(key)
    out = np.zeros((num, 2), np.uint32)
    nat.check(nat.lib().dibs_prng_split(nat.ptr(key), num, int(partitionable), nat.ptr(out)))
    return out


def keys_to_device(keys, device):
    keys = as_key(keys)
    return torch.from_numpy(keys.view(np.int32).copy()).to(device)


def keys_from_device(t):
    return t.detach().cpu().numpy().view(np.uint32).copy()


# Process-level cache of native plans.  A plan owns the workspace, the data-dependent precomputation (QR factor of x,
# BGe statistics), the captured CUDA graphs of the step and -- on several GPUs -- the opened CUDA-IPC peer buffers;
# none of that depends on WHICH model object asks, only on the POD configuration, the data and the process group.
# Keying the cache on exactly those lets a new ``JointDiBS(...)`` on the same problem reuse all of it (the reference
# gets the same effect from jit's compilation cache, svgd.py:270).  Least-recently-used plans beyond the bound are
# destroyed (their device memory is released as soon as no caller still holds them).
_PLAN_CACHE = collections.OrderedDict()
_PLAN_CACHE_MAX = max(1, int(os.environ.get("DIBS_B200_PLAN_CACHE", "8")))


def clear_plan_cache():
    """Drop every cached native plan (frees their device memory once unreferenced)."""
    _PLAN_CACHE.clear()


def _digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        if a is None:
            h.update(b"-")
            continue
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
        a = np.ascontiguousarray(a)
        h.update(str(a.shape).encode() + str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


class DiBS:
    """Backbone shared by :class:`MarginalDiBS` and :class:`JointDiBS` (reference: dibs/inference/dibs.py:12-78)."""

    def __init__(self, *, x, interv_mask, graph_model, likelihood_model, joint, kernel_obj, optimizer, optimizer_param,
                 alpha_linear, beta_linear, tau, n_grad_mc_samples, n_acyclicity_mc_samples, grad_estimator_z,
                 score_function_baseline, latent_prior_std, verbose, device=None, prng_partitionable=False):
        if not torch.cuda.is_available():
            raise RuntimeError("dibs_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.x, self.interv_mask, self._data_digest = self._upload_data(x, interv_mask, likelihood_model)
        self.n_vars = self.x.shape[-1]
        self.graph_model = graph_model
        self.likelihood_model = likelihood_model
        self.joint = joint
        self.kernel = kernel_obj
        if optimizer not in nat.OPTIMIZER:
            raise ValueError()                      # svgd.py:121-122
        self.optimizer = optimizer
        self.optimizer_param = optimizer_param
        self.alpha_linear = alpha_linear
        self.beta_linear = beta_linear
        self.alpha = lambda t: (alpha_linear * t)
        self.beta = lambda t: (beta_linear * t)
        self.tau = tau
        self.n_grad_mc_samples = n_grad_mc_samples
        self.n_acyclicity_mc_samples = n_acyclicity_mc_samples
        self.grad_estimator_z = grad_estimator_z
        self.score_function_baseline = score_function_baseline
        self.latent_prior_std = latent_prior_std
        self.verbose = verbose
        self.prng_partitionable = prng_partitionable
        for obj, what in ((graph_model, "graph_model"), (likelihood_model, "likelihood_model")):
            if not hasattr(obj, "native_kind"):
                raise NotImplementedError(f"{what} {type(obj).__name__} has no native implementation in dibs_b200")
        if hasattr(likelihood_model, "check_native"):
            likelihood_model.check_native()

    # ------------------------------------------------------------------ plan management
    def _dist(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def _std(self, n_dim):
        if self.latent_prior_std:
            return float(self.latent_prior_std)
        return float(np.float32(1.0) / np.sqrt(np.float32(n_dim)))   # svgd.py:142,302 in fp32

    def _upload_data(self, x, interv_mask, likelihood_model):
        """-> (x float32 [N, d] on the device, mask int32 [N, d] or None, digest of the data a plan precomputes from)."""
        xh = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        xd = xh.to(self.device, torch.float32, non_blocking=True).contiguous()
        md, mh = None, None
        if interv_mask is not None:
            mh = interv_mask if isinstance(interv_mask, torch.Tensor) else torch.as_tensor(np.asarray(interv_mask))
            md = mh.to(self.device, torch.int32, non_blocking=True).contiguous()
        mean_obs = getattr(likelihood_model, "mean_obs", None)
        return xd, md, _digest(xh, mh, None if mean_obs is None else np.asarray(mean_obs, np.float32))

    def _call(self, fn, *args):
        """One native call with this model's device current (plans allocate and launch on the current device)."""
        with torch.cuda.device(self.device):
            nat.check(getattr(nat.lib(), fn)(*args))

    def _config(self, n_particles, n_dim, world, rank):
        if self.grad_estimator_z not in nat.ESTIMATOR:
            raise ValueError(f'Unknown gradient estimator `{self.grad_estimator_z}`')     # dibs.py:318
        lm, gm = self.likelihood_model, self.graph_model
        c = nat.DibsConfig()
        c.n_vars, c.n_dim, c.n_particles = self.n_vars, n_dim, n_particles
        c.joint = int(self.joint)
        c.likelihood = nat.LIK[lm.native_kind]
        c.graph_prior = nat.PRIOR[gm.native_kind]
        c.grad_estimator_z = nat.ESTIMATOR[self.grad_estimator_z]
        c.optimizer = nat.OPTIMIZER[self.optimizer]
        c.n_grad_mc_samples = self.n_grad_mc_samples
        c.n_acyclicity_mc_samples = self.n_acyclicity_mc_samples
        c.hidden = getattr(lm, "hidden", 0) if lm.native_kind == "densenn" else 0
        c.activation = nat.ACTIVATION[getattr(lm, "activation", "relu")] if lm.native_kind == "densenn" else 0
        c.prng_partitionable = int(self.prng_partitionable)
        c.alpha_linear, c.beta_linear, c.tau = self.alpha_linear, self.beta_linear, self.tau
        c.score_function_baseline = self.score_function_baseline
        c.latent_prior_std = self._std(n_dim)
        kn = self.kernel
        c.h_latent = getattr(kn, "h_latent", getattr(kn, "h", 5.0))
        c.h_theta = getattr(kn, "h_theta", 500.0)
        c.scale_latent = getattr(kn, "scale_latent", getattr(kn, "scale", 1.0))
        c.scale_theta = getattr(kn, "scale_theta", 1.0)
        c.stepsize = self.optimizer_param["stepsize"]
        c.er_p = getattr(gm, "p", 0.0)
        c.obs_noise = getattr(lm, "obs_noise", 0.1)
        c.mean_edge = getattr(lm, "mean_edge", 0.0)
        c.sig_edge = getattr(lm, "sig_edge", 1.0)
        c.min_edge = getattr(lm, "min_edge", 0.5)
        c.sig_param = getattr(lm, "sig_param", 1.0)
        c.bge_alpha_mu = getattr(lm, "alpha_mu", 1.0)
        c.bge_alpha_lambd = getattr(lm, "alpha_lambd", self.n_vars + 2)
        c.world_size, c.rank = world, rank
        return c

    def _plan(self, n_particles, n_dim=None, sharded=False, data=None):
        """Native plan for (M, k); ``sharded`` plans split particles over the torch.distributed ranks.  ``data`` =
        (x, mask, digest) scores against another data set (held-out likelihoods); default: the model's own."""
        n_dim = n_dim or self.n_vars
        world, rank = self._dist() if sharded else (1, 0)
        c = self._config(int(n_particles), int(n_dim), world, rank)
        x, mask, digest = data or (self.x, self.interv_mask, self._data_digest)
        group = None
        if world > 1:
            import torch.distributed as dist
            group = id(dist.distributed_c10d._get_default_group())
        key = (bytes(c), digest, self.device.index, group)
        plan = _PLAN_CACHE.get(key)
        if plan is not None:
            _PLAN_CACHE.move_to_end(key)
            return plan
        handle = ctypes.c_void_p()
        self._call("dibs_plan_create", ctypes.byref(c), ctypes.byref(handle))
        plan = _Plan(handle, c, self.device)
        mean_obs = getattr(self.likelihood_model, "mean_obs", None)
        mean_np = None if mean_obs is None else np.ascontiguousarray(np.asarray(mean_obs, np.float32))
        self._call("dibs_set_data", handle, nat.ptr(x), nat.ptr(mask), x.shape[0], nat.ptr(mean_np), self._stream())
        if world > 1:
            plan.attach_nccl()
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
            _PLAN_CACHE.popitem(last=False)          # least recently used; destroyed when unreferenced
        return plan

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        return t.to(self.device, torch.float32).contiguous()

    def _flat_theta(self, theta):
        if theta is None:
            return None
        if isinstance(theta, (torch.Tensor, np.ndarray)) and theta.ndim == 2 and \
                theta.shape[1] == self.likelihood_model.theta_dim():
            return self._f32(theta)
        return self._f32(self.likelihood_model.flatten(theta))

    # ------------------------------------------------------------------ backbone functionality (dibs.py:84-184)
    def particle_to_g_lim(self, z):
        """G for alpha = infinity: (U V^T > 0), zero diagonal (dibs.py:84-99).  z [..., d, k, 2] -> int32 [..., d, d]."""
        z = self._f32(z)
        lead, (d, k) = z.shape[:-3], z.shape[-3:-1]
        n = int(np.prod(lead)) if lead else 1
        out = torch.empty((n, d, d), dtype=torch.int32, device=self.device)
        plan = self._plan(max(n, 1), k)
        self._call("dibs_particle_to_g_lim", plan.handle, nat.ptr(z), n, nat.ptr(out), self._stream())
        return out.reshape(*lead, d, d)

    def edge_probs(self, z, t):
        """sigmoid(alpha(t) U V^T), zero diagonal (dibs.py:168-184)."""
        z = self._f32(z)
        lead, (d, k) = z.shape[:-3], z.shape[-3:-1]
        n = int(np.prod(lead)) if lead else 1
        out = torch.empty((n, d, d), dtype=torch.float32, device=self.device)
        plan = self._plan(max(n, 1), k)
        self._call("dibs_edge_probs", plan.handle, nat.ptr(z), n, int(t), nat.ptr(out), self._stream())
        return out.reshape(*lead, d, d)

    def sample_g(self, p, subk, n_samples):
        """Bernoulli graphs from edge probabilities p [d, d] (or [n, d, d] with subk [n, 2]) (dibs.py:102-119)."""
        p = self._f32(p)
        single = p.dim() == 2
        p = p.reshape(-1, self.n_vars, self.n_vars)
        n = p.shape[0]
        keys = keys_to_device(np.asarray(as_key(subk)).reshape(n, 2), self.device)
        out = torch.empty((n, n_samples, self.n_vars, self.n_vars), dtype=torch.int32, device=self.device)
        plan = self._plan(n)
        self._call("dibs_sample_graphs", plan.handle, nat.ptr(p), nat.ptr(keys), n, int(n_samples), nat.ptr(out),
                                               self._stream())
        return out[0] if single else out

    def sample_soft_g(self, z, subk, n_samples, t):
        """logistic noise (random.logistic) + particle_to_soft_graph (dibs.py:121-140,431): z [n,d,k,2] -> [n,S,d,d]."""
        z = self._f32(z)
        single = z.dim() == 3
        z = z.reshape(-1, *z.shape[-3:])
        n, d, k = z.shape[0], z.shape[1], z.shape[2]
        keys = keys_to_device(np.asarray(as_key(subk)).reshape(n, 2), self.device)
        out = torch.empty((n, n_samples, d, d), dtype=torch.float32, device=self.device)
        plan = self._plan(n, k)
        self._call("dibs_soft_graphs", plan.handle, nat.ptr(z), nat.ptr(keys), n, int(n_samples), int(t), nat.ptr(out),
                                             self._stream())
        return out[0] if single else out

    # ------------------------------------------------------------------ likelihood estimators (dibs.py:255-551)
    def eltwise_log_joint_prob(self, gs, single_theta, rng=None):
        """log p(Theta, D | G) batched over graphs: gs [S, d, d] (+ single theta) or gs [n, S, d, d] (+ theta batch)."""
        gs = self._f32(gs)
        single = gs.dim() == 3
        gs = gs.reshape(-1, *gs.shape[-3:]) if not single else gs[None]
        n, s = gs.shape[0], gs.shape[1]
        theta = None
        if self.joint:
            theta = single_theta
            if single:
                theta = _add_leading(theta)
            theta = self._flat_theta(theta)
        out = torch.empty((n, s), dtype=torch.float32, device=self.device)
        plan = self._plan(n)
        self._call("dibs_log_joint_prob", plan.handle, nat.ptr(gs.contiguous()), nat.ptr(theta), n, s, nat.ptr(out),
                                                self._stream())
        return out[0] if single else out

    def eltwise_grad_z_likelihood(self, zs, thetas, baselines, t, subkeys):
        """Batch of estimators of grad_Z log p(Theta, D | Z) (dibs.py:295-321) -> (grads [n,d,k,2], baselines [n])."""
        if self.grad_estimator_z not in nat.ESTIMATOR:
            raise ValueError(f'Unknown gradient estimator `{self.grad_estimator_z}`')
        zs = self._f32(zs)
        n, d, k = zs.shape[0], zs.shape[1], zs.shape[2]
        theta = self._flat_theta(thetas) if self.joint else None
        base = self._f32(baselines)
        keys = keys_to_device(np.asarray(as_key(subkeys)).reshape(n, 2), self.device)
        grad = torch.empty_like(zs)
        base_out = torch.empty_like(base)
        plan = self._plan(n, k)
        self._call("dibs_grad_z_likelihood", plan.handle, nat.ptr(zs), nat.ptr(theta), nat.ptr(base), int(t),
                                                   nat.ptr(keys), n, nat.ptr(grad), nat.ptr(base_out), self._stream())
        return grad, base_out

    def eltwise_grad_theta_likelihood(self, zs, thetas, t, subkeys):
        """Batch of estimators of grad_Theta log p(Theta, D | Z) (dibs.py:467-551) -> flat [n, Dtheta]."""
        zs = self._f32(zs)
        n, k = zs.shape[0], zs.shape[2]
        theta = self._flat_theta(thetas)
        keys = keys_to_device(np.asarray(as_key(subkeys)).reshape(n, 2), self.device)
        grad = torch.empty_like(theta)
        plan = self._plan(n, k)
        self._call("dibs_grad_theta_likelihood", plan.handle, nat.ptr(zs), nat.ptr(theta), int(t), nat.ptr(keys), n,
                                                       nat.ptr(grad), self._stream())
        return grad

    # ------------------------------------------------------------------ latent prior (dibs.py:557-658)
    def eltwise_grad_latent_prior(self, zs, subkeys, t, constraint_only=False):
        """grad_Z log p(Z) = -beta(t) grad E[h(G)] - Z / sigma^2 + grad log p(G_alpha(Z)) (dibs.py:626-658)."""
        zs = self._f32(zs)
        n, k = zs.shape[0], zs.shape[2]
        keys = keys_to_device(np.asarray(as_key(subkeys)).reshape(n, 2), self.device)
        grad = torch.empty_like(zs)
        plan = self._plan(n, k)
        self._call("dibs_grad_latent_prior", plan.handle, nat.ptr(zs), int(t), nat.ptr(keys), n, int(constraint_only),
                                                   nat.ptr(grad), self._stream())
        return grad

    def grad_constraint_gumbel(self, single_z, key, t):
        """mean over Gumbel-soft graphs of grad_Z h(G) for one particle (dibs.py:576-601)."""
        return self.eltwise_grad_latent_prior(self._f32(single_z)[None], np.asarray(as_key(key))[None], t,
                                              constraint_only=True)[0]

    def acyclic_constr(self, g):
        """h(G) = tr((I + G/d)^d) - d, batched (dibs/graph_utils.py:8-30)."""
        g = self._f32(g)
        single = g.dim() == 2
        g = g.reshape(-1, self.n_vars, self.n_vars)
        out = torch.empty((g.shape[0],), dtype=torch.float32, device=self.device)
        plan = self._plan(g.shape[0])
        self._call("dibs_acyclic_constr", plan.handle, nat.ptr(g), g.shape[0], nat.ptr(out), self._stream())
        return out[0] if single else out

    def particle_summary(self, zs, t):
        """Enqueue the device-side progress summary of a particle set (``dibs_particle_summary``): returns
        (record, event) where ``record`` is a pinned host tensor [2 + d*d] = (#cyclic, n, mean edge probabilities) that
        is valid once ``event`` has completed.  Nothing here waits for the GPU."""
        zs = self._f32(zs)
        n, k = zs.shape[0], zs.shape[2]
        rec = torch.empty(2 + self.n_vars * self.n_vars, dtype=torch.float32, pin_memory=True)
        plan = self._plan(n, k)
        self._call("dibs_particle_summary", plan.handle, nat.ptr(zs), n, int(t), nat.ptr(rec), self._stream())
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return rec, ev, zs                      # zs kept alive until the kernels have run

    def _flush_summary(self):
        pend, self._pending_summary = self._pending_summary, None
        if pend is None:
            return
        t, (rec, ev, _) = pend
        ev.synchronize()                        # enqueued a whole chunk ago: complete unless the chunk was tiny
        self.last_summary = dict(t=t, n_cyclic=int(rec[0].item()), n_particles=int(rec[1].item()),
                                 edge_marginals=rec[2:].reshape(self.n_vars, self.n_vars).clone())
        print(f'iteration {t:6d} | alpha {self.alpha(t):6.1f} | beta {self.beta(t):6.1f} '
              f'| #cyclic {self.last_summary["n_cyclic"]:3d}')

    def visualize_callback(self, ipython=True, save_path=None):
        """Progress callback (dibs.py:661-692), text only -- the reference's matplotlib grid is out of scope.  The
        summary (#cyclic graphs among G_lim, mean edge probabilities) is computed on the device and streamed to pinned
        host memory; a callback prints the PREVIOUS chunk's line, so it never blocks on the chunk it was called for
        (``sample()`` flushes the last line when the loop ends; also available as ``self.last_summary``)."""
        self._pending_summary = None

        def callback(**kwargs):
            self._flush_summary()
            self._pending_summary = (kwargs["t"], self.particle_summary(kwargs["zs"], kwargs["t"]))
        return callback


def _add_leading(theta):
    if theta is None:
        return None
    if isinstance(theta, torch.Tensor):
        return theta[None]
    if isinstance(theta, np.ndarray):
        return theta[None]
    return type(theta)(_add_leading(t) for t in theta)


class _Plan:
    """Owns one native ``dibs_plan`` (shared between model objects through the process-level cache)."""

    def __init__(self, handle, cfg, device):
        self.handle, self.cfg, self.device = handle, cfg, device
        self.n_local = cfg.n_particles // cfg.world_size
        self.row0 = cfg.rank * self.n_local
        self.theta_dim = nat.lib().dibs_theta_dim(handle)
        self.exchange = "none"

    def check(self):
        """Raise if a bounded peer wait of the exchange timed out (``dibs_plan_status``)."""
        nat.check(nat.lib().dibs_plan_status(self.handle))

    def attach_nccl(self):
        """Multi-GPU exchange set-up, once per plan: peer-memory pushes over NVLink (CUDA IPC) when every rank can open
        every peer's buffers, else two NCCL communicators -- the gradient exchange (critical path) and the particle
        exchange (side branch of the step graph), independent so the two may run in either order on different ranks."""
        if os.environ.get("DIBS_B200_NO_P2P") != "1" and self._attach_peer_memory():
            self.exchange = "peer-memory"
            return                      # rows travel by peer-memory pushes: no communicator needed
        self._attach_one()
        self._attach_one()
        self.exchange = "nccl"

    def _agree(self, ok, device):
        """MIN over the ranks of a local success flag (every rank takes the same branch afterwards)."""
        import torch.distributed as dist
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return int(flag.item()) == 1

    def _attach_peer_memory(self):
        """Exchange CUDA-IPC handles of the plan's particle / gradient / flag buffers and open the peers' copies: the
        step then pushes rows over NVLink peer memory instead of calling NCCL (kernels_peer.cuh).  Returns False --
        on EVERY rank, with nothing left open -- when any rank cannot export or attach (ranks on different nodes, more
        ranks than the flag table holds, no P2P between two devices); the caller then falls back to NCCL."""
        import torch.distributed as dist
        dev = self.device
        world = self.cfg.world_size
        if world > nat.PEER_MAX:
            return False                # same on every rank: no collective needed to agree
        mine = np.zeros(320, np.uint8)
        with torch.cuda.device(dev):
            ok = nat.lib().dibs_plan_ipc_export(self.handle, nat.ptr(mine)) == 0
        on_gpu = dist.get_backend() == "nccl"
        cdev = dev if on_gpu else torch.device("cpu")
        blob = torch.from_numpy(mine).to(cdev)
        allb = torch.empty(world * 320, dtype=torch.uint8, device=cdev)
        dist.all_gather_into_tensor(allb, blob)
        if not self._agree(ok, cdev):
            return False
        buf = np.ascontiguousarray(allb.cpu().numpy())
        with torch.cuda.device(dev):
            ok = nat.lib().dibs_plan_ipc_attach(self.handle, nat.ptr(buf)) == 0
        if not self._agree(ok, cdev):   # also the barrier: every rank has opened its peers
            with torch.cuda.device(dev):
                nat.lib().dibs_plan_ipc_detach(self.handle)     # close what this rank opened; back to the NCCL path
            return False
        return True

    def _attach_one(self):
        import torch.distributed as dist
        dev = self.device
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.cfg.rank == 0:
            buf = np.zeros(128, np.uint8)
            nat.check(nat.lib().dibs_nccl_unique_id(nat.ptr(buf)))
            ident = torch.from_numpy(buf)
        on_gpu = dist.get_backend() == "nccl"
        ident = ident.to(dev) if on_gpu else ident
        dist.broadcast(ident, src=0)
        buf = np.ascontiguousarray(ident.cpu().numpy())
        with torch.cuda.device(dev):
            nat.check(nat.lib().dibs_plan_attach_nccl(self.handle, nat.ptr(buf)))

    def __del__(self):
        try:
            if self.handle:
                with torch.cuda.device(self.device):
                    nat.lib().dibs_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
