"""``MarginalDiBS`` / ``JointDiBS`` with the reference's constructor and ``sample()`` signatures
(dibs/inference/svgd.py:17-375, 380-844); the SVGD loop itself is one native call per callback chunk.

Differences a caller can observe, all deliberate:
  * arrays are torch CUDA tensors (the reference returns jax / numpy arrays);
  * plugins are recognised by class (``native_kind``), arbitrary differentiable Python plugins raise
    ``NotImplementedError`` -- native code cannot autodiff them (SURVEY 'Hard parts');
  * with an initialised ``torch.distributed`` process group, particles are sharded over the ranks (rows exchanged by
    peer-memory pushes over NVLink inside the native loop, NCCL all-gathers as the fallback); every rank returns
    the full particle set, bit-identical to a single-GPU run.
"""
import numpy as np
import torch

from .. import _native as nat
from ..kernel import AdditiveFrobeniusSEKernel, JointAdditiveFrobeniusSEKernel
from .dibs import DiBS, as_key, split, keys_to_device


class _SVGDBase(DiBS):
    _joint = False

    def _sample_initial_random_particles(self, *, key, n_particles, n_dim=None, plan=None):
        """Z ~ N(0, std^2) (and Theta ~ likelihood_model.sample_parameters) for ALL particles
        (svgd.py:125-148, 489-515).  Returns z [M, d, k, 2] (and flat theta [M, Dtheta])."""
        n_dim = n_dim or self.n_vars
        plan = plan or self._plan(n_particles, n_dim)
        z = torch.empty((n_particles, self.n_vars, n_dim, 2), dtype=torch.float32, device=self.device)
        theta = None
        if self._joint:
            theta = torch.empty((n_particles, plan.theta_dim), dtype=torch.float32, device=self.device)
        kdev = keys_to_device(key, self.device)
        self._call("dibs_init_particles", plan.handle, nat.ptr(kdev), nat.ptr(z), nat.ptr(theta), self._stream())
        return (z, theta) if self._joint else z

    def _f_kernel_mat(self, x_latents, x_thetas=None):
        """Pairwise kernel matrix k((Z,Theta)_a, (Z,Theta)_b) of a particle set with itself (svgd.py:165-176, 537-551)."""
        z = self._f32(x_latents)
        n, k = z.shape[0], z.shape[2]
        theta = self._flat_theta(x_thetas) if self._joint else None
        out = torch.empty((n, n), dtype=torch.float32, device=self.device)
        plan = self._plan(n, k)
        self._call("dibs_kernel_matrix", plan.handle, nat.ptr(z.reshape(n, -1)), nat.ptr(theta), n, nat.ptr(out),
                                               self._stream())
        return out

    def _parallel_update(self, z, theta, grad_z, grad_theta):
        """phi for every particle (svgd.py:218-224, 617-670): returns (phi_z [n,d,k,2], phi_theta flat or None)."""
        z = self._f32(z)
        n, k = z.shape[0], z.shape[2]
        gz = self._f32(grad_z)
        th = self._flat_theta(theta) if self._joint else None
        gth = self._flat_theta(grad_theta) if self._joint else None
        phi_z = torch.empty_like(z)
        phi_th = torch.empty_like(th) if self._joint else None
        plan = self._plan(n, k)
        self._call("dibs_svgd_phi", plan.handle, nat.ptr(z.reshape(n, -1)), nat.ptr(th), nat.ptr(gz.reshape(n, -1)),
                                          nat.ptr(gth), n, nat.ptr(phi_z), nat.ptr(phi_th), self._stream())
        return phi_z, phi_th

    def _svgd_loop(self, start, n_steps, init):
        """``lax.fori_loop(start, start + n_steps, _svgd_step)`` on caller-provided state (svgd.py:270-272, 725-727).

        init = (z [M,d,k,2], theta flat [M,Dtheta] or None, v_z, v_theta, key uint32[2], sf_baseline [M]); returns the
        same tuple after the steps (single-GPU plan, all M particles local).
        """
        z, theta, v_z, v_theta, key, sf = init
        z = self._f32(z).clone()
        m, k = z.shape[0], z.shape[2]
        theta = self._flat_theta(theta).clone() if self._joint else None
        v_z = self._f32(v_z).clone()
        v_theta = self._f32(v_theta).clone() if self._joint else None
        sf = self._f32(sf).clone()
        key_dev = keys_to_device(key, self.device)
        plan = self._plan(m, k)
        self._call("dibs_svgd_steps", plan.handle, int(start), int(n_steps), nat.ptr(z), nat.ptr(theta),
                   nat.ptr(v_z), nat.ptr(v_theta), nat.ptr(key_dev), nat.ptr(sf), self._stream())
        torch.cuda.synchronize(self.device)
        plan.check()
        from .dibs import keys_from_device
        return z, theta, v_z, v_theta, keys_from_device(key_dev), sf

    def _score_held_out(self, g, theta, x_ho, interv_msk_ho):
        """Score n graphs (one per particle) against ANOTHER data set through the native scorer: a plan bound to
        (x_ho, mask) -- cached like every plan -- and one ``dibs_log_joint_prob`` launch with S = 1."""
        data = self._upload_data(x_ho, interv_msk_ho, self.likelihood_model)
        gt = torch.as_tensor(g, device=self.device).to(torch.float32)
        if gt.dim() == 2:
            gt = gt[None]
        n = gt.shape[0]
        th = self._flat_theta(theta) if self._joint else None
        reps = 2 if self._joint else 1      # joint scorers draw sample PAIRS: pass each graph twice, keep the first
        gs = gt[:, None].expand(n, reps, *gt.shape[1:]).contiguous()
        out = torch.empty((n, reps), dtype=torch.float32, device=self.device)
        plan = self._plan(n, data=data)
        self._call("dibs_log_joint_prob", plan.handle, nat.ptr(gs), nat.ptr(th), n, reps, nat.ptr(out), self._stream())
        return out[:, 0]

    def _theta_out(self, flat):
        return self.likelihood_model.unflatten(flat)

    def _run(self, *, key, n_particles, steps, n_dim_particles, callback, callback_every):
        """Shared body of ``sample`` (svgd.py:274-331, 730-795)."""
        import torch.distributed as dist
        n_dim = n_dim_particles or self.n_vars
        plan = self._plan(n_particles, n_dim, sharded=True)
        world, rank = plan.cfg.world_size, plan.cfg.rank
        lo, hi = plan.row0, plan.row0 + plan.n_local

        # randomly sample initial particles (svgd.py:294-295, 751-753)
        key = as_key(key)
        key, subk = split(key, 2, self.prng_partitionable)
        init = self._sample_initial_random_particles(key=subk, n_particles=n_particles, n_dim=n_dim, plan=plan)
        z_all, theta_all = init if self._joint else (init, None)
        if self.latent_prior_std is None:
            self.latent_prior_std = self._std(n_dim)          # side effect like svgd.py:301-302

        z = z_all[lo:hi].contiguous()
        theta = theta_all[lo:hi].contiguous() if self._joint else None
        v_z = torch.zeros_like(z)                              # opt_init (rmsprop second moment)
        v_theta = torch.zeros_like(theta) if self._joint else None
        sf_baseline = torch.zeros(plan.n_local, dtype=torch.float32, device=self.device)
        key_dev = keys_to_device(key, self.device)

        def gather(local):
            if world == 1:
                return local
            if dist.get_backend() != "nccl":            # gloo: collectives on host tensors
                out = torch.empty((n_particles,) + tuple(local.shape[1:]), dtype=local.dtype)
                dist.all_gather_into_tensor(out, local.contiguous().cpu())
                return out.to(local.device)
            out = torch.empty((n_particles,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(out, local.contiguous())
            return out

        callback_every = callback_every or steps
        for t in (range(0, steps, callback_every) if steps else range(0)):
            self._call("dibs_svgd_steps", plan.handle, int(t), int(callback_every), nat.ptr(z), nat.ptr(theta),
                       nat.ptr(v_z), nat.ptr(v_theta), nat.ptr(key_dev), nat.ptr(sf_baseline), self._stream())
            if callback:
                kw = dict(dibs=self, t=t + callback_every, zs=gather(z).clone())
                if self._joint:
                    kw["thetas"] = self._theta_out(gather(theta).clone())
                callback(**kw)

        z_final = gather(z)
        if getattr(self, "_pending_summary", None) is not None:
            self._flush_summary()                              # last progress line of visualize_callback
        self._last_state = dict(z=z_final, theta=gather(theta) if self._joint else None, v_z=v_z, v_theta=v_theta,
                                key=key_dev, sf_baseline=sf_baseline)
        g_final = self.particle_to_g_lim(z_final)
        if world > 1:
            torch.cuda.synchronize(self.device)
            plan.check()                                       # a timed-out peer wait surfaces here, not as a hang
        if self._joint:
            return g_final, self._theta_out(self._last_state["theta"])
        return g_final


def _log_normalise(logp):
    logp = logp.to(torch.float32)
    return logp - torch.logsumexp(logp, dim=0)


class MarginalDiBS(_SVGDBase):
    """SVGD inference of the marginal DAG posterior p(G | D) (reference: dibs/inference/svgd.py:17-375).

    Same keyword-only constructor and defaults as the reference (svgd.py:60-77)."""
    _joint = False

    def __init__(self, *, x, graph_model, likelihood_model, interv_mask=None, kernel=AdditiveFrobeniusSEKernel,
                 kernel_param=None, optimizer="rmsprop", optimizer_param=None, alpha_linear=1.0, beta_linear=1.0,
                 tau=1.0, n_grad_mc_samples=128, n_acyclicity_mc_samples=32, grad_estimator_z="score",
                 score_function_baseline=0.0, latent_prior_std=None, verbose=False, device=None,
                 prng_partitionable=False):
        if kernel_param is None:
            kernel_param = {"h": 5.0}
        if optimizer_param is None:
            optimizer_param = {"stepsize": 0.005}
        if kernel is not AdditiveFrobeniusSEKernel:
            raise NotImplementedError("dibs_b200 implements AdditiveFrobeniusSEKernel for MarginalDiBS")
        super().__init__(x=x, interv_mask=interv_mask, graph_model=graph_model, likelihood_model=likelihood_model,
                         joint=False, kernel_obj=kernel(**kernel_param), optimizer=optimizer,
                         optimizer_param=optimizer_param, alpha_linear=alpha_linear, beta_linear=beta_linear, tau=tau,
                         n_grad_mc_samples=n_grad_mc_samples, n_acyclicity_mc_samples=n_acyclicity_mc_samples,
                         grad_estimator_z=grad_estimator_z, score_function_baseline=score_function_baseline,
                         latent_prior_std=latent_prior_std, verbose=verbose, device=device,
                         prng_partitionable=prng_partitionable)

    def _parallel_update_z(self, z, kxx_unused, z_all, grad_log_prob_z):
        return self._parallel_update(z_all, None, grad_log_prob_z, None)[0]

    def get_empirical(self, g):
        """Empirical particle distribution: unique graphs weighted by their counts (svgd.py:333-351)."""
        from ..metrics import ParticleDistribution
        g_np = g.detach().cpu().numpy() if hasattr(g, "detach") else np.asarray(g)
        unique, counts = np.unique(g_np, axis=0, return_counts=True)
        logp = np.log(counts) - np.log(g_np.shape[0])
        return ParticleDistribution(logp=torch.from_numpy(logp.astype(np.float32)), g=torch.from_numpy(unique))

    def get_mixture(self, g):
        """Mixture particle distribution: graphs weighted by the unnormalised posterior log p(D | G) (svgd.py:353-375);
        the scores come from the native BGe scorer, one CTA per graph (n = M particles with one graph each)."""
        from ..metrics import ParticleDistribution
        gt = torch.as_tensor(g, device=self.device)
        logp = self.eltwise_log_joint_prob(gt.to(torch.float32)[:, None], None)[:, 0]
        return ParticleDistribution(logp=_log_normalise(logp), g=gt)

    def eltwise_log_marginal_likelihood_observ(self, g, x_ho):
        """log p(x_ho | G) for a batch of graphs g [n, d, d] on held-out observational data (svgd.py:110-111)."""
        return self._score_held_out(g, None, x_ho, None)

    def eltwise_log_marginal_likelihood_interv(self, g, x_ho, interv_msk_ho):
        """log p(x_ho | G) on held-out interventional data with its intervention mask (svgd.py:112-113)."""
        return self._score_held_out(g, None, x_ho, interv_msk_ho)

    def sample(self, *, key, n_particles, steps, n_dim_particles=None, callback=None, callback_every=None):
        """SVGD with DiBS: ``n_particles`` samples G ~ p(G | D); returns int32 [n_particles, d, d] (svgd.py:274-331)."""
        return self._run(key=key, n_particles=n_particles, steps=steps, n_dim_particles=n_dim_particles,
                         callback=callback, callback_every=callback_every)


class JointDiBS(_SVGDBase):
    """SVGD inference of the joint posterior p(G, Theta | D) (reference: dibs/inference/svgd.py:380-844).

    Same keyword-only constructor and defaults as the reference (svgd.py:425-442)."""
    _joint = True

    def __init__(self, *, x, graph_model, likelihood_model, interv_mask=None, kernel=JointAdditiveFrobeniusSEKernel,
                 kernel_param=None, optimizer="rmsprop", optimizer_param=None, alpha_linear=0.05, beta_linear=1.0,
                 tau=1.0, n_grad_mc_samples=128, n_acyclicity_mc_samples=32, grad_estimator_z="reparam",
                 score_function_baseline=0.0, latent_prior_std=None, verbose=False, device=None,
                 prng_partitionable=False):
        if kernel_param is None:
            kernel_param = {"h_latent": 5.0, "h_theta": 500.0}
        if optimizer_param is None:
            optimizer_param = {"stepsize": 0.005}
        if kernel is not JointAdditiveFrobeniusSEKernel:
            raise NotImplementedError("dibs_b200 implements JointAdditiveFrobeniusSEKernel for JointDiBS")
        super().__init__(x=x, interv_mask=interv_mask, graph_model=graph_model, likelihood_model=likelihood_model,
                         joint=True, kernel_obj=kernel(**kernel_param), optimizer=optimizer,
                         optimizer_param=optimizer_param, alpha_linear=alpha_linear, beta_linear=beta_linear, tau=tau,
                         n_grad_mc_samples=n_grad_mc_samples, n_acyclicity_mc_samples=n_acyclicity_mc_samples,
                         grad_estimator_z=grad_estimator_z, score_function_baseline=score_function_baseline,
                         latent_prior_std=latent_prior_std, verbose=verbose, device=device,
                         prng_partitionable=prng_partitionable)

    def _parallel_update_z(self, z_unused, theta_unused, kxx_unused, z_all, theta_all, grad_log_prob_z):
        zero = torch.zeros_like(self._flat_theta(theta_all))
        return self._parallel_update(z_all, theta_all, grad_log_prob_z, zero)[0]

    def _parallel_update_theta(self, z_unused, theta_unused, kxx_unused, z_all, theta_all, grad_log_prob_theta):
        zero = torch.zeros_like(self._f32(z_all))
        return self._parallel_update(z_all, theta_all, zero, grad_log_prob_theta)[1]

    def get_empirical(self, g, theta):
        """Empirical particle distribution; every (G, Theta) particle is unique, weight 1/N each (svgd.py:798-817)."""
        from ..metrics import ParticleDistribution
        n = g.shape[0]
        logp = torch.full((n,), -float(np.log(n)), dtype=torch.float32)
        return ParticleDistribution(logp=logp, g=g, theta=theta)

    def get_mixture(self, g, theta):
        """Mixture particle distribution: particles weighted by log p(Theta, D | G) (svgd.py:819-844); scored natively,
        particle i against its own Theta_i (each graph is passed twice so the launch takes the paired-sample path)."""
        from ..metrics import ParticleDistribution
        gt = torch.as_tensor(g, device=self.device).to(torch.float32)
        logp = self.eltwise_log_joint_prob(torch.stack([gt, gt], dim=1), theta)[:, 0]
        return ParticleDistribution(logp=_log_normalise(logp), g=torch.as_tensor(g, device=self.device), theta=theta)

    def eltwise_log_likelihood_observ(self, g, theta, x_ho):
        """log p(Theta, x_ho | G) per particle (g [n, d, d], theta batch) on held-out observational data
        (svgd.py:475-476 -- the reference's name notwithstanding it calls interventional_log_joint_prob)."""
        return self._score_held_out(g, theta, x_ho, None)

    def eltwise_log_likelihood_interv(self, g, theta, x_ho, interv_msk_ho):
        """log p(Theta, x_ho | G) per particle on held-out interventional data (svgd.py:477-478)."""
        return self._score_held_out(g, theta, x_ho, interv_msk_ho)

    def sample(self, *, key, n_particles, steps, n_dim_particles=None, callback=None, callback_every=None):
        """SVGD with DiBS: samples (G, Theta) ~ p(G, Theta | D); returns (int32 [M, d, d], theta pytree) (svgd.py:730-795)."""
        return self._run(key=key, n_particles=n_particles, steps=steps, n_dim_particles=n_dim_particles,
                         callback=callback, callback_every=callback_every)
