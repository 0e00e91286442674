"""Linear-Gaussian likelihood plugins: plain descriptors the native plan consumes.

Mirror the constructors and defaults of the reference (dibs/models/linearGaussian.py:35-57, 190-197).
The jit-pure methods DiBS binds (``interventional_log_marginal_prob`` :150-170,
``interventional_log_joint_prob`` :323-338) and their gradients run inside the CUDA Monte-Carlo kernels
(dibs_b200/csrc/kernels_mc_bge.cuh, kernels_mc.cuh); ``sample_parameters`` (:212-227) runs in
dibs_b200/csrc/kernels_init.cuh.
"""


class BGe:
    """BGe marginal likelihood log p(D | G) with Normal-Wishart prior (Geiger & Heckerman; Kuipers et al.)."""
    native_kind = "bge"

    def __init__(self, *, n_vars, mean_obs=None, alpha_mu=None, alpha_lambd=None):
        self.n_vars = n_vars
        self.mean_obs = mean_obs
        self.alpha_mu = alpha_mu or 1.0
        self.alpha_lambd = alpha_lambd or (self.n_vars + 2)
        assert self.alpha_lambd > self.n_vars + 1

    def get_theta_shape(self, *, n_vars):
        raise NotImplementedError("Not available for BGe score; use `LinearGaussian` model instead.")

    def sample_parameters(self, *, key, n_vars, n_particles=0, batch_size=0):
        raise NotImplementedError("Not available for BGe score; use `LinearGaussian` model instead.")

    def theta_dim(self):
        return 0


class LinearGaussian:
    """Linear SEM with additive Gaussian noise and Gaussian edge weights."""
    native_kind = "lingauss"

    def __init__(self, *, n_vars, obs_noise=0.1, mean_edge=0.0, sig_edge=1.0, min_edge=0.5):
        self.n_vars = n_vars
        self.obs_noise = obs_noise
        self.mean_edge = mean_edge
        self.sig_edge = sig_edge
        self.min_edge = min_edge

    def get_theta_shape(self, *, n_vars):
        return (n_vars, n_vars)

    def theta_dim(self):
        return self.n_vars * self.n_vars

    def unflatten(self, flat):
        """[M, d*d] -> [M, d, d] (the reference's theta layout)."""
        return flat.reshape(flat.shape[0], self.n_vars, self.n_vars)

    def flatten(self, theta):
        return theta.reshape(theta.shape[0], -1)
