"""Dense nonlinear-Gaussian likelihood plugin: plain descriptor the native plan consumes.

Mirrors the constructor of the reference (dibs/models/nonlinearGaussian.py:105-135).  Natively implemented:
one hidden layer with bias (the reference's default in dibs/target.py:270-271) and any of its four activations
(relu / tanh / sigmoid / leakyrelu, nonlinearGaussian.py:52-61); anything else raises
like an unknown plugin would.  The forward/backward math (nonlinearGaussian.py:248-326) runs in
dibs_b200/csrc/kernels_mc_nn.cuh, the stax initialisation (:155-186) in kernels_init.cuh.
"""


class DenseNonlinearGaussian:
    native_kind = "densenn"

    def __init__(self, *, n_vars, hidden_layers, obs_noise=0.1, sig_param=1.0, activation='relu', bias=True):
        if activation not in ('sigmoid', 'tanh', 'relu', 'leakyrelu'):
            raise KeyError(f'Invalid activation function `{activation}`')
        self.n_vars = n_vars
        self.obs_noise = obs_noise
        self.sig_param = sig_param
        self.hidden_layers = tuple(hidden_layers)
        self.activation = activation
        self.bias = bias

    def check_native(self):
        if len(self.hidden_layers) != 1 or not self.bias:
            raise NotImplementedError("dibs_b200 implements DenseNonlinearGaussian with one hidden layer and bias=True, "
                                      "any of the reference's activations (SURVEY 8f rank 3 lists the rest)")

    @property
    def hidden(self):
        return int(self.hidden_layers[0])

    def theta_dim(self):
        d, h = self.n_vars, self.hidden
        return d * (d * h + 2 * h + 1)

    def unflatten(self, flat):
        """[M, Dtheta] -> the reference's stax pytree [(W1[M,d,d,H], b1[M,d,H]), (), (W2[M,d,H,1], b2[M,d,1])]."""
        m, d, h = flat.shape[0], self.n_vars, self.hidden
        o = 0
        w1 = flat[:, o:o + d * d * h].reshape(m, d, d, h); o += d * d * h
        b1 = flat[:, o:o + d * h].reshape(m, d, h); o += d * h
        w2 = flat[:, o:o + d * h].reshape(m, d, h, 1); o += d * h
        b2 = flat[:, o:o + d].reshape(m, d, 1)
        return [(w1, b1), (), (w2, b2)]

    def flatten(self, theta):
        import torch
        (w1, b1), _, (w2, b2) = theta
        m = w1.shape[0]
        return torch.cat([w1.reshape(m, -1), b1.reshape(m, -1), w2.reshape(m, -1), b2.reshape(m, -1)], dim=1)
