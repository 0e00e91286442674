"""Kernel plugins: descriptors of the squared-exponential kernels (dibs/kernel.py:4-71).

The all-pairs evaluation and its gradient run in dibs_b200/csrc/kernels_pair.cuh; ``eval`` here is the
single-pair convenience of the reference's interface and is not on the SVGD path.
"""
import torch


class AdditiveFrobeniusSEKernel:
    """k(Z, Z') = scale * exp(-||Z - Z'||_F^2 / h)"""

    def __init__(self, *, h=20.0, scale=1.0):
        self.h = h
        self.scale = scale

    def eval(self, *, x, y):
        return self.scale * torch.exp(-torch.sum((x - y) ** 2.0) / self.h)


class JointAdditiveFrobeniusSEKernel:
    """k = scale_z exp(-||Z-Z'||^2 / h_z) + scale_theta exp(-||Theta-Theta'||^2 / h_theta)"""

    def __init__(self, *, h_latent=5.0, h_theta=500.0, scale_latent=1.0, scale_theta=1.0):
        self.h_latent = h_latent
        self.h_theta = h_theta
        self.scale_latent = scale_latent
        self.scale_theta = scale_theta

    def eval(self, *, x_latent, x_theta, y_latent, y_theta):
        def sqn(a, b):
            if isinstance(a, torch.Tensor):
                return torch.sum((a - b) ** 2.0)
            return sum(sqn(u, v) for u, v in zip(a, b))
        return (self.scale_latent * torch.exp(-sqn(x_latent, y_latent) / self.h_latent)
                + self.scale_theta * torch.exp(-sqn(x_theta, y_theta) / self.h_theta))
