"""TEST INFRASTRUCTURE ONLY -- ``jax.scipy.stats.norm.logpdf`` (jax/_src/scipy/stats/norm.py)."""
import math

import torch


class norm:
    @staticmethod
    def logpdf(x, loc=0.0, scale=1.0):
        scale = scale if isinstance(scale, torch.Tensor) else torch.tensor(float(scale), dtype=torch.float32)
        scale_sqrd = scale * scale
        log_normalizer = torch.log(torch.tensor(2.0 * math.pi, dtype=torch.float32) * scale_sqrd)
        quadratic = (x - loc) * (x - loc) / scale_sqrd
        return -(log_normalizer + quadratic) / 2
