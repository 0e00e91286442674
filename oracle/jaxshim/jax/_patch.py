"""TEST INFRASTRUCTURE ONLY -- numpy/jax-style conveniences on ``torch.Tensor``.

Applied only inside the golden-generation process (never in ``dibs_b200``).
"""
import torch


class _AtIndexer:
    def __init__(self, t):
        self._t = t

    def __getitem__(self, idx):
        return _AtSetter(self._t, idx)


class _AtSetter:
    def __init__(self, t, idx):
        self._t, self._idx = t, idx

    def set(self, value):
        out = self._t.clone()
        out[self._idx] = value
        return out

    def add(self, value):
        out = self._t.clone()
        out[self._idx] += value
        return out


torch.Tensor.at = property(lambda self: _AtIndexer(self))
torch.Tensor.astype = lambda self, dtype: self.to(dtype)

_orig_transpose = torch.Tensor.transpose


def _transpose(self, *dims):
    if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
        return self.permute(*dims[0])
    if len(dims) == 0:
        return self.permute(*reversed(range(self.dim())))
    if len(dims) > 2:
        return self.permute(*dims)
    return _orig_transpose(self, *dims)


torch.Tensor.transpose = _transpose

_orig_reshape = torch.Tensor.reshape


def _reshape(self, *shape):
    flat = []
    for s in (shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape):
        flat.append(int(s))
    return _orig_reshape(self, tuple(flat))


torch.Tensor.reshape = _reshape
