"""TEST INFRASTRUCTURE ONLY -- ``vmap`` / ``grad`` / ``jit`` on top of ``torch.func``."""
import functools

import torch
from torch.utils import _pytree as pytree


def jit(fun=None, static_argnums=None, static_argnames=None, **_):
    """``jax.jit`` -> identity (eager torch execution)."""
    if fun is None:
        return lambda f: f
    return fun


def device_get(x):
    return x


def grad(fun, argnums=0):
    return torch.func.grad(fun, argnums=argnums)


class _Static:
    """Carrier for a non-tensor output leaf of a vmapped function."""

    def __init__(self, value):
        self.value = value


def vmap(fun, in_axes=0, out_axes=0):
    """``jax.vmap`` with positional ``in_axes`` / ``out_axes``.

    torch.func.vmap insists that outputs are tensors; the reference's stax
    initialisers return ``(output_shape, params)`` with a tuple of ints, so
    non-tensor output leaves are carried around the transform.
    """

    @functools.wraps(fun)
    def wrapped(*args):
        side = {}

        def inner(*a):
            out = fun(*a)
            leaves, spec = pytree.tree_flatten(out)
            tens, layout = [], []
            for leaf in leaves:
                if isinstance(leaf, torch.Tensor):
                    layout.append(len(tens))
                    tens.append(leaf)
                else:
                    layout.append(_Static(leaf))
            side["spec"], side["layout"] = spec, layout
            return tuple(tens)

        in_dims = in_axes if not isinstance(in_axes, list) else tuple(in_axes)
        if isinstance(in_dims, int):
            in_dims = tuple(in_dims for _ in args)
        # JAX treats None / leafless pytrees as empty nodes: any in_axis is fine for them
        in_dims = tuple(None if not any(isinstance(l, torch.Tensor) for l in pytree.tree_flatten(a)[0]) else ax
                        for a, ax in zip(args, in_dims))
        # out_axes: a single int applies to every tensor leaf (all uses in the reference)
        tens = torch.func.vmap(inner, in_dims=in_dims, out_dims=out_axes if isinstance(out_axes, int) else 0)(*args)
        if not isinstance(out_axes, int):
            flat_axes = pytree.tree_flatten(out_axes)[0]
            tens = tuple(t if ax == 0 else t.movedim(0, ax) for t, ax in zip(tens, flat_axes))
        leaves = [l.value if isinstance(l, _Static) else tens[l] for l in side["layout"]]
        return pytree.tree_unflatten(leaves, side["spec"])

    return wrapped
