"""TEST INFRASTRUCTURE ONLY -- ``jax.example_libraries.optimizers.{sgd,rmsprop}`` (leaf-wise over pytrees)."""
import torch

from ..tree_util import tree_map


class OptimizerState:
    def __init__(self, tree):
        self.tree = tree  # pytree whose leaves are tuples (x, *aux)


def _optimizer(init_leaf, update_leaf, params_leaf):
    def _is_state_leaf(v):
        return isinstance(v, tuple) and len(v) and isinstance(v[0], torch.Tensor) and getattr(v, "_state", True)

    class _S(tuple):
        pass

    def init(x0_tree):
        return OptimizerState(tree_map(lambda x: _S(init_leaf(x)), x0_tree))

    def _map_states(f, tree, *rest):
        if isinstance(tree, _S):
            return f(tree, *rest)
        if tree is None:
            return None
        if isinstance(tree, (list, tuple)):
            return type(tree)(_map_states(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
        raise TypeError(type(tree))

    def update(i, g_tree, state):
        def _upd(s, g):
            return _S(update_leaf(i, g, tuple(s)))
        # walk g_tree and state.tree together
        def walk(st, g):
            if isinstance(st, _S):
                return _upd(st, g)
            if st is None:
                return None
            return type(st)(walk(a, b) for a, b in zip(st, g))
        return OptimizerState(walk(state.tree, g_tree))

    def get_params(state):
        return _map_states(lambda s: params_leaf(tuple(s)), state.tree)

    return init, update, get_params


def sgd(step_size):
    return _optimizer(lambda x0: (x0,),
                      lambda i, g, s: (s[0] - step_size * g,),
                      lambda s: s[0])


def rmsprop(step_size, gamma=0.9, eps=1e-8):
    def init(x0):
        return x0, torch.zeros_like(x0)

    def update(i, g, state):
        x, avg_sq_grad = state
        avg_sq_grad = avg_sq_grad * gamma + (g * g) * (1. - gamma)
        x = x - step_size * g / torch.sqrt(avg_sq_grad + eps)
        return x, avg_sq_grad

    return _optimizer(init, update, lambda s: s[0])
