"""TEST INFRASTRUCTURE ONLY -- ``jax.tree_util`` subset; ``None`` and ``()`` are empty nodes like in JAX."""
import functools

import torch


def _is_leaf(x):
    return not isinstance(x, (list, tuple, dict)) and x is not None


def tree_map(f, tree, *rest):
    if tree is None:
        return None
    if isinstance(tree, (list, tuple)):
        out = [tree_map(f, t, *[r[i] for r in rest]) for i, t in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    return f(tree, *rest)


def tree_leaves(tree):
    if tree is None:
        return []
    if isinstance(tree, (list, tuple)):
        return [l for t in tree for l in tree_leaves(t)]
    if isinstance(tree, dict):
        return [l for k in sorted(tree) for l in tree_leaves(tree[k])]
    return [tree]


def tree_reduce(f, tree, *init):
    return functools.reduce(f, tree_leaves(tree), *init)


def tree_flatten(tree):
    leaves = tree_leaves(tree)
    return leaves, tree


def tree_unflatten(treedef, leaves):
    it = iter(leaves)
    return tree_map(lambda _: next(it), treedef)
