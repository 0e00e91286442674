"""TEST INFRASTRUCTURE ONLY -- ``jax.lax`` subset (fori_loop, cond)."""
import torch


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        # the loop index is a traced int32 scalar in JAX, so `alpha_linear * t` is an fp32 product
        val = body_fun(torch.tensor(i, dtype=torch.int32), val)
    return val


def cond(pred, true_fun, false_fun, *operands, operand=None):
    if isinstance(pred, torch.Tensor):
        pred = bool(pred)
    if operands:
        return true_fun(*operands) if pred else false_fun(*operands)
    return true_fun(operand) if pred else false_fun(operand)
