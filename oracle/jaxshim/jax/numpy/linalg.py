"""TEST INFRASTRUCTURE ONLY -- ``jnp.linalg`` subset (matrix_power, slogdet)."""
import torch


def matrix_power(a, n):
    """Same multiplication order as ``jnp.linalg.matrix_power`` (LSB-first binary exponentiation)."""
    n = int(n)
    if n == 0:
        return torch.eye(a.shape[-1], dtype=a.dtype).expand_as(a).clone()
    if n == 1:
        return a
    if n == 2:
        return a @ a
    if n == 3:
        return (a @ a) @ a
    z = result = None
    while n > 0:
        z = a if z is None else z @ z
        n, bit = divmod(n, 2)
        if bit:
            result = z if result is None else result @ z
    return result


def slogdet(a):
    """Partial-pivot LU log-determinant (LAPACK getrf), like XLA:CPU."""
    out = torch.linalg.slogdet(a)
    return out.sign, out.logabsdet
