"""TEST INFRASTRUCTURE ONLY -- ``jax.example_libraries.stax`` subset (Dense, serial, elementwise activations)."""
import torch

from .. import random
from ..nn.initializers import normal as _normal_init


def Dense(out_dim, W_init=None, b_init=None):
    W_init = W_init or _normal_init()
    b_init = b_init or _normal_init()

    def init_fun(rng, input_shape):
        output_shape = tuple(input_shape[:-1]) + (out_dim,)
        k1, k2 = random.split(rng)
        W, b = W_init(k1, (input_shape[-1], out_dim)), b_init(k2, (out_dim,))
        return output_shape, (W, b)

    def apply_fun(params, inputs, **kwargs):
        W, b = params
        return inputs @ W + b

    return init_fun, apply_fun


def elementwise(fun, **fun_kwargs):
    init_fun = lambda rng, input_shape: (input_shape, ())
    apply_fun = lambda params, inputs, **kwargs: fun(inputs, **fun_kwargs)
    return init_fun, apply_fun


Relu = elementwise(torch.relu)
Tanh = elementwise(torch.tanh)
Sigmoid = elementwise(torch.sigmoid)
LeakyRelu = elementwise(lambda x: torch.where(x >= 0, x, 0.01 * x))


def serial(*layers):
    nlayers = len(layers)
    init_funs, apply_funs = zip(*layers)

    def init_fun(rng, input_shape):
        params = []
        for init in init_funs:
            rng, layer_rng = random.split(rng)
            input_shape, param = init(layer_rng, input_shape)
            params.append(param)
        return input_shape, params

    def apply_fun(params, inputs, **kwargs):
        for fun, param in zip(apply_funs, params):
            inputs = fun(param, inputs, **kwargs)
        return inputs

    return init_fun, apply_fun
