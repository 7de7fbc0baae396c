"""jax.lax subset: scan as a Python loop over the leading axis, cond as a Python branch."""
import torch as _torch
from . import numpy as _jnp


def _leaves_len(xs):
    if isinstance(xs, (tuple, list)):
        return _leaves_len(xs[0])
    return xs.shape[0]


def _index(xs, i):
    if isinstance(xs, (tuple, list)):
        return tuple(_index(x, i) for x in xs)
    return xs[i]


def _stack(ys):
    first = ys[0]
    if isinstance(first, (tuple, list)):
        return tuple(_stack([y[k] for y in ys]) for k in range(len(first)))
    ys = [y if isinstance(y, _torch.Tensor) else _torch.as_tensor(float(y)) for y in ys]
    return _torch.stack(ys)


def _empty_like_out(f, init, xs):
    raise NotImplementedError('scan over an empty sequence: shapes unknown to the shim')


def scan(f, init, xs, length=None, reverse=False):
    if isinstance(xs, (tuple, list)):
        xs = tuple(_jnp.asarray(x) for x in xs)
    else:
        xs = _jnp.asarray(xs)
    n = _leaves_len(xs)
    carry = init
    ys = []
    order = range(n - 1, -1, -1) if reverse else range(n)
    for i in order:
        carry, y = f(carry, _index(xs, i))
        ys.append(y)
    if reverse:
        ys = ys[::-1]
    if n == 0:
        return carry, None
    return carry, _stack(ys)


def cond(pred, true_fun, false_fun, *operands):
    if bool(pred):
        return true_fun(*operands)
    return false_fun(*operands)
