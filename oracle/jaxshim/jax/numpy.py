"""jax.numpy subset on torch float64 (see jax/__init__.py of this shim: test infrastructure only)."""
import math as _math
import numpy as _np
import torch as _torch

ndarray = _torch.Tensor
pi = _math.pi
float64 = _torch.float64


def _is_t(x):
    return isinstance(x, _torch.Tensor)


def asarray(x, dtype=None):
    if _is_t(x):
        return x
    if isinstance(x, (list, tuple)):
        return array(x)
    a = _np.asarray(x)
    if a.dtype.kind in 'fc' or dtype is not None:
        return _torch.as_tensor(a, dtype=_torch.float64)
    if a.dtype.kind in 'iub':
        return _torch.as_tensor(a)
    return _torch.as_tensor(a, dtype=_torch.float64)


def _stack_nested(x):
    if _is_t(x):
        return x
    if isinstance(x, (list, tuple)):
        parts = [_stack_nested(e) for e in x]
        if len(parts) == 0:
            return _torch.zeros((0,), dtype=_torch.float64)
        parts = [p if _is_t(p) else _torch.as_tensor(float(p), dtype=_torch.float64) for p in parts]
        parts = [p.to(_torch.float64) for p in parts]
        return _torch.stack(parts)
    if isinstance(x, _np.ndarray):
        return asarray(x)
    return _torch.as_tensor(float(x), dtype=_torch.float64)


def array(x, dtype=None):
    """jnp.array: nested lists may mix python floats, numpy scalars and (traced) tensors."""
    if isinstance(x, _np.ndarray):
        return asarray(x)
    t = _stack_nested(x)
    return t


def zeros(shape, dtype=None):
    return _torch.zeros(shape, dtype=_torch.float64)


def ones(shape, dtype=None):
    return _torch.ones(shape, dtype=_torch.float64)


def eye(n):
    return _torch.eye(n, dtype=_torch.float64)


def zeros_like(x):
    return _torch.zeros_like(asarray(x))


def ones_like(x):
    return _torch.ones_like(asarray(x))


def empty_like(x):
    return _torch.zeros_like(asarray(x))


def diag(x):
    return _torch.diag(asarray(x))


def _un(f):
    def g(x):
        return f(asarray(x))
    return g


exp = _un(_torch.exp)
log = _un(_torch.log)
sin = _un(_torch.sin)
cos = _un(_torch.cos)
sqrt = _un(_torch.sqrt)
abs = _un(_torch.abs)


def outer(a, b):
    a, b = asarray(a), asarray(b)
    return a[:, None] * b[None, :]


def dot(a, b):
    return _torch.dot(asarray(a), asarray(b))


def einsum(spec, *ops):
    return _torch.einsum(spec, *[asarray(o) for o in ops])


def vstack(xs):
    xs = [asarray(x) for x in xs]
    xs = [x[None] if x.dim() == 1 else x for x in xs]
    return _torch.cat(xs, dim=0)


def hstack(xs):
    return _torch.cat([asarray(x) for x in xs], dim=-1)


def concatenate(xs, axis=0):
    return _torch.cat([asarray(x) for x in xs], dim=axis)


def reshape(x, shape):
    return asarray(x).reshape(shape)


def linspace(a, b, n):
    return _torch.as_tensor(_np.linspace(a, b, n))


def arange(*a):
    return _torch.as_tensor(_np.arange(*a))


def diff(x):
    x = asarray(x)
    return x[1:] - x[:-1]


def sum(x, axis=None):
    x = asarray(x)
    return x.sum() if axis is None else x.sum(dim=axis)


def mean(x, axis=None):
    x = asarray(x)
    return x.mean() if axis is None else x.mean(dim=axis)


class _Linalg:
    @staticmethod
    def cholesky(a):
        return _torch.linalg.cholesky(asarray(a))

    @staticmethod
    def solve(a, b):
        return _torch.linalg.solve(asarray(a), asarray(b))

    @staticmethod
    def inv(a):
        return _torch.linalg.inv(asarray(a))


linalg = _Linalg()
