"""Minimal torch-float64-backed stand-in for the parts of JAX that chirpgp's hot path touches.

TEST INFRASTRUCTURE ONLY (lives under oracle/): it exists so that the UNMODIFIED reference sources
(/root/reference/chirpgp/{filters_smoothers,quadratures,models}.py) can be imported and executed in this
container, where jax/jaxlib are not installed, to (a) generate the golden vectors committed under
tests/golden/ and (b) validate the C restatement in oracle/.  It restates JAX *primitives* (scan, cond,
jacfwd, vmap, cholesky, cho_solve, block_diag, norm.logpdf ...), not chirpgp's algorithm.  Nothing in the
product package may import it.
"""
import torch as _torch
from . import numpy, lax, random, scipy  # noqa: F401

_torch.set_default_dtype(_torch.float64)


class _Config:
    def update(self, *_a, **_k):
        pass


config = _Config()


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def _wrap_in(args):
    return tuple(numpy.asarray(a) if not callable(a) else a for a in args)


def vmap(fun, in_axes=0, out_axes=0):
    """jax.vmap -> torch.func.vmap (in_axes may be int / list / tuple, None = broadcast)."""
    if isinstance(in_axes, list):
        in_axes = tuple(in_axes)

    def wrapped(*args):
        return _torch.func.vmap(fun, in_dims=in_axes, out_dims=out_axes)(*args)

    return wrapped


def jacfwd(fun, argnums=0):
    def wrapped(*args):
        args = list(args)
        args[argnums] = numpy.asarray(args[argnums])
        return _torch.func.jacfwd(fun, argnums=argnums)(*args)

    return wrapped


def jacrev(fun, argnums=0):
    def wrapped(*args):
        args = list(args)
        args[argnums] = numpy.asarray(args[argnums])
        return _torch.func.jacrev(fun, argnums=argnums)(*args)

    return wrapped


def grad(fun, argnums=0):
    def wrapped(*args):
        args = list(args)
        args[argnums] = numpy.asarray(args[argnums])
        return _torch.func.grad(fun, argnums=argnums)(*args)

    return wrapped


def value_and_grad(fun, argnums=0):
    def wrapped(*args):
        args = list(args)
        args[argnums] = numpy.asarray(args[argnums])
        g, v = _torch.func.grad(lambda *a: (lambda r: (r, r))(fun(*a)), argnums=argnums, has_aux=True)(*args)
        return v, g

    return wrapped


def hessian(fun, argnums=0):
    return jacfwd(jacrev(fun, argnums=argnums), argnums=argnums)


def cond(pred, true_fun, false_fun, *operands):
    return lax.cond(pred, true_fun, false_fun, *operands)
