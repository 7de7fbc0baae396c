"""jax.scipy.linalg subset. cholesky / cho_factor read ONE triangle only (LAPACK potrf semantics)."""
import torch as _torch
from .. import numpy as _jnp


def _sym_from_lower(a):
    lo = _torch.tril(a)
    return lo + _torch.tril(a, -1).transpose(-1, -2)


def cholesky(a, lower=False):
    a = _jnp.asarray(a)
    if lower:
        # potrf('L') reads the lower triangle
        return _torch.linalg.cholesky(_sym_from_lower(a))
    up = _torch.triu(a)
    s = up + _torch.triu(a, 1).transpose(-1, -2)
    return _torch.linalg.cholesky(s).transpose(-1, -2)


def cho_factor(a, lower=False):
    return cholesky(a, lower=lower), lower


def cho_solve(c_and_lower, b):
    c, lower = c_and_lower
    b = _jnp.asarray(b)
    vec = b.dim() == 1
    if vec:
        b = b[:, None]
    x = _torch.cholesky_solve(b, c, upper=not lower)
    return x[:, 0] if vec else x


def block_diag(*arrs):
    """Built from cat of zero-padded row blocks so that it also works under torch.func.vmap / jacfwd."""
    ts = []
    for a in arrs:
        a = _jnp.asarray(a)
        if a.dim() == 0:
            a = a.reshape(1, 1)
        elif a.dim() == 1:
            a = a[None, :]
        ts.append(a.to(_torch.float64))
    ncols = sum(t.shape[1] for t in ts)
    rows = []
    c0 = 0
    for t in ts:
        r, c = t.shape
        parts = []
        if c0 > 0:
            parts.append(_torch.zeros((r, c0), dtype=_torch.float64))
        parts.append(t)
        if ncols - c0 - c > 0:
            parts.append(_torch.zeros((r, ncols - c0 - c), dtype=_torch.float64))
        rows.append(_torch.cat(parts, dim=1))
        c0 += c
    return _torch.cat(rows, dim=0)


def expm(a):
    return _torch.linalg.matrix_exp(_jnp.asarray(a))
