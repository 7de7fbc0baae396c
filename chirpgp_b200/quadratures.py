"""Sigma-point tables and Gauss--Hermite post-processing (host side).

Mirrors the reference interface ``chirpgp.quadratures`` (/root/reference/chirpgp/quadratures.py):

* ``SigmaPoints`` NamedTuple ``(d, n_points, w, wc, xi)``                      quadratures.py:84-110
* ``SigmaPoints.cubature(d)``                                                   quadratures.py:139-150
* ``SigmaPoints.gauss_hermite(d, order=3)``                                     quadratures.py:157-196
* ``SigmaPoints.unscented`` raises ``NotImplementedError``                      quadratures.py:153-154
* ``gaussian_expectation``                                                      quadratures.py:234-274

The tables are integer/index work plus a handful of float64 operations; "bit-exact" means they are produced
on the host by the same NumPy operations in the same order as the reference (``np.roots`` is an eigen-solve,
so e.g. the order-3 roots are *not* symmetric to the last ulp) and uploaded to the GPU as data -- nothing is
hard-coded in the kernels.  ``w``/``xi`` are NumPy float64 arrays (the reference wraps the same NumPy arrays
into ``jnp.array``).
"""
import math
from typing import NamedTuple, Optional

import numpy as np

__all__ = ['SigmaPoints', 'gaussian_expectation', 'hermite_rule_1d']


def _physicists_hermite(order: int):
    """Coefficients (highest power first) of H_0 .. H_order, H_{k+1} = 2x H_k - 2k H_{k-1}."""
    polys = [np.array([1]), np.array([2, 0])]
    for k in range(1, order):
        shifted = np.concatenate([polys[k], [0]])           # x * H_k
        lower = np.concatenate([[0, 0], polys[k - 1]])      # H_{k-1}, aligned
        polys.append(2 * shifted - 2 * k * lower)
    return polys[:order + 1]


def hermite_rule_1d(order: int):
    """1-D Gauss--Hermite nodes (roots of H_order, in the reference's order) and weights."""
    polys = _physicists_hermite(order)
    roots = np.flip(np.roots(polys[order]))
    w_1d = np.zeros((order,))
    for i in range(order):
        w_1d[i] = (2 ** (order - 1) * float(math.factorial(order)) * np.sqrt(np.pi)
                   / (order ** 2 * (np.polyval(polys[order - 1], roots[i])) ** 2))
    return roots, w_1d


def gh_index_table(d: int, order: int) -> np.ndarray:
    """(d, order**d) int64 table; column i holds the base-`order` digits of i, dimension 0 varying fastest."""
    idx = np.arange(order ** d, dtype=np.int64)
    return np.stack([(idx // (order ** j)) % order for j in range(d)], axis=0)


class SigmaPoints(NamedTuple):
    d: int
    n_points: int
    w: np.ndarray
    wc: Optional[np.ndarray]
    xi: np.ndarray

    @classmethod
    def cubature(cls, d: int):
        n_points = 2 * d
        w = np.ones((n_points,)) / n_points
        xi = math.sqrt(d) * np.concatenate([np.eye(d), -np.eye(d)], axis=0)
        return cls(d=d, n_points=n_points, w=w, wc=None, xi=xi)

    @classmethod
    def unscented(cls, d: int, alpha: float, beta: float, lam: float):
        raise NotImplementedError('Unscented transform is not implemented.')

    @classmethod
    def gauss_hermite(cls, d: int, order: int = 3):
        n_points = order ** d
        roots, w_1d = hermite_rule_1d(order)
        table = gh_index_table(d, order)
        s = 1 / (np.sqrt(np.pi) ** d)
        w = s * np.prod(w_1d[table], axis=0)
        xi = (math.sqrt(2) * roots[table]).T
        return cls(d=d, n_points=n_points, w=w, wc=None, xi=np.ascontiguousarray(xi))

    def gauss_hermite_order(self) -> int:
        """Order p if this table is bit-identical to ``gauss_hermite(d, p)`` (lets the host pick the kernel
        specialisation that shares transcendental evaluations between points), else 0."""
        d, n = int(self.d), int(self.n_points)
        for p in range(2, 12):
            if p ** d == n:
                ref = SigmaPoints.gauss_hermite(d, p)
                if np.array_equal(ref.xi, np.asarray(self.xi)) and np.array_equal(ref.w, np.asarray(self.w)):
                    return p
            if p ** d > n:
                break
        return 0

    def is_cubature(self) -> bool:
        """True if this table is bit-identical to ``cubature(d)`` (lets the host pick kernels that use its structure)."""
        ref = SigmaPoints.cubature(int(self.d))
        return (int(self.n_points) == ref.n_points and np.array_equal(ref.xi, np.asarray(self.xi))
                and np.array_equal(ref.w, np.asarray(self.w)))

    # host-side helpers with the reference's names (quadratures.py:198-231); NumPy only, never on the hot path
    def gen_sigma_points(self, m, chol_of_P):
        return np.asarray(m) + np.einsum('ij,...j->...i', np.asarray(chol_of_P), self.xi)

    def expectation(self, evals_of_integrand):
        return np.einsum('i,i...->...', self.w, np.asarray(evals_of_integrand))

    def expectation_from_nodes(self, v_f, chi):
        return np.einsum('i,i...->...', self.w, np.asarray(v_f(chi)))


def gaussian_expectation(ms, chol_Ps, func=None, d: int = 1, order: int = 10, force_shape: bool = False):
    """E[func(V)] for V ~ N(m, chol chol^T) by Gauss--Hermite (quadratures.py:234-274): the post-processing step right
    after the smoothers, e.g. ``gaussian_expectation(ms=mss[:, 2], chol_Ps=sqrt(Pss[:, 2, 2]),
    force_shape=True)`` for the frequency estimate (demos/ghfs_mle.py:87-89).

    CUDA tensors with the default integrand (``func=None`` -> the softplus ``g``) and d = 1 are evaluated on the device by
    ``cgp_gaussian_expectation_softplus_f64`` -- strided views into the smoother output are read in place, only the (T, 1)
    result exists afterwards.  Everything else (NumPy inputs, a user ``func``, d > 1) is vectorised NumPy on the host.
    """
    import torch
    from .models import g as _g
    if (func is None and d == 1 and isinstance(ms, torch.Tensor) and ms.is_cuda and isinstance(chol_Ps, torch.Tensor)
            and chol_Ps.is_cuda and ms.dtype == torch.float64 and chol_Ps.dtype == torch.float64):
        return _gaussian_expectation_device(ms, chol_Ps, order)
    func = _g if func is None else func
    if isinstance(ms, torch.Tensor):
        ms = ms.detach().cpu().numpy()
    if isinstance(chol_Ps, torch.Tensor):
        chol_Ps = chol_Ps.detach().cpu().numpy()
    ms = np.asarray(ms, dtype=np.float64)
    chol_Ps = np.asarray(chol_Ps, dtype=np.float64)
    if force_shape:
        ms = ms.reshape(-1, 1)
        chol_Ps = chol_Ps.reshape(-1, 1, 1)
    sgps = SigmaPoints.gauss_hermite(d=d, order=order)
    chi = ms[:, None, :] + np.einsum('tij,sj->tsi', chol_Ps, sgps.xi)       # (T, s, d)
    return np.einsum('s,ts...->t...', sgps.w, np.asarray(func(chi)))


def _gaussian_expectation_device(ms, chol_Ps, order: int):
    """ms, chol_Ps: CUDA float64 tensors with the same number of elements (any shape: d = 1, so (T,), (T, 1), (T, 1, 1) are
    the same thing; batches (B, T) are flattened).  Returns ms.shape + (1,) if ms has no trailing unit axis, like the
    reference's (T, 1)."""
    import ctypes as C
    import torch
    from . import _native as N
    if ms.numel() != chol_Ps.numel():
        raise ValueError('gaussian_expectation: ms has %d elements, chol_Ps %d' % (ms.numel(), chol_Ps.numel()))
    out_shape = tuple(ms.shape) if (ms.dim() >= 2 and ms.shape[-1] == 1) else tuple(ms.shape) + (1,)

    def flat(t):
        """(data pointer, element stride) of a 1-d walk over t without copying when t is an evenly strided view."""
        t = t.detach()
        t = t.reshape(-1) if t.is_contiguous() else t.squeeze()
        if t.dim() == 0:
            t = t.reshape(1)
        if t.dim() != 1 or (t.numel() > 1 and t.stride(0) < 1):
            t = t.contiguous().reshape(-1)
        return t, (int(t.stride(0)) if t.numel() > 1 else 1)

    m1, ms_stride = flat(ms)
    c1, sd_stride = flat(chol_Ps)
    n = int(m1.numel())
    sg = SigmaPoints.gauss_hermite(d=1, order=order)
    w = np.ascontiguousarray(sg.w, dtype=np.float64)
    xi = np.ascontiguousarray(np.asarray(sg.xi)[:, 0], dtype=np.float64)
    out = torch.empty((n,), dtype=torch.float64, device=ms.device)
    if n == 0:
        return out.reshape(out_shape)
    with torch.cuda.device(ms.device):
        stream = C.c_void_p(torch.cuda.current_stream(ms.device).cuda_stream)
        rc = N.lib().cgp_gaussian_expectation_softplus_f64(n, C.c_void_p(m1.data_ptr()), ms_stride, C.c_void_p(c1.data_ptr()),
                                                           sd_stride, 0, w.ctypes.data_as(C.c_void_p),
                                                           xi.ctypes.data_as(C.c_void_p), int(order),
                                                           C.c_void_p(out.data_ptr()), stream)
    N.check(rc, 'gaussian_expectation')
    return out.reshape(out_shape)
