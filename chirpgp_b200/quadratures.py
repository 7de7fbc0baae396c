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
    """E[func(V)] for V ~ N(m, chol chol^T) by Gauss--Hermite (quadratures.py:234-274).

    Post-processing step *after* the hot path: vectorised NumPy on the host.
    """
    from .models import g as _g
    func = _g if func is None else func
    ms = np.asarray(ms, dtype=np.float64)
    chol_Ps = np.asarray(chol_Ps, dtype=np.float64)
    if force_shape:
        ms = ms.reshape(-1, 1)
        chol_Ps = chol_Ps.reshape(-1, 1, 1)
    sgps = SigmaPoints.gauss_hermite(d=d, order=order)
    chi = ms[:, None, :] + np.einsum('tij,sj->tsi', chol_Ps, sgps.xi)       # (T, s, d)
    return np.einsum('s,ts...->t...', sgps.w, np.asarray(func(chi)))
