"""Measures the summation-order noise floor of the reference algorithm on the CPU oracle and pins the figures the GPU parity
tolerances are derived from (tests/parity_tolerances.py: tolerance = 3 x floor).  The only change between the two oracle
runs is a permutation of the sigma points -- a pure re-ordering of the sums in filters_smoothers.py:119-120, :525 -- or, for
the EKS, the transposition of the (symmetric to rounding) filtering covariances.  The oracle is deterministic, so the
measured figures must reproduce the pinned ones (asserted to within [0.5, 1.05] x pinned)."""
import numpy as np

from oracle import oracle as orc
from chirpgp_b200 import toymodels
from chirpgp_b200.quadratures import SigmaPoints
from parity_tolerances import NOISE_FLOOR


class _SG:
    def __init__(self, w, xi):
        self.w, self.xi, self.n_points = w, xi, w.shape[0]


def _noise(spec, sg, m0, P0, H, ys, dt, nperm=6):
    rng = np.random.default_rng(0)
    f = orc.sgp_filter(spec, sg, H, 0.1, m0, P0, dt, ys)
    s = orc.sgp_smoother(spec, sg, f[0], f[1], dt)
    out = np.zeros(5)
    for _ in range(nperm):
        perm = rng.permutation(sg.w.shape[0])
        sp = _SG(sg.w[perm], sg.xi[perm])
        f2 = orc.sgp_filter(spec, sp, H, 0.1, m0, P0, dt, ys)
        s2 = orc.sgp_smoother(spec, sp, f[0], f[1], dt)
        out = np.maximum(out, [np.abs(f2[0] - f[0]).max(), np.abs(f2[1] - f[1]).max(), np.abs(s2[0] - s[0]).max(),
                               np.abs(s2[1] - s[1]).max(), np.abs((f2[2] - f[2]) / f[2]).max()])
    return dict(zip(('mf', 'Pf', 'ms', 'Ps', 'nll'), out))


def _assert_pinned(name, measured):
    pinned = NOISE_FLOOR[name]
    print('%s noise floor: %s' % (name, '  '.join('%s %.2e (pinned %.2e)' % (k, measured[k], pinned[k]) for k in pinned)))
    for k, v in pinned.items():
        assert 0.5 * v <= measured[k] <= 1.05 * v, (name, k, measured[k], v)


def test_noise_floor_chirp_gh3():
    T, dt = 3141, 1e-3
    _, ys, _ = toymodels.synthetic_batch(24, T, dt, Xi=0.1, seed=2)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1.)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7.)
    _assert_pinned('chirp_gh3', _noise(spec, SigmaPoints.gauss_hermite(4, 3), m0, P0, H, ys, dt))


def test_noise_floor_harmonic_cubature():
    T, dt = 3141, 1e-3
    _, ys, _ = toymodels.synthetic_batch(8, T, dt, Xi=0.1, num_harmonics=3, seed=4)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1., num_harmonics=3)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7., num_harmonics=3, kind='harmonic')
    _assert_pinned('harmonic_cub', _noise(spec, SigmaPoints.cubature(8), m0, P0, H, ys, dt))


def test_noise_floor_long_sequence():
    B, T, dt, Xi = 6, 20000, 1.5e-4, 0.1
    rng = np.random.default_rng(0)
    ts = np.linspace(dt, dt * T, T)
    ys = np.sin(2 * np.pi * (500 * np.exp(-5 / np.sin(ts)) + 8 * ts))[None] + np.sqrt(Xi) * rng.standard_normal((B, T))
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1.)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7.)
    _assert_pinned('chirp_gh3_T20000', _noise(spec, SigmaPoints.gauss_hermite(4, 3), m0, P0, H, ys[:2], dt))


def test_noise_floor_eks_triangle():
    T, dt = 3141, 1e-3
    _, ys, _ = toymodels.synthetic_batch(24, T, dt, Xi=0.1, seed=2)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1.)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7.)
    f = orc.ekf(spec, H, 0.1, m0, P0, dt, ys)
    s = orc.eks(spec, f[0], f[1], dt)
    PT = np.ascontiguousarray(np.swapaxes(f[1], -1, -2))
    assert 0 < np.abs(PT - f[1]).max() < 1e-14                 # symmetric to rounding only
    s2 = orc.eks(spec, f[0], PT, dt)
    _assert_pinned('chirp_eks', dict(ms=np.abs(s2[0] - s[0]).max(), Ps=np.abs(s2[1] - s[1]).max()))
