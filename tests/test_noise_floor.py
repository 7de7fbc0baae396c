"""Measures the summation-order noise floor of the reference algorithm on the CPU oracle: the only change is a
permutation of the sigma points (a pure re-ordering of the sums in filters_smoothers.py:119-120, :525).  The GPU
parity tolerances in tests/test_gpu_parity.py are set ~10x above these figures."""
import numpy as np

from oracle import oracle as orc
from chirpgp_b200 import toymodels
from chirpgp_b200.quadratures import SigmaPoints


class _SG:
    def __init__(self, w, xi):
        self.w, self.xi, self.n_points = w, xi, w.shape[0]


def _noise(spec, sg, m0, P0, H, ys, dt):
    rng = np.random.default_rng(0)
    f = orc.sgp_filter(spec, sg, H, 0.1, m0, P0, dt, ys)
    s = orc.sgp_smoother(spec, sg, f[0], f[1], dt)
    perm = rng.permutation(sg.w.shape[0])
    sp = _SG(sg.w[perm], sg.xi[perm])
    f2 = orc.sgp_filter(spec, sp, H, 0.1, m0, P0, dt, ys)
    s2 = orc.sgp_smoother(spec, sp, f[0], f[1], dt)
    return (np.abs(f2[0] - f[0]).max(), np.abs(f2[1] - f[1]).max(), np.abs((f2[2] - f[2]) / f[2]).max(),
            np.abs(s2[0] - s[0]).max(), np.abs(s2[1] - s[1]).max())


def test_noise_floor_chirp_gh3():
    T, dt = 3141, 1e-3
    _, ys, _ = toymodels.synthetic_batch(3, T, dt, Xi=0.1, seed=2)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1.)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7.)
    mf, Pf, nll, ms, Ps = _noise(spec, SigmaPoints.gauss_hermite(4, 3), m0, P0, H, ys, dt)
    print('chirp d=4 GH3 noise floor: mf %.1e Pf %.1e nll %.1e ms %.1e Ps %.1e' % (mf, Pf, nll, ms, Ps))
    # non-zero (the algorithm is not summation-order invariant) but far below the stated tolerances
    assert 0 < mf < 5e-9 and ms < 5e-9 and Pf < 5e-9 and Ps < 5e-9 and nll < 1e-10


def test_noise_floor_harmonic_cubature():
    T, dt = 3141, 1e-3
    _, ys, _ = toymodels.synthetic_batch(3, T, dt, Xi=0.1, num_harmonics=3, seed=4)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1., num_harmonics=3)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7., num_harmonics=3, kind='harmonic')
    mf, Pf, nll, ms, Ps = _noise(spec, SigmaPoints.cubature(8), m0, P0, H, ys, dt)
    print('harmonic d=8 cubature noise floor: mf %.1e Pf %.1e nll %.1e ms %.1e Ps %.1e' % (mf, Pf, nll, ms, Ps))
    assert 0 < mf < 2e-8 and ms < 2e-8 and Pf < 2e-8 and Ps < 2e-8 and nll < 1e-10
