"""Chirp instantaneous-frequency estimation, end to end on the GPU -- the flow of the reference's five demos
(/root/reference/demos/{ekfs_mle,ghfs_mle,cd_ekfs_mle,cd_ghfs_mle,ghfs_harmonics_mle}.py) without the plotting:
MLE of the hyper-parameters (L-BFGS-B) -> filter -> smoother -> E[g(V)] by Gauss-Hermite -> RMSE against the truth.

    python demos/if_estimation.py [ekfs|ghfs|cd_ekfs|cd_ghfs|ghfs_harmonics] [--no-mle]
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chirpgp_b200 as cg  # noqa: E402
from chirpgp_b200 import mle, toymodels  # noqa: E402
from chirpgp_b200.models import g, g_inv  # noqa: E402


def rmse(a, b):                                   # chirpgp/tools.py:279-293
    return float(np.sqrt(np.mean((np.asarray(a) - np.asarray(b)) ** 2)))


def main():
    method = next((a for a in sys.argv[1:] if not a.startswith('--')), 'ghfs')
    do_mle = '--no-mle' not in sys.argv
    dt, T, Xi = 0.001, 3141, 0.1                  # demos/ekfs_mle.py:16-18
    ts = np.linspace(dt, dt * T, T)
    freq, phase = toymodels.meow_freq(offset=8.)
    rng = np.random.default_rng(555)
    h = 3 if method == 'ghfs_harmonics' else 1
    if h == 1:
        build = cg.build_chirp_model
        sgps = cg.SigmaPoints.gauss_hermite(d=4, order=3)
    else:
        build = lambda p: cg.build_harmonic_chirp_model(p, num_harmonics=h)
        sgps = cg.SigmaPoints.cubature(d=2 * h + 2)
    for name, mag in [('constant', toymodels.constant_mag(1.)), ('damped', toymodels.damped_exp_mag(0.3)),
                      ('OU', toymodels.random_ou_mag(1., 1., rng))]:
        clean = toymodels.gen_chirp(ts, mag, phase) if h == 1 else toymodels.gen_harmonic_chirp(ts, [mag] * h, phase)
        ys = clean + math.sqrt(Xi) * rng.standard_normal(T)
        theta = g_inv(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
        H = build(g(theta))[5]
        t0 = time.time()
        filt = {'ekfs': 'ekf', 'ghfs': 'sgp_filter', 'cd_ekfs': 'cd_ekf', 'cd_ghfs': 'cd_sgp_filter',
                'ghfs_harmonics': 'sgp_filter'}[method]
        if do_mle:
            theta, res = mle.fit_mle(build, theta, H, Xi, dt, ys, method=filt, sgps=sgps if 'sgp' in filt else None, maxiter=100)
            ok = bool(res.success)
        else:
            ok = True
        drift, dispersion, m_and_cov, m0, P0, H = build(g(theta))
        if method == 'ekfs':
            f = cg.ekf(m_and_cov, H, Xi, m0, P0, dt, ys)
            s = cg.eks(m_and_cov, f[0], f[1], dt)
        elif method in ('ghfs', 'ghfs_harmonics'):
            f = cg.sgp_filter(m_and_cov, sgps, H, Xi, m0, P0, dt, ys)
            s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], dt)
        elif method == 'cd_ekfs':
            f = cg.cd_ekf(drift, dispersion, H, Xi, m0, P0, dt, ys)
            s = cg.cd_eks(drift, dispersion, f[0], f[1], dt)
        else:
            bm = dispersion(np.eye(4))
            f = cg.cd_sgp_filter(drift, bm, sgps, H, Xi, m0, P0, dt, ys)
            s = cg.cd_sgp_smoother(drift, bm, sgps, f[0], f[1], dt)
        v = 2 * h                                  # index of the V state
        est = cg.gaussian_expectation(ms=s[0][:, v], chol_Ps=np.sqrt(s[1][:, v, v]), func=g, force_shape=True)[:, 0]
        print('%-14s %-9s MLE ok=%s params=%s nll=%.3f  RMSE(freq)=%.4f  (%.2f s)'
              % (method, name, ok, np.round(g(theta), 4), f[2][-1], rmse(freq(ts), est), time.time() - t0), flush=True)


if __name__ == '__main__':
    main()
