"""Chirp instantaneous-frequency estimation, end to end on the GPU -- the flow of the reference's five demos
(/root/reference/demos/{ekfs_mle,ghfs_mle,cd_ekfs_mle,cd_ghfs_mle,ghfs_harmonics_mle}.py, plus the KPT pipeline of
tetralith/jobs/kpt_mle.py) without the plotting:
MLE of the hyper-parameters (L-BFGS-B) -> filter -> smoother -> E[g(V)] by Gauss-Hermite -> RMSE against the truth.
The measurements go to the GPU once; filtering, smoothing and the frequency estimate stay on the device.

    python demos/if_estimation.py [ekfs|ghfs|cd_ekfs|cd_ghfs|ghfs_harmonics|kpt] [--no-mle]
"""
import math
import os
import sys
import time

import numpy as np
import torch

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
        ys = torch.as_tensor(clean + math.sqrt(Xi) * rng.standard_normal(T)).cuda()
        t0 = time.time()
        if method == 'kpt':                        # tetralith/jobs/kpt_mle.py:37-70
            fsamp = 1. / dt
            theta = g_inv(np.array([0.02, 1e-5, 1e-5, 8., 1.]))
            build_kpt = lambda p: cg.build_kpt_chirp_model(p, fsamp, num_harmonics=1)
            ok = True
            if do_mle:
                theta, res = mle.fit_mle(build_kpt, theta, None, Xi, dt, ys, method='ekf_for_kpt', maxiter=100)
                ok = bool(res.success)
            F, Sigma, m0, P0, hfun = build_kpt(g(theta))
            f = cg.ekf_for_kpt(F, Sigma, hfun, Xi, m0, P0, dt, ys)
            s = cg.rts(F, Sigma, f[0], f[1])
            scale = fsamp / 2 / math.pi
            est = cg.gaussian_expectation(ms=s[0][:, 0] * scale, chol_Ps=torch.sqrt(s[1][:, 0, 0]) * scale, force_shape=True)
        else:
            theta = g_inv(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
            H = build(g(theta))[5]
            filt = {'ekfs': 'ekf', 'ghfs': 'sgp_filter', 'cd_ekfs': 'cd_ekf', 'cd_ghfs': 'cd_sgp_filter',
                    'ghfs_harmonics': 'sgp_filter'}[method]
            if do_mle:
                theta, res = mle.fit_mle(build, theta, H, Xi, dt, ys, method=filt, sgps=sgps if 'sgp' in filt else None,
                                         maxiter=100)
                ok = bool(res.success)
            else:
                ok = True
            drift, dispersion, m_and_cov, m0, P0, H = build(g(theta))
            if method == 'ekfs':
                f = cg.ekf(m_and_cov, H, Xi, m0, P0, dt, ys)
                s = cg.eks(m_and_cov, f[0], f[1], dt)
            elif method in ('ghfs', 'ghfs_harmonics'):
                f = cg.sgp_filter(m_and_cov, sgps, H, Xi, m0, P0, dt, ys)        # leaves the smoother gains on f[0]
                s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], dt)             # -> sweep only
            elif method == 'cd_ekfs':
                f = cg.cd_ekf(drift, dispersion, H, Xi, m0, P0, dt, ys)
                s = cg.cd_eks(drift, dispersion, f[0], f[1], dt)
            else:
                bm = dispersion(np.eye(4))
                f = cg.cd_sgp_filter(drift, bm, sgps, H, Xi, m0, P0, dt, ys)
                s = cg.cd_sgp_smoother(drift, bm, sgps, f[0], f[1], dt)
            v = 2 * h                                  # index of the V state
            est = cg.gaussian_expectation(ms=s[0][:, v], chol_Ps=torch.sqrt(s[1][:, v, v]), force_shape=True)   # on the device
        est = est[:, 0].cpu().numpy()
        print('%-14s %-9s MLE ok=%s params=%s nll=%.3f  RMSE(freq)=%.4f  (%.2f s)'
              % (method, name, ok, np.round(g(theta), 5), float(f[2][-1]), rmse(freq(ts), est), time.time() - t0), flush=True)


if __name__ == '__main__':
    main()
