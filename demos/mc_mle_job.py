"""The Monte-Carlo MLE jobs of the reference as ONE batched run (tetralith/jobs/{ekfs,ghfs,cd_ekfs,cd_ghfs}_mle.py:26-86):
`num_mcs` runs x 3 magnitudes (constant, damped, OU) = 3 num_mcs independent chirps, each with its own L-BFGS-B fit of the
hyper-parameters, then filter -> smoother -> E[g(V)] -> RMSE against the true frequency, failed fits -> NaN.

The reference runs the 300 fits one after the other; here `mle.fit_mle_batched` advances all of them in lock-step (one
batched nll / gradient launch per iteration serves every running optimiser) and the filter / smoother / read-out of all
chirps with their own fitted parameters is one call.  Measurements come from numpy's generator instead of the JAX keys in
rnd_keys.npy (no bit parity with jax.random is possible).

    python demos/mc_mle_job.py [ekfs|ghfs|cd_ekfs|cd_ghfs] [--mcs 100] [--maxiter 100] [--out results.npz]
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('method', nargs='?', default='ghfs', choices=['ekfs', 'ghfs', 'cd_ekfs', 'cd_ghfs'])
    ap.add_argument('--mcs', type=int, default=100)
    ap.add_argument('--maxiter', type=int, default=100)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    dt, T, Xi = 0.001, 3141, 0.1
    ts = np.linspace(dt, dt * T, T)
    freq, phase = toymodels.meow_freq(offset=8.)
    names = ('const', 'damped', 'ou')
    ys = np.empty((a.mcs, 3, T))
    for mc in range(a.mcs):
        rng = np.random.default_rng([2024, mc])
        noise = math.sqrt(Xi) * rng.standard_normal(T)          # the reference uses ONE noise draw per MC run for all 3 magnitudes
        for k, mag in enumerate((toymodels.constant_mag(1.), toymodels.damped_exp_mag(0.3), toymodels.random_ou_mag(1., 1., rng))):
            ys[mc, k] = toymodels.gen_chirp(ts, mag, phase) + noise
    ys = ys.reshape(-1, T)
    B = ys.shape[0]
    filt = {'ekfs': 'ekf', 'ghfs': 'sgp_filter', 'cd_ekfs': 'cd_ekf', 'cd_ghfs': 'cd_sgp_filter'}[a.method]
    sgps = cg.SigmaPoints.gauss_hermite(d=4, order=3) if 'gh' in a.method else None
    theta0 = g_inv(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
    H = np.array([0., 1., 0., 0.])
    ys_d = torch.as_tensor(ys).cuda()

    t0 = time.time()
    thetas, results = mle.fit_mle_batched(cg.build_chirp_model, theta0, H, Xi, dt, ys_d, method=filt, sgps=sgps, maxiter=a.maxiter)
    t_fit = time.time() - t0
    ok = np.array([r.success for r in results])
    params = g(np.where(ok[:, None], thetas, theta0))            # failed fits: any valid parameters, results masked below
    # one batched filter + smoother + read-out with per-chirp parameters
    drift, dispersion, m_and_cov, m0, P0, _ = cg.build_chirp_model(params)
    t0 = time.time()
    if a.method == 'ekfs':
        est, = cg.ekf_smoother(m_and_cov, H, Xi, m0, P0, dt, ys_d, readout='freq')
    elif a.method == 'ghfs':
        est, = cg.sgp_filter_smoother(m_and_cov, sgps, H, Xi, m0, P0, dt, ys_d, readout='freq')
    elif a.method == 'cd_ekfs':
        est, = cg.cd_ekf_smoother(drift, dispersion, H, Xi, m0, P0, dt, ys_d, readout='freq')
    else:
        est, = cg.cd_sgp_filter_smoother(drift, dispersion.matrix(), sgps, H, Xi, m0, P0, dt, ys_d, readout='freq')
    est = est.cpu().numpy()
    t_smooth = time.time() - t0
    rmse = np.sqrt(np.mean((est - freq(ts)[None]) ** 2, axis=1))
    rmse[~ok] = np.nan                                           # tetralith/jobs/ghfs_mle.py:75-78
    print('%s: %d chirps, %d L-BFGS-B fits in %.1f s (%d batched launches, max %d evaluations per fit), %d converged; '
          'filter + smoother + read-out of all chirps %.3f s' % (a.method, B, B, t_fit, mle.fit_mle_batched.last_launches,
                                                                 max(r.nfev for r in results), int(ok.sum()), t_smooth))
    for k, nm in enumerate(names):
        r = rmse.reshape(a.mcs, 3)[:, k]
        print('  %-7s RMSE(freq) mean %.4f  std %.4f  (%d of %d runs converged)'
              % (nm, np.nanmean(r), np.nanstd(r), int(np.isfinite(r).sum()), a.mcs))
    if a.out:
        np.savez(a.out, thetas=thetas, success=ok, rmse=rmse.reshape(a.mcs, 3), params=np.where(ok[:, None], params, np.nan))
        print('Results saved in ' + a.out)


if __name__ == '__main__':
    main()
