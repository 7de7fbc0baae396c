"""CPU experiment behind the [G | c | C] smoother record: does regrouping the reference's recursion
    ms = mf + G (ms' - mp),  Ps = Pf + G (Ps' - Pp) G^T          (filters_smoothers.py:83-84)
as  ms = c + G ms',          Ps = C + G Ps' G^T,   c = mf - G mp,  C = Pf - G Pp G^T = Pf - W W^T  (W = D Lp^-T)
stay inside the summation-order noise floor of the reference algorithm (tests/parity_tolerances.py)?
NumPy float64, gains from the oracle's model functions; prints max |difference| of the two recursions and of the
reference-order recursion against the C oracle's smoother (sanity).  Run: python profiles/scripts/regroup_noise.py"""
import os
import sys

import numpy as np
import scipy.linalg as sl

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
from oracle import oracle as orc  # noqa: E402
from chirpgp_b200 import toymodels  # noqa: E402
from chirpgp_b200.quadratures import SigmaPoints  # noqa: E402
from parity_tolerances import NOISE_FLOOR  # noqa: E402


def gains_sgp(spec, sg, mf, Pf, dt):
    L = np.linalg.cholesky(Pf)
    chi = mf + sg.xi @ L.T
    ev = np.stack([orc.disc_mean_cov(spec, dt, x, want_jac=False)[0] for x in chi])
    Sig = orc.disc_mean_cov(spec, dt, chi[0], want_jac=False)[2]
    mp = sg.w @ ev
    Pp = np.einsum('i,ij,ik->jk', sg.w, ev, ev) + Sig - np.outer(mp, mp)
    D = np.einsum('i,ij,ik->jk', sg.w, chi, ev) - np.outer(mf, mp)
    return mp, Pp, D


def gains_ekf(spec, mf, Pf, dt):
    mp, J, Sig = orc.disc_mean_cov(spec, dt, mf, want_jac=True)
    return mp, J @ Pf @ J.T + Sig, Pf @ J.T


VARIANT = os.environ.get('CGP_REGROUP', 'WWt')      # 'GDt': C = Pf - tril(G D^T) mirrored (what the kernels do)


def run(name, gains, mfs, Pfs, ref):
    T, d = mfs.shape
    ms_a, Ps_a = mfs[-1].copy(), Pfs[-1].copy()
    ms_b, Ps_b = mfs[-1].copy(), Pfs[-1].copy()
    out = np.zeros(4)
    for k in range(T - 2, -1, -1):
        mp, Pp, D = gains(mfs[k], Pfs[k])
        cf = sl.cho_factor(Pp, lower=True)
        G = sl.cho_solve(cf, D.T).T
        ms_a = mfs[k] + G @ (ms_a - mp)
        Ps_a = Pfs[k] + G @ (Ps_a - Pp) @ G.T
        W = sl.solve_triangular(cf[0], D.T, lower=True).T          # G Pp G^T = W W^T
        c = mfs[k] - G @ mp
        Cm = Pfs[k] - (np.tril(G @ D.T) + np.tril(G @ D.T, -1).T if VARIANT == 'GDt' else W @ W.T)
        ms_b = c + G @ ms_b
        Ps_b = Cm + G @ Ps_b @ G.T
        out = np.maximum(out, [np.abs(ms_b - ms_a).max(), np.abs(Ps_b - Ps_a).max(), np.abs(ms_a - ref[0][k]).max(),
                               np.abs(Ps_a - ref[1][k]).max()])
    fl = NOISE_FLOOR[name]
    print('%-14s regrouped vs reference order: ms %.2e (floor %.2e)  Ps %.2e (floor %.2e)   | numpy reference order vs C oracle: '
          'ms %.2e  Ps %.2e' % (name, out[0], fl['ms'], out[1], fl['Ps'], out[2], out[3]))


def main():
    T, dt = 3141, 1e-3
    nch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    _, ys, _ = toymodels.synthetic_batch(24, T, dt, Xi=0.1, seed=2)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1.)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7.)
    sg = SigmaPoints.gauss_hermite(4, 3)
    f = orc.sgp_filter(spec, sg, H, 0.1, m0, P0, dt, ys[:nch])
    s = orc.sgp_smoother(spec, sg, f[0], f[1], dt)
    for b in range(nch):
        run('chirp_gh3', lambda m, P: gains_sgp(spec, sg, m, P, dt), f[0][b], f[1][b], (s[0][b], s[1][b]))
    f = orc.ekf(spec, H, 0.1, m0, P0, dt, ys[:nch])
    s = orc.eks(spec, f[0], f[1], dt)
    for b in range(nch):
        run('chirp_eks', lambda m, P: gains_ekf(spec, m, P, dt), f[0][b], f[1][b], (s[0][b], s[1][b]))
    _, ys, _ = toymodels.synthetic_batch(8, T, dt, Xi=0.1, num_harmonics=3, seed=4)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1., num_harmonics=3)
    m0, P0, H = orc.chirp_m0_P0_H(0.1, 1., 1., 7., num_harmonics=3, kind='harmonic')
    sg = SigmaPoints.cubature(8)
    f = orc.sgp_filter(spec, sg, H, 0.1, m0, P0, dt, ys[:nch])
    s = orc.sgp_smoother(spec, sg, f[0], f[1], dt)
    for b in range(nch):
        run('harmonic_cub', lambda m, P: gains_sgp(spec, sg, m, P, dt), f[0][b], f[1][b], (s[0][b], s[1][b]))


if __name__ == '__main__':
    main()
