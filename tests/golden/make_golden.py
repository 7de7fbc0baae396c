"""Generate the golden vectors in tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE SOURCES.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

jax/jaxlib are not installed, so the reference's hot-path modules (chirpgp/filters_smoothers.py,
quadratures.py, models.py) are imported over oracle/jaxshim, a torch-float64 stand-in for the JAX primitives
they use (scan, cond, vmap, jacfwd, grad, cholesky, cho_solve, block_diag, norm.logpdf).  Everything stored here
is therefore "reference algorithm, reference source code, torch rounding" -- not real-XLA rounding.

Inputs are stored next to outputs so the fixtures are self-contained.  Cases:
  linear_a{0,1}.npz  the reference's own test data (test/test_filters_smoothers.py:19-85, np.random.seed(666))
  chirp.npz          chirp model d=4 (models.py:437-459): all 8 nonlinear functions, GH order 3 + cubature,
                     nll gradients w.r.t. theta (jax.grad of filter(...)[-1][-1], demos/ekfs_mle.py:42-45)
  chirp_lam0.npz     exact lam == 0 branch of disc_chirp_lcd (models.py:302-303)
  harmonic.npz       harmonic model h=3, d=8 with cubature (demos/ghfs_harmonics_mle.py:25-27)
  lascala.npz        La Scala model (models.py:497-519)
  short.npz          T = 2 warm-up call (demos/ekfs_mle.py:65-66)
  tables.npz         sigma-point tables (bit-exact targets)
  kpt.npz, kpt_h2.npz  ekf_for_kpt + rts on the KPT model (models.py:522-580), 1 and 2 harmonics, with jax.grad of the nll
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

fs, qd, md = ref_loader.load()
import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402

sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location('_toy', os.path.join(ROOT, 'chirpgp_b200', 'toymodels.py'))
toy = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(toy)

OUT = os.path.dirname(os.path.abspath(__file__))


def npy(x):
    if isinstance(x, (tuple, list)):
        return [npy(e) for e in x]
    return np.asarray(x.detach().numpy() if hasattr(x, 'detach') else x, dtype=np.float64)


def save(name, **kw):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in kw.items()})
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def linear_cases():
    np.random.seed(666)
    for idx, (a, b) in enumerate(([1., 1.], [2.1, 0.4])):
        dim_x, dt = 3, 0.01
        A = -a * jnp.eye(dim_x)
        B = b * jnp.eye(dim_x)
        drift = lambda u: A @ u
        dispersion = lambda _: B
        F = math.exp(-a * dt) * jnp.eye(dim_x)
        Sigma = b ** 2 / (2 * a) * (1 - math.exp(-2 * a * dt)) * jnp.eye(dim_x)
        Xi = 0.1
        H = jnp.ones((dim_x,))
        m0 = jnp.zeros((dim_x,))
        P0 = 0.1 * jnp.eye(dim_x)
        num_measurements = 1000
        yy = np.zeros((num_measurements,))
        x = np.array(m0).copy()
        for i in range(num_measurements):     # identical draw order to the reference test (:43-54)
            x = npy(F) @ x + np.sqrt(npy(Sigma)) @ np.random.randn(dim_x)
            yy[i] = npy(H) @ x + np.sqrt(Xi) * np.random.randn()
        T = 300                                # fixtures keep the first 300 samples of the 1000 drawn
        ys = jnp.asarray(yy[:T])
        m_and_cov = lambda u, _: (F @ u, Sigma)
        sg = qd.SigmaPoints.gauss_hermite(d=dim_x, order=4)
        out = dict(a=a, b=b, dt=dt, Xi=Xi, ys=npy(ys), F=npy(F), Sigma=npy(Sigma), A=npy(A), B=npy(B), H=npy(H),
                   m0=npy(m0), P0=npy(P0), sg_w=npy(sg.w), sg_xi=npy(sg.xi))
        r = {}
        r['kf'] = fs.kf(F, Sigma, H, Xi, m0, P0, ys)
        r['ekf'] = fs.ekf(m_and_cov, H, Xi, m0, P0, dt, ys)
        r['cd_ekf'] = fs.cd_ekf(drift, dispersion, H, Xi, m0, P0, dt, ys)
        r['sgp_filter'] = fs.sgp_filter(m_and_cov, sg, H, Xi, m0, P0, dt, ys)
        r['cd_sgp_filter'] = fs.cd_sgp_filter(drift, B, sg, H, Xi, m0, P0, dt, ys)
        r['rts'] = fs.rts(F, Sigma, r['kf'][0], r['kf'][1])
        r['eks'] = fs.eks(m_and_cov, r['ekf'][0], r['ekf'][1], dt)
        r['cd_eks'] = fs.cd_eks(drift, dispersion, r['cd_ekf'][0], r['cd_ekf'][1], dt)
        r['sgp_smoother'] = fs.sgp_smoother(m_and_cov, sg, r['sgp_filter'][0], r['sgp_filter'][1], dt)
        r['cd_sgp_smoother'] = fs.cd_sgp_smoother(drift, B, sg, r['cd_sgp_filter'][0], r['cd_sgp_filter'][1], dt)
        for k, v in r.items():
            for j, arr in enumerate(v):
                out['%s_%d' % (k, j)] = npy(arr)
        save('linear_a%d' % idx, **out)


def nonlinear_case(name, builder, params, T, dt, ys, sigmas, with_grad, num_harmonics=1, Xi=0.1):
    """Runs all 8 nonlinear functions of the reference for every sigma-point rule in `sigmas`."""
    params_t = jnp.array(params)
    drift, dispersion, m_and_cov, m0, P0, H = builder(params_t)
    d = m0.shape[0]
    ys_t = jnp.asarray(ys)
    out = dict(params=np.asarray(params), dt=dt, Xi=Xi, ys=np.asarray(ys), m0=npy(m0), P0=npy(P0), H=npy(H),
               num_harmonics=num_harmonics, disp=npy(dispersion(jnp.eye(d))))
    r = {}
    r['ekf'] = fs.ekf(m_and_cov, H, Xi, m0, P0, dt, ys_t)
    r['eks'] = fs.eks(m_and_cov, r['ekf'][0], r['ekf'][1], dt)
    r['cd_ekf'] = fs.cd_ekf(drift, dispersion, H, Xi, m0, P0, dt, ys_t)
    r['cd_eks'] = fs.cd_eks(drift, dispersion, r['cd_ekf'][0], r['cd_ekf'][1], dt)
    for tag, sg in sigmas.items():
        out['sg_w_' + tag] = npy(sg.w)
        out['sg_xi_' + tag] = npy(sg.xi)
        bm = dispersion(jnp.eye(d))
        r['sgp_filter_' + tag] = fs.sgp_filter(m_and_cov, sg, H, Xi, m0, P0, dt, ys_t)
        r['sgp_smoother_' + tag] = fs.sgp_smoother(m_and_cov, sg, r['sgp_filter_' + tag][0], r['sgp_filter_' + tag][1], dt)
        r['cd_sgp_filter_' + tag] = fs.cd_sgp_filter(drift, bm, sg, H, Xi, m0, P0, dt, ys_t)
        r['cd_sgp_smoother_' + tag] = fs.cd_sgp_smoother(drift, bm, sg, r['cd_sgp_filter_' + tag][0],
                                                         r['cd_sgp_filter_' + tag][1], dt)
    for k, v in r.items():
        for j, arr in enumerate(v):
            out['%s_%d' % (k, j)] = npy(arr)
    if with_grad:
        theta = md.g_inv(params_t)
        out['theta'] = npy(theta)

        def obj_ekf(th):
            _, _, mc, m0_, P0_, H_ = builder(md.g(th))
            return fs.ekf(mc, H_, Xi, m0_, P0_, dt, ys_t)[-1][-1]

        def obj_cd_ekf(th):
            dr, di, _, m0_, P0_, H_ = builder(md.g(th))
            return fs.cd_ekf(dr, di, H_, Xi, m0_, P0_, dt, ys_t)[-1][-1]

        out['grad_ekf'] = npy(jax.grad(obj_ekf)(theta))
        out['grad_cd_ekf'] = npy(jax.grad(obj_cd_ekf)(theta))
        for tag, sg in sigmas.items():
            def obj_sgp(th, sg=sg):
                _, _, mc, m0_, P0_, H_ = builder(md.g(th))
                return fs.sgp_filter(mc, sg, H_, Xi, m0_, P0_, dt, ys_t)[-1][-1]

            out['grad_sgp_filter_' + tag] = npy(jax.grad(obj_sgp)(theta))

            def obj_cd_sgp(th, sg=sg):                      # demos/cd_ghfs_mle.py:46-58
                dr, di, _, m0_, P0_, H_ = builder(md.g(th))
                return fs.cd_sgp_filter(dr, di(jnp.eye(d)), sg, H_, Xi, m0_, P0_, dt, ys_t)[-1][-1]

            out['grad_cd_sgp_filter_' + tag] = npy(jax.grad(obj_cd_sgp)(theta))
    save(name, **out)


def kpt_case(name, params, num_harmonics, T, dt, ys, Xi=0.1):
    """ekf_for_kpt + rts on the KPT model (filters_smoothers.py:267-314, models.py:522-580; the pipeline of
    tetralith/jobs/kpt_mle.py:39-62) and jax.grad of its nll w.r.t. theta (:39-42)."""
    fsamp = 1. / dt
    params_t = jnp.array(params)
    F, Sigma, m0, P0, h = md.build_kpt_chirp_model(params_t, fsamp, num_harmonics=num_harmonics)
    ys_t = jnp.asarray(ys)
    out = dict(params=np.asarray(params), fs=fsamp, dt=dt, Xi=Xi, ys=np.asarray(ys), F=npy(jnp.asarray(F)), Sigma=npy(Sigma),
               m0=npy(m0), P0=npy(P0), num_harmonics=num_harmonics)
    f = fs.ekf_for_kpt(jnp.asarray(F), Sigma, h, Xi, m0, P0, dt, ys_t)
    s = fs.rts(jnp.asarray(F), Sigma, f[0], f[1])
    for j in range(3):
        out['ekf_for_kpt_%d' % j] = npy(f[j])
    for j in range(2):
        out['rts_%d' % j] = npy(s[j])
    out['h_at_m0'] = npy(h(m0))
    theta = md.g_inv(params_t)

    def obj(th):
        F_, Sigma_, m0_, P0_, h_ = md.build_kpt_chirp_model(md.g(th), fsamp, num_harmonics=num_harmonics)
        return fs.ekf_for_kpt(jnp.asarray(F_), Sigma_, h_, Xi, m0_, P0_, dt, ys_t)[-1][-1]

    out['theta'] = npy(theta)
    out['grad_ekf_for_kpt'] = npy(jax.grad(obj)(theta))
    save(name, **out)


def kpt_cases():
    dt = 1e-3
    _, ys3, _ = toy.synthetic_batch(3, 3141, dt, Xi=0.1, seed=2)
    kpt_case('kpt', [0.02, 1e-3, 1e-2, 8., 1.], 1, 400, dt, ys3[0, 1200:1600])
    _, ysh, _ = toy.synthetic_batch(2, 3141, dt, Xi=0.1, num_harmonics=3, seed=4)
    kpt_case('kpt_h2', [0.05, 1e-3, 1e-2, 8., 0.7], 2, 250, dt, ysh[1, 1200:1450])


def main():
    if '--only-kpt' in sys.argv:       # added after the other fixtures were committed: leaves them untouched
        kpt_cases()
        return
    # sigma-point tables (bit-exact targets for chirpgp_b200.quadratures)
    tabs = {}
    for d, o in [(4, 3), (3, 4), (1, 5), (1, 10), (2, 3), (8, 2), (4, 5)]:
        s = qd.SigmaPoints.gauss_hermite(d=d, order=o)
        tabs['gh_w_%d_%d' % (d, o)] = npy(s.w)
        tabs['gh_xi_%d_%d' % (d, o)] = npy(s.xi)
    for d in [1, 3, 4, 8, 10, 12]:
        s = qd.SigmaPoints.cubature(d)
        tabs['cub_w_%d' % d] = npy(s.w)
        tabs['cub_xi_%d' % d] = npy(s.xi)
    save('tables', **tabs)

    linear_cases()

    dt = 1e-3
    # chirp: first T samples of the SURVEY 8(d) config-2 synthetic batch, chirps 0..2 (three magnitudes)
    T = 400
    _, ys3, _ = toy.synthetic_batch(3, 3141, dt, Xi=0.1, seed=2)
    # take a stretch in the middle of the record where the frequency actually sweeps
    sl = slice(1200, 1200 + T)
    gh = {'gh3': qd.SigmaPoints.gauss_hermite(d=4, order=3), 'cub': qd.SigmaPoints.cubature(4)}
    nonlinear_case('chirp', md.build_chirp_model, [0.1, 0.1, 0.1, 1., 1., 7.], T, dt, ys3[2, sl], gh, True)
    nonlinear_case('chirp_lam0', md.build_chirp_model, [0., 0.3, 0.2, 0.8, 1.5, 3.], 200, dt, ys3[0, sl][:200],
                   {'gh3': gh['gh3']}, False)
    nonlinear_case('lascala', md.build_lascala_model, [0.1, 1., 1., 7.], 200, dt, ys3[1, sl][:200],
                   {'gh3': gh['gh3']}, False)
    nonlinear_case('short', md.build_chirp_model, [0.1, 0.1, 0.1, 1., 1., 7.], 2, dt, np.ones(2),
                   {'gh3': gh['gh3']}, False)

    _, ysh, _ = toy.synthetic_batch(2, 3141, dt, Xi=0.1, num_harmonics=3, seed=4)
    hb = lambda p: md.build_harmonic_chirp_model(p, num_harmonics=3)
    nonlinear_case('harmonic', hb, [0.1, 0.1, 0.1, 1., 1., 7.], 250, dt, ysh[1, 1200:1450],
                   {'cub': qd.SigmaPoints.cubature(8)}, True, num_harmonics=3)
    hb2 = lambda p: md.build_harmonic_chirp_model(p, num_harmonics=2, freq_scale=1.7)
    nonlinear_case('harmonic2', hb2, [0.2, 0.15, 0.1, 0.7, 1.2, 5.], 120, dt, ysh[0, 1200:1320],
                   {'cub': qd.SigmaPoints.cubature(6)}, False, num_harmonics=2)
    kpt_cases()


if __name__ == '__main__':
    main()
