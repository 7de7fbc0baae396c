"""MLE path on the GPU: nll kernel vs the filter kernel, adjoint kernel vs (a) the gradients that the reference's own
sources produce under jax.grad (golden fixtures), (b) a torch autodiff twin of the EKF (cotangents of every kernel
input, including the unsymmetrised P0 convention), (c) central finite differences at full length.

Tolerances: nll rtol 1e-11 against the filter kernel; gradients rtol 1e-7 against autodiff, 1e-5 against
finite differences."""
import numpy as np
import numpy.testing as npt
import pytest
import torch

import chirpgp_b200 as cg
from chirpgp_b200 import mle, toymodels
from chirpgp_b200.models import g as gfun

pytestmark = pytest.mark.gpu


def _theta_grad(builder, theta_np, H, Xi, dt, ys, **kw):
    theta = torch.tensor(theta_np, dtype=torch.float64, requires_grad=True)
    _, _, mc, m0, P0, _ = builder(gfun(theta))
    val = mle.ekf_nll(mc, H, Xi, m0, P0, dt, ys, **kw)
    grad, = torch.autograd.grad(val.sum(), theta)
    return val.detach().cpu().numpy(), grad.numpy()


@pytest.mark.parametrize('name,builder', [
    ('chirp', cg.build_chirp_model),
    ('harmonic', lambda p: cg.build_harmonic_chirp_model(p, num_harmonics=3)),
])
def test_gradient_matches_reference_jax_grad(golden, name, builder):
    z = golden(name)
    val, grad = _theta_grad(builder, z['theta'], z['H'], float(z['Xi']), float(z['dt']), z['ys'], ckpt_every=37)
    npt.assert_allclose(val, z['ekf_2'][-1], rtol=1e-11)
    npt.assert_allclose(grad, z['grad_ekf'], rtol=1e-7, atol=1e-9)


def test_nll_equals_filter_and_fd_at_full_length():
    B, T, dt = 6, 3141, 1e-3
    _, ys, _ = toymodels.synthetic_batch(B, T, dt, Xi=0.1, seed=2)
    theta0 = np.log(np.exp(np.array([0.1, 0.1, 0.1, 1., 1., 7.])) - 1.)
    val, grad = _theta_grad(cg.build_chirp_model, theta0, np.array([0., 1., 0., 0.]), 0.1, dt, ys)
    _, _, mc, m0, P0, H = cg.build_chirp_model(gfun(torch.tensor(theta0)))
    f = cg.ekf(mc, H, 0.1, m0, P0, dt, ys)
    npt.assert_allclose(val, f[2][:, -1], rtol=1e-11)
    fd = np.zeros(6)
    for i in range(6):
        h = 1e-5 * max(1., abs(theta0[i]))
        tp, tm = theta0.copy(), theta0.copy()
        tp[i] += h; tm[i] -= h
        vp = mle.ekf_nll(cg.build_chirp_model(gfun(torch.tensor(tp)))[2], H, 0.1, *cg.build_chirp_model(gfun(torch.tensor(tp)))[3:5], dt, ys).sum()
        vm = mle.ekf_nll(cg.build_chirp_model(gfun(torch.tensor(tm)))[2], H, 0.1, *cg.build_chirp_model(gfun(torch.tensor(tm)))[3:5], dt, ys).sum()
        fd[i] = float(vp - vm) / (2 * h)
    npt.assert_allclose(grad, fd, rtol=1e-5, atol=1e-6)


def test_cotangents_of_every_kernel_input_vs_autodiff_twin():
    from oracle import ekf_torch
    T, dt, nh = 60, 1e-3, 1
    rng = np.random.default_rng(3)
    _, ys, _ = toymodels.synthetic_batch(1, 3141, dt, Xi=0.1, seed=9)
    ys = ys[0, 1500:1500 + T]
    _, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.3, 0.2, 0.15, 0.9, 1.2, 6.]))
    A = rng.standard_normal((4, 4)) * 0.05
    P0 = P0 + torch.as_tensor(A @ A.T)                      # dense SPD so every entry of P0_bar is exercised
    consts = mc.consts(dt).clone().requires_grad_(True)
    m0c = m0.clone().requires_grad_(True)
    P0c = P0.clone().requires_grad_(True)
    Xi = torch.tensor(0.1, dtype=torch.float64, requires_grad=True)
    want = ekf_torch.ekf_nll(consts, H, Xi, m0c, P0c, dt, torch.as_tensor(ys), nh)
    gw = torch.autograd.grad(want, [consts, m0c, P0c, Xi])

    class _M(cg.models.LCDModel):                           # feed the same constants tensor through the kernel path
        def __init__(self, c):
            self._c, self.num_harmonics, self.d, self.freq_scale = c, 1, 4, 1.

        def consts(self, dt):
            return self._c

    for raw in (True, False):
        c2 = consts.detach().clone().requires_grad_(True)
        m2, P2 = m0.clone().requires_grad_(True), P0.clone().requires_grad_(True)
        X2 = torch.tensor(0.1, dtype=torch.float64, requires_grad=True)
        got = mle.ekf_nll(_M(c2), H, X2, m2, P2, dt, ys, ckpt_every=7, raw_p0_cotangent=raw)
        gg = torch.autograd.grad(got, [c2, m2, P2, X2])
        npt.assert_allclose(got.item(), want.item(), rtol=1e-11)
        for a, b, nm in zip(gg, gw, ['consts', 'm0', 'P0', 'Xi']):
            a, b = a.cpu().numpy(), b.numpy()
            if nm == 'consts':
                a, b = a[..., :9], b[..., :9]
            if nm == 'P0' and not raw:                     # default: the symmetrised covariance cotangent
                b = 0.5 * (b + b.T)
            npt.assert_allclose(a, b, rtol=1e-7, atol=1e-9, err_msg='%s raw=%s' % (nm, raw))
    assert np.abs(gw[2].numpy() - gw[2].numpy().T).max() > 1e-3    # the case does exercise the antisymmetric part


def test_general_measurement_row_and_ragged_chains():
    """H that is not a unit vector (generic kernel instance), a number of problems that is not a multiple of the 32-problem
    chain, several segments per chain, against the filter kernel and central differences."""
    B, T, dt = 37, 300, 1e-3
    _, ys, _ = toymodels.synthetic_batch(B, 3141, dt, Xi=0.1, seed=11)
    ys = ys[:, 700:700 + T]
    Hrow = np.array([0.3, 1., 0., 0.2])
    theta0 = np.log(np.exp(np.array([0.2, 0.15, 0.1, 1.1, 0.9, 6.5])) - 1.)
    val, grad = _theta_grad(cg.build_chirp_model, theta0, Hrow, 0.1, dt, ys, ckpt_every=16)
    _, _, mc, m0, P0, _ = cg.build_chirp_model(gfun(torch.tensor(theta0)))
    f = cg.ekf(mc, Hrow, 0.1, m0, P0, dt, ys)
    npt.assert_allclose(val, f[2][:, -1], rtol=1e-11)
    fd = np.zeros(6)
    for i in range(6):
        h = 1e-5 * max(1., abs(theta0[i]))
        v = []
        for sgn in (1., -1.):
            th = theta0.copy(); th[i] += sgn * h
            _, _, mc1, m01, P01, _ = cg.build_chirp_model(gfun(torch.tensor(th)))
            v.append(float(mle.ekf_nll(mc1, Hrow, 0.1, m01, P01, dt, ys).sum()))
        fd[i] = (v[0] - v[1]) / (2 * h)
    npt.assert_allclose(grad, fd, rtol=1e-5, atol=1e-6)


def test_config5_shape_nll_and_adjoint():
    """BASELINE configs[4] shape: T = 1e5 steps (dt = 3.141e-5), chirps x hyper-parameter candidates, DEFAULT checkpoint stride
    (6250 segments of 16 steps, units of several segments, chains migrating between warps).  nll against the filter kernel's
    n_ell[-1] at rtol 1e-11; the adjoint gradient against (a) the forward-mode tangent kernel at full length -- an independent
    derivative code path -- at rtol 1e-9, (b) central finite differences at full length, (c) the torch autodiff twin (jacfwd
    of the model mean, reverse mode through the Python loop; ~15 ms per step on the host) on a T = 2000 prefix at rtol 1e-7."""
    from oracle import ekf_torch
    T, G, Xi = 100000, 4, 0.1
    dt = 3.141 / T
    rng = np.random.default_rng(5)
    ts = np.linspace(dt, dt * T, T)
    ys = np.sin(2 * np.pi * (500 * np.exp(-5 / np.sin(ts)) + 8 * ts))[None] + np.sqrt(Xi) * rng.standard_normal((2, T))
    grid = np.array([[0.1, 0.05, 0.1, 1., 1., 7.], [0.4, 0.1, 0.1, 1., 1., 7.], [0.7, 0.2, 0.1, 1., 1., 7.], [1.0, 0.4, 0.1, 1., 1., 7.]])
    theta_np = np.log(np.exp(grid) - 1.)
    H = np.array([0., 1., 0., 0.])
    theta = torch.tensor(theta_np, dtype=torch.float64, device='cuda', requires_grad=True)
    _, _, mc, m0, P0, _ = cg.build_chirp_model(gfun(theta))
    nll = mle.ekf_nll(mc, H, Xi, m0, P0, dt, ys, candidates=True)                       # (2, G), default ckpt_every
    assert mle._default_ckpt.last['ckpt_every'] == 16
    grad, = torch.autograd.grad(nll.sum(), theta)                                       # (G, 6): summed over the 2 chirps
    nll, grad = nll.detach().cpu().numpy(), grad.cpu().numpy()
    for gi in range(G):
        _, _, mc1, m01, P01, _ = cg.build_chirp_model(grid[gi])
        f = cg.ekf(mc1, H, Xi, m01, P01, dt, ys)
        npt.assert_allclose(nll[:, gi], f[2][:, -1], rtol=1e-11)
    # (a) tangent kernel, full length
    for gi in (0, 3):
        v, gt = mle.filter_nll_grad('ekf', cg.build_chirp_model, theta_np[gi], H, Xi, dt, ys)
        npt.assert_allclose(v.cpu().numpy(), nll[:, gi], rtol=1e-11)
        npt.assert_allclose(gt.sum(0).cpu().numpy(), grad[gi], rtol=1e-9, atol=1e-9)
    # (b) central differences, full length, candidate 1
    fd = np.zeros(6)
    for i in range(6):
        h = 1e-5 * max(1., abs(theta_np[1, i]))
        v = []
        for sgn in (1., -1.):
            th = theta_np[1].copy(); th[i] += sgn * h
            _, _, mcp, m0p, P0p, _ = cg.build_chirp_model(gfun(torch.tensor(th)))
            v.append(float(mle.ekf_nll(mcp, H, Xi, m0p, P0p, dt, ys).sum()))
        fd[i] = (v[0] - v[1]) / (2 * h)
    npt.assert_allclose(grad[1], fd, rtol=1e-5, atol=1e-4)
    # (c) autodiff twin on a prefix (one chirp, candidate 2)
    Tp = 2000
    th = torch.tensor(theta_np[2], dtype=torch.float64, requires_grad=True)
    _, _, mct, m0t, P0t, Ht = cg.build_chirp_model(gfun(th))
    want = ekf_torch.ekf_nll(mct.consts(dt), Ht, torch.tensor(Xi, dtype=torch.float64), m0t, P0t, dt, torch.as_tensor(ys[0, :Tp]), 1)
    gw, = torch.autograd.grad(want, th)
    th2 = torch.tensor(theta_np[2], dtype=torch.float64, device='cuda', requires_grad=True)
    _, _, mc2, m02, P02, _ = cg.build_chirp_model(gfun(th2))
    got = mle.ekf_nll(mc2, H, Xi, m02, P02, dt, ys[0, :Tp])
    gg, = torch.autograd.grad(got, th2)
    npt.assert_allclose(got.item(), want.item(), rtol=1e-11)
    npt.assert_allclose(gg.cpu().numpy(), gw.numpy(), rtol=1e-7, atol=1e-9)


def test_candidate_grid_and_fit():
    B, T, dt = 4, 400, 1e-3
    _, ys, _ = toymodels.synthetic_batch(B, 3141, dt, Xi=0.1, seed=5)
    ys = ys[:, 1000:1000 + T]
    lam = np.array([0.1, 0.4, 0.7, 1.0])
    bb = np.array([0.05, 0.1, 0.2, 0.4])
    grid = np.array([[l, b_, 0.1, 1., 1., 7.] for l in lam for b_ in bb])
    _, _, mc, m0, P0, H = cg.build_chirp_model(grid)
    nll = mle.ekf_nll(mc, H, 0.1, m0, P0, dt, ys, candidates=True)
    assert nll.shape == (B, 16)
    for gi in (0, 7, 15):
        _, _, mc1, m01, P01, _ = cg.build_chirp_model(grid[gi])
        f = cg.ekf(mc1, H, 0.1, m01, P01, dt, ys)
        npt.assert_allclose(nll[:, gi].cpu().numpy(), f[2][:, -1], rtol=1e-11)
    theta0 = np.log(np.exp(np.array([0.1, 0.1, 0.1, 1., 1., 7.])) - 1.)
    theta, res = mle.fit_mle(cg.build_chirp_model, theta0, H, 0.1, dt, ys[:1], maxiter=15)
    assert res.fun < float(mle.ekf_nll(*[cg.build_chirp_model(gfun(torch.tensor(theta0)))[k] for k in (2,)], H, 0.1,
                                       *cg.build_chirp_model(gfun(torch.tensor(theta0)))[3:5], dt, ys[:1]).sum())


def _builder(name):
    return cg.build_chirp_model if name == 'chirp' else (lambda p: cg.build_harmonic_chirp_model(p, num_harmonics=3))


@pytest.mark.parametrize('name,method,tag', [
    ('chirp', 'ekf', None), ('chirp', 'cd_ekf', None), ('chirp', 'sgp_filter', 'gh3'), ('chirp', 'sgp_filter', 'cub'),
    ('chirp', 'cd_sgp_filter', 'gh3'), ('chirp', 'cd_sgp_filter', 'cub'),
    ('harmonic', 'ekf', None), ('harmonic', 'cd_ekf', None), ('harmonic', 'sgp_filter', 'cub'), ('harmonic', 'cd_sgp_filter', 'cub'),
])
def test_tangent_kernels_match_reference_jax_grad(golden, name, method, tag):
    """The MLE demos of the sigma-point / continuous-discrete filters (demos/ghfs_mle.py:54-61, cd_ekfs_mle.py,
    cd_ghfs_mle.py) take jax.grad of their nll.  Forward-mode tangent kernels vs the gradients the reference's own sources
    produce (fixtures): value rtol 1e-10, gradient rtol 1e-7."""
    z = golden(name)
    key = method if tag is None else '%s_%s' % (method, tag)
    if 'grad_' + key not in z:
        pytest.skip('fixture has no grad_%s' % key)
    d = int(z['m0'].shape[0])
    sgps = None
    if tag is not None:
        sgps = cg.SigmaPoints.gauss_hermite(d, 3) if tag == 'gh3' else cg.SigmaPoints.cubature(d)
    val, grad = mle.filter_nll_grad(method, _builder(name), z['theta'], z['H'], float(z['Xi']), float(z['dt']), z['ys'], sgps=sgps)
    npt.assert_allclose(val.cpu().numpy()[0], z[key + '_2'][-1], rtol=1e-10)
    npt.assert_allclose(grad.cpu().numpy()[0], z['grad_' + key], rtol=1e-7, atol=1e-9)


def test_tangent_kernel_per_chirp_parameters_and_fit_mle():
    """theta (B, P): one parameter vector per chirp in one launch == B separate launches; fit_mle on the exact gradients."""
    B, T, dt = 5, 200, 1e-3
    _, ys, _ = toymodels.synthetic_batch(B, 3141, dt, Xi=0.1, seed=21)
    ys = ys[:, 900:900 + T]
    rng = np.random.default_rng(5)
    theta0 = np.log(np.exp(np.array([0.1, 0.1, 0.1, 1., 1., 7.])) - 1.)
    thetas = theta0 + 0.05 * rng.standard_normal((B, 6))
    H = np.array([0., 1., 0., 0.])
    sg = cg.SigmaPoints.gauss_hermite(4, 3)
    for method, sgps in (('sgp_filter', sg), ('cd_ekf', None)):
        v, gr = mle.filter_nll_grad(method, cg.build_chirp_model, thetas, H, 0.1, dt, ys, sgps=sgps)
        for i in (0, B - 1):
            v1, g1 = mle.filter_nll_grad(method, cg.build_chirp_model, thetas[i], H, 0.1, dt, ys[i], sgps=sgps)
            npt.assert_allclose(v[i].item(), v1[0].item(), rtol=1e-14)
            npt.assert_allclose(gr[i].cpu().numpy(), g1[0].cpu().numpy(), rtol=1e-12, atol=1e-12)
    theta, res = mle.fit_mle(cg.build_chirp_model, theta0, H, 0.1, dt, ys[:2], method='sgp_filter', sgps=sg, maxiter=8)
    v0, _ = mle.filter_nll_grad('sgp_filter', cg.build_chirp_model, theta0, H, 0.1, dt, ys[:2], sgps=sg)
    assert res.fun < float(v0.sum())


def test_fit_mle_batched_follows_the_sequential_fits():
    """tetralith/jobs/ekfs_mle.py:26-86 runs one L-BFGS-B fit per Monte-Carlo chirp; fit_mle_batched advances all of them in
    lock-step with one batched nll / gradient launch per iteration.  Each chirp must reach the optimum its own sequential
    fit_mle reaches (same SciPy optimiser, same objective values: the iterates coincide)."""
    B, T, dt = 24, 300, 1e-3
    _, ys, _ = toymodels.synthetic_batch(B, 3141, dt, Xi=0.1, seed=31)
    ys = ys[:, 1200:1200 + T]
    theta0 = np.log(np.exp(np.array([0.1, 0.1, 0.1, 1., 1., 7.])) - 1.)
    H = np.array([0., 1., 0., 0.])
    thetas, results = mle.fit_mle_batched(cg.build_chirp_model, theta0, H, 0.1, dt, ys, maxiter=12, nan_on_failure=False)
    assert thetas.shape == (B, 6)
    launches = mle.fit_mle_batched.last_launches
    assert launches <= max(r.nfev for r in results)          # lock-step: one launch serves every running fit
    for i in (0, 7, 23):
        th_i, res_i = mle.fit_mle(cg.build_chirp_model, theta0, H, 0.1, dt, ys[i:i + 1], maxiter=12)
        npt.assert_allclose(thetas[i], th_i, rtol=1e-9, atol=1e-12)
        npt.assert_allclose(results[i].fun, res_i.fun, rtol=1e-12)
        assert results[i].nit == res_i.nit
    # NaN convention of the reference for failed fits
    th_nan, res_nan = mle.fit_mle_batched(cg.build_chirp_model, theta0, H, 0.1, dt, ys[:3], maxiter=1)
    for row, r in zip(th_nan, res_nan):
        assert np.all(np.isnan(row)) == (not r.success)


def test_kpt_mle_objective_and_gradient(golden):
    """tetralith/jobs/kpt_mle.py:39-42: obj_func(theta) = ekf_for_kpt(...)[-1][-1] and its jax.grad, here by five-point
    differences over a candidate batch (fit_mle(method='ekf_for_kpt'))."""
    import scipy.optimize
    for name in ('kpt', 'kpt_h2'):
        z = golden(name)
        nh, fsamp = int(z['num_harmonics']), float(z['fs'])
        calls = []

        def spy(fun, x0, jac, method, options):
            calls.append(fun(np.asarray(x0)))

            class R:
                x, success = np.asarray(x0), True
            return R()

        orig = scipy.optimize.minimize
        scipy.optimize.minimize = spy
        try:
            mle.fit_mle(lambda p: cg.build_kpt_chirp_model(p, fsamp, nh), z['theta'], None, float(z['Xi']), float(z['dt']),
                        z['ys'], method='ekf_for_kpt')
        finally:
            scipy.optimize.minimize = orig
        v, gr = calls[0]
        npt.assert_allclose(v, z['ekf_for_kpt_2'][-1], rtol=1e-10)
        npt.assert_allclose(gr, z['grad_ekf_for_kpt'], rtol=2e-6, atol=1e-6)


def test_nll_path_and_arbitrary_cotangent():
    """ekf(...)[-1] as a differentiable (T,) output: values == the filter kernel's n_ell; the gradient of a
    randomly weighted sum of ALL steps matches the torch autodiff twin, and the all-weight-on-the-last-step case reproduces
    ekf_nll's gradient."""
    from oracle import ekf_torch
    from chirpgp_b200.models import g_inv
    rng = np.random.default_rng(12)
    T, dt, Xi = 40, 1e-3, 0.1
    _, ys, _ = toymodels.synthetic_batch(2, 3141, dt, Xi=Xi, seed=2)
    ys = ys[:, 1200:1200 + T]
    theta0 = g_inv(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
    ct = rng.standard_normal((2, T))

    def run(use_path, cot):
        theta = torch.tensor(theta0, dtype=torch.float64, requires_grad=True)
        _, _, mc, m0, P0, H = cg.build_chirp_model(gfun(theta))
        if use_path:
            nell = mle.ekf_nll_path(mc, H, Xi, m0, P0, dt, ys)
            assert tuple(nell.shape) == (2, T)
            loss = (nell * torch.as_tensor(cot, device=nell.device)).sum()
        else:
            nell = None
            loss = mle.ekf_nll(mc, H, Xi, m0, P0, dt, ys).sum()
        grad, = torch.autograd.grad(loss, theta)
        return (None if nell is None else nell.detach().cpu().numpy()), grad.cpu().numpy()

    nell, grad = run(True, ct)
    _, _, mc, m0, P0, H = cg.build_chirp_model(gfun(torch.tensor(theta0)))
    f = cg.ekf(mc, H, Xi, m0, P0, dt, ys)
    npt.assert_allclose(nell, f[2], rtol=1e-12)
    # autodiff twin: n_ell_k is the final nll of the first k + 1 samples
    theta = torch.tensor(theta0, dtype=torch.float64, requires_grad=True)
    _, _, mc_t, m0_t, P0_t, H_t = cg.build_chirp_model(gfun(theta))
    consts = mc_t.consts(dt)
    total = torch.zeros((), dtype=torch.float64)
    for b in range(2):
        for k in range(T):
            total = total + float(ct[b, k]) * ekf_torch.ekf_nll(consts, H_t, Xi, m0_t, P0_t, dt, torch.as_tensor(ys[b, :k + 1]), 1)
    want, = torch.autograd.grad(total, theta)
    npt.assert_allclose(grad, want.numpy(), rtol=1e-7, atol=1e-8)
    # weight on the last step only == gradient of the final nll
    last = np.zeros((2, T)); last[:, -1] = 1.
    _, g_last = run(True, last)
    _, g_final = run(False, None)
    npt.assert_allclose(g_last, g_final, rtol=1e-12, atol=1e-13)
