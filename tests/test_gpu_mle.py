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

    c2 = consts.detach().clone().requires_grad_(True)
    m2, P2 = m0.clone().requires_grad_(True), P0.clone().requires_grad_(True)
    X2 = torch.tensor(0.1, dtype=torch.float64, requires_grad=True)
    got = mle.ekf_nll(_M(c2), H, X2, m2, P2, dt, ys, ckpt_every=7)
    gg = torch.autograd.grad(got, [c2, m2, P2, X2])
    npt.assert_allclose(got.item(), want.item(), rtol=1e-11)
    for a, b, nm in zip(gg, gw, ['consts', 'm0', 'P0', 'Xi']):
        npt.assert_allclose(a.cpu().numpy()[..., :9] if nm == 'consts' else a.cpu().numpy(),
                            b.numpy()[..., :9] if nm == 'consts' else b.numpy(), rtol=1e-7, atol=1e-9, err_msg=nm)


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


def test_fd_gradient_mle_for_sigma_point_and_cd_filters(golden):
    """The MLE demos of the sigma-point / continuous-discrete filters (demos/ghfs_mle.py:54-61, cd_ekfs_mle.py,
    cd_ghfs_mle.py) take jax.grad of their nll; here those objectives are differentiated by five-point central differences
    over a candidate batch.  Check objective and gradient against the reference's jax.grad fixtures."""
    z = golden('chirp')
    H, Xi, dt, ys = z['H'], float(z['Xi']), float(z['dt']), z['ys']
    sg = cg.SigmaPoints.gauss_hermite(4, 3)
    cases = [('sgp_filter', sg, z['grad_sgp_filter_gh3'], z['sgp_filter_gh3_2'][-1]),
             ('cd_ekf', None, z['grad_cd_ekf'], z['cd_ekf_2'][-1])]
    for method, sgps, want_grad, want_val in cases:
        calls = []

        def spy(fun, x0, jac, method, options):
            v, gr = fun(np.asarray(x0))
            calls.append((v, gr))

            class R:
                x, success = np.asarray(x0), True
                fun_ = v
            return R()

        import scipy.optimize
        orig = scipy.optimize.minimize
        scipy.optimize.minimize = spy
        try:
            mle.fit_mle(cg.build_chirp_model, z['theta'], H, Xi, dt, ys, method=method, sgps=sgps)
        finally:
            scipy.optimize.minimize = orig
        v, gr = calls[0]
        npt.assert_allclose(v, want_val, rtol=1e-10)
        # measured: 2.8e-7 (sgp_filter) -- the floor is the nll's own rounding noise (~1e-12 relative) divided by h
        npt.assert_allclose(gr, want_grad, rtol=2e-6, atol=1e-7)


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
