"""CPU tests of the host-side logic and of the C-ABI surface (no GPU, no compute calls):
 * the shared library loads and exports every symbol include/chirpgp_b200.h declares;
 * model constants / host evaluation against the oracle and against SciPy expm (the reference's test/test_models.py:26-78
   and test/test_m32.py:14-30, restated without JAX);
 * quadrature invariants (test/test_quadratures.py:19-59) and gaussian_expectation (test/test_utils.py:84-95);
 * product code never imports the oracle; missing CUDA raises instead of falling back."""
import ctypes
import math
import os
import re

import numpy as np
import numpy.testing as npt
import pytest
import scipy.linalg
import torch

import chirpgp_b200 as cg
from chirpgp_b200 import _native, models
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'chirpgp_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(cgp_[a-z0-9_]+)\s*\(', hdr)))
    assert len(declared) >= 18
    assert os.path.exists(_native.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_native.EXPORTED) == declared
    lib.cgp_abi_version.restype = ctypes.c_int
    assert lib.cgp_abi_version() == _native.ABI_VERSION
    # argument errors are reported, not thrown (no GPU work involved)
    lib.cgp_ekf_f64.restype = ctypes.c_int
    assert lib.cgp_ekf_f64(None, None, None, None, None, 0, None) == -1
    p = _native.CgpProblem()
    lib.cgp_workspace_bytes.restype = ctypes.c_size_t
    p.B, p.T, p.d = 3, 5, 4
    assert lib.cgp_workspace_bytes(b'eks', ctypes.byref(p)) == 3 * 5 * 30 * 8          # [G 16 | c 4 | C 10] per (chirp, step)


def test_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, 'include', 'chirpgp_b200.h')).read()
    body = hdr[hdr.index('typedef struct CgpProblem {'):hdr.index('} CgpProblem;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'(?:int64_t|int32_t|double|const double \*)\s*\**(\w+)\s*;', body)
    assert names == [f[0] for f in _native.CgpProblem._fields_]


def test_no_cpu_fallback_and_no_oracle_import():
    pkg = os.path.join(ROOT, 'chirpgp_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src.replace('# oracle', ''), fn
    if not torch.cuda.is_available():
        _, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
        with pytest.raises(RuntimeError):
            cg.ekf(mc, H, 0.1, m0, P0, 1e-3, np.ones(4))
        with pytest.raises(RuntimeError):         # the batch-sequence call has no CPU path either
            next(cg.filter_smoother_batches(cg.ekf_smoother, mc, H, 0.1, m0, P0, 1e-3, batches=[np.ones(4)]))


def test_filter_smoother_batches_argument_checks():
    _, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
    with pytest.raises(TypeError):                # only the four pair functions can be sequenced
        next(cg.filter_smoother_batches(cg.ekf, mc, H, 0.1, m0, P0, 1e-3, batches=[np.ones(4)]))
    with pytest.raises(ValueError):
        next(cg.filter_smoother_batches(cg.ekf_smoother, mc, H, 0.1, m0, P0, 1e-3, batches=[np.ones(4)], depth=0))


@pytest.mark.parametrize('ell,sigma,dt', [(0.1, 0.1, 0.1), (1., 1., 0.1), (0.5, 2., 0.01), (2.2, 0.3, 1.)])
def test_m32_solution_vs_lti_discretisation(ell, sigma, dt):
    """test/test_m32.py:14-30: closed form == matrix-fraction discretisation of the Matern-3/2 SDE."""
    gam = math.sqrt(3) / ell
    A = np.array([[0., 1.], [-gam ** 2, -2 * gam]])
    Bm = np.array([[0.], [2 * sigma * gam ** 1.5]])
    F = scipy.linalg.expm(A * dt)
    n = 2
    phi = scipy.linalg.expm(np.block([[A, Bm @ Bm.T], [np.zeros((n, n)), -A.T]]) * dt)
    Sig = phi[:n, n:] @ F.T
    f, s_ = models._m32_solution(torch.tensor(ell, dtype=torch.float64), torch.tensor(sigma, dtype=torch.float64), dt)
    host = (np.array([float(v) for v in f]).reshape(2, 2), np.array([float(s_[i]) for i in (0, 1, 1, 2)]).reshape(2, 2))
    for Ft, St in (orc.m32_solution(ell, sigma, dt), host):
        npt.assert_allclose(Ft, F, atol=1e-12)
        npt.assert_allclose(St, Sig, atol=1e-12)


@pytest.mark.parametrize('lam,b,ell', [(0.1, 0.1, 0.1), (1., 1., 1.), (0.1, 1., 0.1)])
@pytest.mark.parametrize('h', [1, 2, 3])
def test_lcd_mean_is_expm_of_frozen_drift(lam, b, ell, h):
    """test/test_models.py:26-78: with the frequency state frozen the LCD mean matrix is expm(A dt) and the LCD
    covariance is the exact LTI discretisation."""
    sigma, dt = 1.3, 0.1
    d = 2 * h + 2
    rng = np.random.default_rng(0)
    u = rng.standard_normal(d)
    mc = cg.disc_harmonic_chirp_lcd(lam, b, ell, sigma, num_harmonics=h)
    drift = models.SDEDrift(lam, ell, h)
    disp = models.Dispersion(b, ell, sigma, h).matrix().numpy()
    w = 2 * math.pi * models.g(u[d - 2])
    gam = math.sqrt(3) / ell
    A = np.zeros((d, d))
    for k in range(1, h + 1):
        A[2 * k - 2:2 * k, 2 * k - 2:2 * k] = [[-lam, -w * k], [w * k, -lam]]
    A[d - 2:, d - 2:] = [[0., 1.], [-gam ** 2, -2 * gam]]
    npt.assert_allclose(drift(u).numpy(), A @ u, rtol=1e-13)
    mean, cov = mc(u, dt)
    npt.assert_allclose(mean.numpy(), scipy.linalg.expm(A * dt) @ u, rtol=1e-10, atol=1e-12)
    phi = scipy.linalg.expm(np.block([[A, disp @ disp.T], [np.zeros((d, d)), -A.T]]) * dt)
    npt.assert_allclose(cov.numpy(), phi[:d, d:] @ scipy.linalg.expm(A * dt).T, rtol=1e-10, atol=1e-10)
    # host model == oracle model (independent restatements of models.py:295-309 / :369-384)
    om, oJ, oS = orc.disc_mean_cov(orc.ChirpSpec(lam, b, ell, sigma, num_harmonics=h), dt, u)
    npt.assert_allclose(mean.numpy(), om, rtol=1e-14, atol=1e-15)
    npt.assert_allclose(cov.numpy(), oS, rtol=1e-14, atol=1e-18)
    # closed-form Jacobians of the oracle against finite differences
    for i in range(d):
        e = np.zeros(d); e[i] = 1e-6
        fd = (orc.disc_mean_cov(orc.ChirpSpec(lam, b, ell, sigma, num_harmonics=h), dt, u + e)[0]
              - orc.disc_mean_cov(orc.ChirpSpec(lam, b, ell, sigma, num_harmonics=h), dt, u - e)[0]) / 2e-6
        npt.assert_allclose(oJ[:, i], fd, rtol=1e-6, atol=1e-8)


def test_lam0_branch_and_lascala():
    mc0 = cg.disc_chirp_lcd(0., 0.3, 1., 1.)
    c = mc0.consts(0.01)
    assert float(c[0]) == 1. and float(c[5]) == 0.3 ** 2 * 0.01            # models.py:302-303
    las = cg.disc_model_lascala_lcd(1., 1.)
    cl = las.consts(0.01)
    assert float(cl[0]) == 1. and float(cl[5]) == 0.
    npt.assert_array_equal(cl[1:5].numpy(), c[1:5].numpy())                # test/test_models.py:127
    x = np.array([0.1, 2., 3., 50.])
    npt.assert_allclose(models.g_inv(models.g(x)), x, rtol=1e-14)          # test/test_models.py:24


def test_build_models_shapes_and_gradients():
    theta = torch.tensor(np.log(np.exp(np.array([0.1, 0.1, 0.1, 1., 1., 7.])) - 1.), requires_grad=True)
    drift, disp, mc, m0, P0, H = cg.build_chirp_model(models.g(theta))
    assert m0.shape == (4,) and P0.shape == (4, 4) and H.tolist() == [0., 1., 0., 0.]
    (mc.consts(1e-3).sum() + P0.sum() + m0.sum()).backward()
    assert torch.isfinite(theta.grad).all()
    grid = np.tile(np.array([0.1, 0.1, 0.1, 1., 1., 7.]), (5, 1))
    _, _, mcg, m0g, P0g, _ = cg.build_harmonic_chirp_model(grid, num_harmonics=3)
    assert mcg.consts(1e-3).shape == (5, models.NC_LCD) and m0g.shape == (5, 8) and P0g.shape == (5, 8, 8)
    npt.assert_array_equal(m0g[0].numpy(), [0., 1., 0., 1., 0., 1., 7., 0.])   # models.py:489


def test_quadrature_invariants():
    """test/test_quadratures.py:19-59."""
    for s in (cg.SigmaPoints.cubature(4), cg.SigmaPoints.gauss_hermite(1, 5), cg.SigmaPoints.gauss_hermite(4, 3)):
        npt.assert_allclose(s.w.sum(), 1., rtol=1e-14)
    with pytest.raises(NotImplementedError):
        cg.SigmaPoints.unscented(4, 1., 2., 0.)
    rng = np.random.default_rng(1)
    d = 3
    m = rng.standard_normal(d)
    A = rng.standard_normal((d, d)); P = A @ A.T + np.eye(d)
    Q = rng.standard_normal((d, d))
    for s in (cg.SigmaPoints.cubature(d), cg.SigmaPoints.gauss_hermite(d, 3)):
        chi = s.gen_sigma_points(m, np.linalg.cholesky(P))
        val = s.expectation(np.einsum('ni,ij,nj->n', chi, Q, chi))
        npt.assert_allclose(val, np.trace(Q @ P) + m @ Q @ m, rtol=1e-12)     # exact on quadratics
    gh = cg.SigmaPoints.gauss_hermite(1, 10)
    chi = gh.gen_sigma_points(np.array([0.3]), np.array([[0.7]]))
    npt.assert_allclose(gh.expectation(np.sin(chi[:, 0])), math.sin(0.3) * math.exp(-0.7 ** 2 / 2), rtol=1e-9)


def test_gaussian_expectation_closed_form():
    """test/test_utils.py:84-95: E[exp(V)] = exp(m + P/2)."""
    ms = np.array([0.1, -0.4, 1.2]); Ps = np.array([0.2, 0.5, 0.05])
    got = cg.gaussian_expectation(ms, np.sqrt(Ps), func=np.exp, force_shape=True)[:, 0]
    npt.assert_allclose(got, np.exp(ms + Ps / 2), rtol=1e-8)


def test_toymodels_phase_derivative_is_frequency():
    """test/test_toymodels.py:42-55."""
    f, ph = cg.meow_freq(offset=8.)
    ts = np.linspace(0.5, 2.5, 2001)
    npt.assert_allclose(np.gradient(ph(ts), ts)[5:-5], f(ts)[5:-5], rtol=1e-4)
    _, ys, _ = cg.toymodels.synthetic_batch(3, 64, 1e-3, seed=2)
    _, ys2, _ = cg.toymodels.synthetic_batch(3, 64, 1e-3, seed=2)
    npt.assert_array_equal(ys, ys2)


def test_kpt_model_builder_matches_reference_fixture(golden):
    """build_kpt_chirp_model (models.py:522-580) and its measurement function on the host."""
    import chirpgp_b200 as cg
    for name in ('kpt', 'kpt_h2'):
        z = golden(name)
        nh = int(z['num_harmonics'])
        F, Sigma, m0, P0, h = cg.build_kpt_chirp_model(z['params'], float(z['fs']), nh)
        for a, b in ((F, z['F']), (Sigma, z['Sigma']), (m0, z['m0']), (P0, z['P0'])):
            npt.assert_allclose(a.numpy(), b, rtol=1e-15, atol=0)
        npt.assert_allclose(float(h(m0)), float(z['h_at_m0']), rtol=1e-15)
        # batched parameters: one model per chirp
        Fb, Sb, m0b, P0b, _ = cg.build_kpt_chirp_model(np.tile(z['params'], (3, 1)), float(z['fs']), nh)
        assert Sb.shape == (3, nh + 2, nh + 2) and m0b.shape == (3, nh + 2) and P0b.shape == (3, nh + 2, nh + 2)
