"""Monte-Carlo input side on the device: the in-kernel Philox4x32-10 / Box-Muller generator against the
Random123 known-answer vectors and the NumPy restatement (oracle/sim_oracle.py), simulated trajectories against that
restatement sample by sample, and the reference's own Monte-Carlo test (test/test_crlb.py:19-73) on the CUDA path."""
import ctypes as C
import math

import numpy as np
import numpy.testing as npt
import pytest
import scipy.linalg
import torch

import chirpgp_b200 as cg
from chirpgp_b200 import _native as N
from chirpgp_b200 import tools
from oracle import oracle as orc
from oracle import sim_oracle as so

pytestmark = pytest.mark.gpu


def test_philox_known_answers_and_normals():
    L = N.lib()
    out = torch.zeros(4, dtype=torch.int32, device='cuda')
    for ctr, key, want in so.KAT:
        assert L.cgp_test_philox(*[C.c_uint32(v) for v in (key + ctr)], C.c_void_p(out.data_ptr()), None) == 0
        got = [int(v) & 0xFFFFFFFF for v in out.cpu().tolist()]
        assert got == list(want)
    n, seed = 5000, 0x1234567890ABCDEF
    z = torch.empty(2 * n, dtype=torch.float64, device='cuda')
    assert L.cgp_test_normals(C.c_uint64(seed), n, C.c_void_p(z.data_ptr()), None) == 0
    i = np.arange(n, dtype=np.uint64)
    z0, z1 = so.normal2(seed, i & np.uint64(3), i >> np.uint64(2), i * np.uint64(7919))
    got = z.cpu().numpy().reshape(n, 2)
    npt.assert_allclose(got[:, 0], z0, rtol=1e-13, atol=1e-15)       # libm vs CUDA log / sincospi: a few ulp
    npt.assert_allclose(got[:, 1], z1, rtol=1e-13, atol=1e-15)


def test_simulated_trajectories_match_the_restatement():
    """Chirp LCD model and a linear model: x0, xs, ys sample by sample; independence of the batch split."""
    params = np.array([0.1, 0.1, 0.1, 1., 1., 7.])
    _, _, mc, m0, P0, H = cg.build_chirp_model(params)
    spec = orc.ChirpSpec(0.1, 0.1, 1., 1.)
    dt, T, B, seed = 0.01, 40, 64, 666
    x0, xs, ys = tools.simulate(mc, H, 0.1, m0, P0, dt, T, B, seed)
    mean_fn = lambda x: np.stack([orc.disc_mean_cov(spec, dt, u, want_jac=False)[0] for u in x])
    Sigma = orc.disc_mean_cov(spec, dt, np.zeros(4), want_jac=False)[-1]
    ox0, oxs, oys = so.simulate(mean_fn, Sigma, H.numpy(), 0.1, m0.numpy(), P0.numpy(), T, B, seed)
    npt.assert_allclose(x0.cpu().numpy(), ox0, rtol=1e-12, atol=1e-13)
    npt.assert_allclose(xs.cpu().numpy(), oxs, rtol=1e-10, atol=1e-11)
    npt.assert_allclose(ys.cpu().numpy(), oys, rtol=1e-10, atol=1e-11)
    # trajectory i depends on (seed, i) only: two half batches == one batch, bit for bit
    a = tools.simulate(mc, H, 0.1, m0, P0, dt, T, B // 2, seed)
    b = tools.simulate(mc, H, 0.1, m0, P0, dt, T, B // 2, seed, first_trajectory=B // 2)
    assert torch.equal(torch.cat([a[2], b[2]]), ys) and torch.equal(torch.cat([a[1], b[1]]), xs)
    assert not torch.equal(tools.simulate(mc, H, 0.1, m0, P0, dt, T, B, seed + 1)[2], ys)
    # reference-named single-trajectory helpers (tools.py:81-170)
    F = np.array([[0.9, 0.1], [0., 0.8]]); Sig = np.diag([0.1, 0.2])
    tr = tools.simulate_lgssm(F, Sig, np.array([1., -1.]), 30, 5)
    _, oxs2, _ = so.simulate(lambda x: x @ F.T, Sig, np.zeros(2), 0., np.array([1., -1.]), np.zeros((2, 2)), 30, 1, 5)
    npt.assert_allclose(tr.cpu().numpy(), oxs2[0], rtol=1e-12, atol=1e-13)
    assert tuple(tools.simulate_sde(mc, m0, P0, dt, 25, 9).shape) == (25, 4)


def test_crlb_lgssm_monte_carlo():
    """test/test_crlb.py:19-73 on the CUDA path: 10^6 simulated trajectories of the Matern-3/2 LGSSM, batched kf;
    covariances bit-identical across the batch (:64-66) and the Monte-Carlo error covariance ~ Pf (atol 1e-1, :71-73).
    and inv(PCRLB) == Pf (atol 1e-12, :75-87) with posterior_cramer_rao (models.py:583-644)."""
    ell, sigma, dt, T = 1., 1., 0.1, 10
    A = np.array([[0., 1.], [-3 / ell ** 2, -2 * math.sqrt(3) / ell]])
    Bv = np.array([[0.], [2 * sigma * (math.sqrt(3) / ell) ** 1.5]])
    F = scipy.linalg.expm(A * dt)                                    # lti_sde_to_disc, tools.py:44-78 (Van Loan)
    d = 2
    M = np.zeros((2 * d, 2 * d)); M[:d, :d] = A; M[:d, d:] = Bv @ Bv.T; M[d:, d:] = -A.T
    E = scipy.linalg.expm(M * dt)
    Sigma = E[:d, d:] @ F.T
    Xi, H = 1., np.array([1., 0.])
    m0, P0 = np.zeros(2), np.diag([sigma ** 2, 3 / ell ** 2 * sigma ** 2])
    num_mcs = 1000000
    x0, xs, ys = tools.simulate(cg.LinearDisc(F, Sigma), H, Xi, m0, P0, dt, T, num_mcs, 666)
    mfs, Pfs, _ = cg.kf(F, Sigma, H, Xi, m0, P0, ys)
    assert torch.equal(Pfs[3], Pfs[77]) and torch.equal(Pfs[0], Pfs[-1])
    res = mfs - xs
    Emc = torch.einsum('bti,btj->tij', res, res) / num_mcs
    npt.assert_allclose(Emc.cpu().numpy(), Pfs[0].cpu().numpy(), atol=1e-1)
    npt.assert_allclose(Emc.cpu().numpy(), Pfs[0].cpu().numpy(), rtol=2e-2, atol=2e-3)     # what 10^6 samples actually give
    # x0 ~ N(m0, P0)
    npt.assert_allclose(np.cov(x0.cpu().numpy().T), P0, atol=2e-2)
    # :75-87: the PCRLB of the linear Gaussian model is the Kalman covariance
    Ft, St, Ht = torch.as_tensor(F).cuda(), torch.as_tensor(Sigma).cuda(), torch.as_tensor(H).cuda()
    Sinv = torch.linalg.inv(St)
    logdet = torch.logdet(St)

    def logpdf_transition(xt, xs):
        r = xt - Ft @ xs
        return -0.5 * (r @ Sinv @ r) - 0.5 * logdet - math.log(2 * math.pi)

    def logpdf_likelihood(yt, xt):
        return -0.5 * (yt - Ht @ xt) ** 2 / Xi - 0.5 * math.log(2 * math.pi * Xi)

    n = 2000                                                   # the Hessians are constant for this model: no MC noise
    xss = torch.cat([x0[:n, None], xs[:n]], dim=1).transpose(0, 1).contiguous()
    js = cg.posterior_cramer_rao(xss, ys[:n].T.contiguous(), torch.linalg.inv(torch.as_tensor(P0).cuda()),
                                 logpdf_transition, logpdf_likelihood)
    npt.assert_allclose(torch.linalg.inv(js).cpu().numpy(), Pfs[0].cpu().numpy(), atol=1e-12)
