"""A Monte-Carlo study as a SEQUENCE of batches: what tetralith/jobs/ghfs_mle.py:26-86 does one run at a time
(`for mc in range(num_mcs)`: draw a noisy chirp, filter, smooth, E[g(V)], RMSE), here with 1000 runs per batch and the batches
handed to `cg.filter_smoother_batches`, which keeps several of them in flight on the GPU and yields, in order, only the
requested read-outs (posterior frequency estimate + its marginal variance, 16 bytes per step) in pinned host memory.

    python demos/mc_batches.py [ghfs|ekfs|cd_ekfs] [--batches 24] [--chirps 1000] [--depth 8]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chirpgp_b200 as cg  # noqa: E402
from chirpgp_b200 import toymodels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('method', nargs='?', default='ghfs', choices=['ghfs', 'ekfs', 'cd_ekfs'])
    ap.add_argument('--batches', type=int, default=24)
    ap.add_argument('--chirps', type=int, default=1000)
    ap.add_argument('--depth', type=int, default=8)
    a = ap.parse_args()
    dt, T, Xi = 1e-3, 3141, 0.1                                   # demos/ghfs_mle.py:16-18
    params = np.array([0.1, 0.1, 0.1, 1., 1., 7.])                # lam, b, delta, ell, sigma, m0 of V
    drift, disp, m_and_cov, m0, P0, H = cg.build_chirp_model(params)
    dev = torch.device('cuda', 0)
    m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
    if a.method == 'ghfs':
        pair, args = cg.sgp_filter_smoother, (m_and_cov, cg.SigmaPoints.gauss_hermite(d=4, order=3), H, Xi, m0, P0, dt)
    elif a.method == 'ekfs':
        pair, args = cg.ekf_smoother, (m_and_cov, H, Xi, m0, P0, dt)
    else:
        pair, args = cg.cd_ekf_smoother, (drift, disp, H, Xi, m0, P0, dt)

    # Measurement batches: a few distinct ones drawn up front into pinned host memory (the filter kernel of the fused pair reads
    # pinned measurements in place) and cycled -- a real study would draw every batch, e.g. with chirpgp_b200.tools.simulate
    pool, truth = [], None
    for s in range(min(a.batches, 4)):
        _, ys, truth = toymodels.synthetic_batch(a.chirps, T, dt, Xi=Xi, seed=100 + s)
        pool.append(torch.as_tensor(ys).pin_memory())
    truth_t = torch.as_tensor(truth)

    def study(n):
        """n batches through the pair; per-run RMSE of the frequency estimate (chirpgp/tools.py:279-293) on the host."""
        rm, first, v_var = [], None, None
        t0 = time.perf_counter()
        for freq, v_var in cg.filter_smoother_batches(pair, *args, batches=(pool[i % len(pool)] for i in range(n)),
                                                      readout=('freq', 'v_var'), depth=a.depth):
            if first is None:
                first = time.perf_counter() - t0
            rm.append((freq - truth_t).square_().mean(dim=1).sqrt_())
        return torch.cat(rm).numpy(), first, time.perf_counter() - t0, v_var

    # the first sequence of a process allocates every stream's device pool and the pinned result blocks (cudaMalloc /
    # cudaHostAlloc synchronise): a short one up front, then the study itself
    _, first_cold, wall_cold, _ = study(2 * a.depth)
    torch.cuda.synchronize()
    r, first, wall, v_var = study(a.batches)
    print('%s: %d batches x %d chirps x %d steps, %d in flight: %.2f ms per batch host to host incl. the RMSE on the host (first result '
          'after %.1f ms), %.2e smoothed steps/s;  cold start before it: %d batches in %.0f ms'
          % (a.method, a.batches, a.chirps, T, a.depth, wall * 1e3 / a.batches, first * 1e3, a.batches * a.chirps * T / wall,
             2 * a.depth, wall_cold * 1e3))
    print('RMSE of E[g(V)] against the true frequency over %d runs: mean %.3f Hz, median %.3f Hz, worst %.3f Hz; '
          'mean posterior std of V %.3f' % (r.size, r.mean(), np.median(r), r.max(), float(np.sqrt(v_var.numpy()).mean())))


if __name__ == '__main__':
    main()
