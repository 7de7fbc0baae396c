"""Monte-Carlo error of the EKF on the chirp model -- the reference's largest batched job
(/root/reference/tetralith/jobs/crlb_ekf.py: 10^6 simulated trajectories x 500 steps, vmap(ekf), mean / std of the squared
errors of the chirp and frequency-state estimates per time step), entirely on the device:
trajectories and measurements from the in-kernel Philox generator (chirpgp_b200.tools.simulate), batched EKF, reductions.

    python demos/crlb_ekf.py [-lam 0.1 -b 0.1 -delta 0.1 -ell 1 -sigma 1 -Xi 0.1] [--num-mcs 1000000] [--chunk 250000]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chirpgp_b200 as cg  # noqa: E402
from chirpgp_b200 import tools  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    for name, default in (('lam', 0.1), ('b', 0.1), ('delta', 0.1), ('ell', 1.), ('sigma', 1.), ('Xi', 0.1)):
        ap.add_argument('-' + name, type=float, default=default)
    ap.add_argument('--num-mcs', type=int, default=1000000)
    ap.add_argument('--chunk', type=int, default=250000)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    _, _, m0, P0, H = cg.model_chirp(a.lam, a.b, a.ell, a.sigma, a.delta)         # crlb_ekf.py:28-29
    m_and_cov = cg.disc_chirp_lcd(a.lam, a.b, a.ell, a.sigma)
    dt, T = 0.01, 500                                                             # :31-33
    s1 = torch.zeros((2, T), dtype=torch.float64, device='cuda')                  # sum of squared errors (chirp, v)
    s2 = torch.zeros((2, T), dtype=torch.float64, device='cuda')                  # sum of their squares
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    while done < a.num_mcs:
        n = min(a.chunk, a.num_mcs - done)
        _, xs, ys = tools.simulate(m_and_cov, H, a.Xi, m0, P0, dt, T, n, seed=666, first_trajectory=done)   # :41-64
        mfs, _, _ = cg.ekf(m_and_cov, H, a.Xi, m0, P0, dt, ys)                                             # :68-79
        e = (mfs[:, :, 1:3] - xs[:, :, 1:3]) ** 2                                                          # :82-90
        s1 += e.sum(0).T
        s2 += (e ** 2).sum(0).T
        done += n
        del xs, ys, mfs, e
    mean = s1 / a.num_mcs
    std = torch.sqrt(torch.clamp(s2 / a.num_mcs - mean ** 2, min=0.))
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    mean, std = mean.cpu().numpy(), std.cpu().numpy()
    print('%d trajectories x %d steps simulated + filtered in %.2f s (%.2f G steps/s incl. simulation and reductions)'
          % (a.num_mcs, T, sec, a.num_mcs * T / sec / 1e9))
    print('squared error of the chirp state: mean %.4g (t = 1) ... %.4g (t = T);   of the frequency state: %.4g ... %.4g'
          % (mean[0, 0], mean[0, -1], mean[1, 0], mean[1, -1]))
    if a.out:
        np.savez(a.out, ts=np.linspace(dt, T * dt, T), err_mean_chirps=mean[0], err_std_chirps=std[0], err_mean_vs=mean[1],
                 err_std_vs=std[1])                                                                        # :92-95


if __name__ == '__main__':
    main()
