"""Config 4 (harmonic chirps d = 8, cubature): filter (+ gains) and sweep timings at several batch sizes.
   python profiles/scripts/time_cub.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels

PARAMS = np.array([0.1, 0.1, 0.1, 1., 1., 7.])


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    Bs = [int(a) for a in sys.argv[1:]] or [500, 1000, 2000]
    T, dt = 3141, 1e-3
    _, ys_all, _ = toymodels.synthetic_batch(max(Bs), T, dt, Xi=0.1, num_harmonics=3, seed=4)
    drift, disp, mc, m0, P0, H = cg.build_harmonic_chirp_model(PARAMS, num_harmonics=3)
    sg = cg.SigmaPoints.cubature(8)
    Hd, m0d, P0d = H.cuda(), m0.cuda(), P0.cuda()
    for B in Bs:
        ys = torch.as_tensor(ys_all[:B]).cuda()
        f = cg.sgp_filter(mc, sg, Hd, 0.1, m0d, P0d, dt, ys)
        t_fused = timed(lambda: cg.sgp_filter(mc, sg, Hd, 0.1, m0d, P0d, dt, ys))
        t_plain = timed(lambda: cg.sgp_filter(mc, sg, Hd, 0.1, m0d, P0d, dt, ys, smoother_gains=False))
        t_sweep = timed(lambda: cg.sgp_smoother(mc, sg, f[0], f[1], dt))
        fc, Pc = f[0].clone(), f[1].clone()
        t_alone = timed(lambda: cg.sgp_smoother(mc, sg, fc, Pc, dt))
        cyc = t_fused * 1e-3 * 1.965e9 / T
        print('B=%6d  filter+gains %.3f ms (%.0f cycles per step)  plain filter %.3f ms  sweep %.3f ms  stand-alone smoother %.3f ms'
              '   pair %.3e steps/s' % (B, t_fused, cyc, t_plain, t_sweep, t_alone, B * T / ((t_fused + t_sweep) * 1e-3)), flush=True)


if __name__ == '__main__':
    main()
