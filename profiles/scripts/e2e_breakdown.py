#!/usr/bin/env python
"""Where the time of one blocking host->host call goes (config 2): python profiles/scripts/e2e_breakdown.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels

B, T, DT, XI = 1000, 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
_, ys, _ = toymodels.synthetic_batch(B, T, DT, Xi=XI, seed=2)
ys_host = torch.as_tensor(ys).pin_memory()
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
sg = cg.SigmaPoints.gauss_hermite(4, 3)
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)


def t(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


print('pinned alloc 25 MB      %.3f ms' % t(lambda: torch.empty((B, T), dtype=torch.float64, pin_memory=True)))
print('pinned alloc 402 MB     %.3f ms' % t(lambda: torch.empty((B, T, 4, 4), dtype=torch.float64, pin_memory=True)))
keep = []
print('pinned alloc 402 MB, previous kept alive %.3f ms' % t(lambda: keep.append(torch.empty((B, T, 4, 4), dtype=torch.float64, pin_memory=True)) or keep[-3:] and None))
del keep
ysd = ys_host.to(dev)
print('H2D ys                  %.3f ms' % t(lambda: ys_host.to(dev, non_blocking=True)))
print('filter+smoother device  %.3f ms' % t(lambda: cg.sgp_filter_smoother(mc, sg, H, XI, m0, P0, DT, ysd)))
print('  + readout freq,v_var (device) %.3f ms' % t(lambda: cg.sgp_filter_smoother(mc, sg, H, XI, m0, P0, DT, ysd, readout=('freq', 'v_var'))))
print('host ys -> readout host %.3f ms' % t(lambda: cg.sgp_filter_smoother(mc, sg, H, XI, m0, P0, DT, ys_host, readout=('freq', 'v_var'))))
print('host ys -> mss,Pss host %.3f ms' % t(lambda: cg.sgp_filter_smoother(mc, sg, H, XI, m0, P0, DT, ys_host, readout=('mss', 'Pss'))))
f = cg.sgp_filter_smoother(mc, sg, H, XI, m0, P0, DT, ysd)
buf = torch.empty(f[4].shape, dtype=torch.float64, pin_memory=True)
print('bare D2H Pss 402 MB into an existing pinned buffer %.3f ms' % t(lambda: buf.copy_(f[4], non_blocking=True)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    cg.sgp_filter_smoother(mc, sg, H, XI, m0, P0, DT, ys_host, readout=('freq', 'v_var'))
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
