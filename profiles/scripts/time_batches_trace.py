"""Per-batch wall-clock trace of filter_smoother_batches (config 2, pinned host in / out): transient vs steady state."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
B, T, dt, Xi = 1000, 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
hosts = [torch.as_tensor(toymodels.synthetic_batch(B, T, dt, Xi=Xi, seed=s)[1]).pin_memory() for s in (1, 2, 3)]
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
sg = cg.SigmaPoints.gauss_hermite(d=4, order=3)
args = (mc, sg, H, Xi, m0, P0, dt)
for readout in (('freq', 'v_var'), ('mss', 'Pss')):
    for depth in (2, 3):
        for rep in range(2):
            n = 30
            ts = []
            t0 = time.perf_counter()
            for out in cg.filter_smoother_batches(cg.sgp_filter_smoother, *args, batches=(hosts[i % 3] for i in range(n)),
                                                  readout=readout, depth=depth):
                ts.append(time.perf_counter())
            d = np.diff(np.array([t0] + ts)) * 1e3
            print(readout, 'depth', depth, 'rep', rep, ' '.join('%.1f' % x for x in d), '| mean of last 15: %.3f ms' % d[-15:].mean())
