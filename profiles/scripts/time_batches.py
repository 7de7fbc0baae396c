"""End-to-end (pinned host in -> pinned host out) rate of config 2 over a sequence of batches: blocking product calls one
after the other vs cg.filter_smoother_batches with 2 / 3 batches in flight.  python profiles/scripts/time_batches.py [B] [n]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
T, dt, Xi = 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
hosts = [torch.as_tensor(toymodels.synthetic_batch(B, T, dt, Xi=Xi, seed=s)[1]).pin_memory() for s in (1, 2, 3)]
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
sg = cg.SigmaPoints.gauss_hermite(d=4, order=3)
args = (mc, sg, H, Xi, m0, P0, dt)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    t0 = time.perf_counter()
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


for readout in (('freq', 'v_var'), ('mss', 'Pss')):
    def blocking(k):
        for i in range(k):
            out = cg.sgp_filter_smoother(*args, hosts[i % 3], readout=readout)
        return out
    ms, wall = timed(blocking)
    print('readout=%-18s blocking calls            %.3f ms per batch (wall %.3f)  %.1f M steps/s' % (readout, ms, wall, B * T / ms / 1e3))
    for depth in (1, 2, 3, 4):
        def piped(k):
            for out in cg.filter_smoother_batches(cg.sgp_filter_smoother, *args, batches=(hosts[i % 3] for i in range(k)),
                                                  readout=readout, depth=depth):
                pass
            return out
        ms, wall = timed(piped)
        print('readout=%-18s filter_smoother_batches depth %d  %.3f ms per batch (wall %.3f)  %.1f M steps/s'
              % (readout, depth, ms, wall, B * T / ms / 1e3))
