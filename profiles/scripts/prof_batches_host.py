import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
os.environ['CGP_GH_OCT'] = '1'
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
def run(n, depth=8):
    for out in cg.filter_smoother_batches(cg.sgp_filter_smoother, *args, batches=(hosts[i % 3] for i in range(n)), readout=('freq', 'v_var'), depth=depth):
        pass
run(32); run(32)
pr = cProfile.Profile(); pr.enable(); run(48); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
