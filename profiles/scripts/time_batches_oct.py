"""Batch sequences of config 2 (1000 chirps x 3141 per batch) with the large-batch kernel forced (CGP_GH_OCT=1: 8 lanes per chirp, fewer FP64
instructions per chirp and step but a longer chain per warp) against the warp-pair kernel, over the number of batches in flight."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
B, T, dt, Xi = 1000, 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
hosts = [torch.as_tensor(toymodels.synthetic_batch(B, T, dt, Xi=Xi, seed=s)[1]).pin_memory() for s in (1, 2, 3)]
devs = [h.to(dev) for h in hosts]
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
sg = cg.SigmaPoints.gauss_hermite(d=4, order=3)
args = (mc, sg, H, Xi, m0, P0, dt)
def run(src, readout, n, depth):
    for out in cg.filter_smoother_batches(cg.sgp_filter_smoother, *args, batches=(src[i % 3] for i in range(n)), readout=readout, depth=depth):
        pass
for oct_ in ('0', '1'):
    os.environ['CGP_GH_OCT'] = oct_
    for depth in (2, 3, 4, 6, 8):
        for name, src, ro in (('dev  none', devs, None), ('host freq,v_var', hosts, ('freq', 'v_var'))):
            run(src, ro, 4 * depth, depth); run(src, ro, 4 * depth, depth)
            torch.cuda.synchronize(); t0 = time.perf_counter(); run(src, ro, 48, depth); torch.cuda.synchronize()
            print('oct=%s depth %d %-16s %.3f ms per batch' % (oct_, depth, name, (time.perf_counter() - t0) * 1e3 / 48), flush=True)
