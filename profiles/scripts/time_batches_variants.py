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
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 3
def run(src, readout, n, depth=DEPTH):
    for out in cg.filter_smoother_batches(cg.sgp_filter_smoother, *args, batches=(src[i % 3] for i in range(n)), readout=readout, depth=depth):
        pass
for name, src, ro in (('dev  none', devs, None), ('dev  freq,v_var', devs, ('freq', 'v_var')), ('host v_mean', hosts, ('v_mean',)),
                      ('host v_mean,v_var', hosts, ('v_mean', 'v_var')), ('host freq', hosts, ('freq',)), ('host freq,v_var', hosts, ('freq', 'v_var')),
                      ('host n_ell_last', hosts, ('n_ell_last',))):
    run(src, ro, 4 * DEPTH); run(src, ro, 4 * DEPTH)
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(src, ro, 80); torch.cuda.synchronize()
    print('%-20s %.3f ms per batch' % (name, (time.perf_counter() - t0) * 1e3 / 80))
