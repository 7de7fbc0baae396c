#!/usr/bin/env python
"""One fused filter + gains launch at B chirps (for ncu captures of the large-batch kernel):  python profiles/scripts/run_oct.py [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
T, DT, XI = 3141, 1e-3, 0.1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device('cuda', 0)
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
sg = cg.SigmaPoints.gauss_hermite(4, 3)
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
_, ys, _ = toymodels.synthetic_batch(min(B, 1000), T, DT, Xi=XI, seed=2)
ys = torch.as_tensor(np.tile(ys, (-(-B // ys.shape[0]), 1))[:B]).to(dev)
for it in range(2):
    f = cg.sgp_filter(mc, sg, H, XI, m0, P0, DT, ys)
torch.cuda.synchronize()
print(float(f[2][:, -1].sum()))
