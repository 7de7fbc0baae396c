#!/usr/bin/env python
"""One cd_sgp_filter + cd_sgp_smoother pass (chirp SDE, Gauss-Hermite order 3, warp-per-chirp kernels) at B chirps, for ncu
captures:  python profiles/scripts/run_cd_gh.py [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
T, DT, XI = 3141, 1e-3, 0.1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device('cuda', 0)
drift, disp, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
sg = cg.SigmaPoints.gauss_hermite(4, 3)
_, ys, _ = toymodels.synthetic_batch(min(B, 1000), T, DT, Xi=XI, seed=2)
ys = torch.as_tensor(np.tile(ys, (-(-B // ys.shape[0]), 1))[:B]).to(dev)
for it in range(2):
    f = cg.cd_sgp_filter(drift, disp(None), sg, H, XI, m0, P0, DT, ys)
    s = cg.cd_sgp_smoother(drift, disp(None), sg, f[0], f[1], DT)
torch.cuda.synchronize()
print(float(f[2][:, -1].sum()), float(s[0].sum()))
