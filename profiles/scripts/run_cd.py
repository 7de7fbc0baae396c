#!/usr/bin/env python
"""One cd_ekf + cd_eks pass (chirp SDE, half-warp-per-chirp kernels) at B chirps, for ncu captures:  python profiles/scripts/run_cd.py [B]"""
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
_, ys, _ = toymodels.synthetic_batch(min(B, 1000), T, DT, Xi=XI, seed=2)
ys = torch.as_tensor(np.tile(ys, (-(-B // ys.shape[0]), 1))[:B]).to(dev)
for it in range(2):
    f = cg.cd_ekf(drift, disp, H, XI, m0, P0, DT, ys)
    s = cg.cd_eks(drift, disp, f[0], f[1], DT)
torch.cuda.synchronize()
print(float(f[2][:, -1].sum()), float(s[0].sum()))
