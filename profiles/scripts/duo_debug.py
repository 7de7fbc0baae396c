#!/usr/bin/env python
"""Debug build only (-DCGP_DUO_DEBUG): cycles the producer waits for buffers / the consumer waits for steps / spends flushing."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
os.environ['CGP_DUO'] = '2'; os.environ['CGP_DUO_SEL'] = '4'
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
T = 3141
dev = torch.device('cuda', 0)
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
sg = cg.SigmaPoints.gauss_hermite(4, 3)
_, ys, _ = toymodels.synthetic_batch(1000, T, 1e-3, Xi=0.1, seed=2)
ys = torch.as_tensor(ys).to(dev)
for gains in (False, True):
    for _ in range(2):
        f = cg.sgp_filter(mc, sg, H.to(dev), 0.1, m0.to(dev), P0.to(dev), 1e-3, ys, smoother_gains=gains)
    d = f[0][:, 0, :].cpu().numpy()
    print('gains=%d per step: producer waits %.0f of %.0f cycles; consumer waits %.0f, flushes %.0f (per flush %.0f)'
          % (gains, d[:, 0].mean() / T, d[:, 1].mean() / T, d[:, 2].mean() / T, d[:, 3].mean() / T, d[:, 3].mean() / (T / 32)))
    print('   producer total cycles/step: min %.0f median %.0f max %.0f' % (d[:, 1].min() / T, np.median(d[:, 1]) / T, d[:, 1].max() / T))
