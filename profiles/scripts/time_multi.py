#!/usr/bin/env python
"""Experiment: NCH chirps interleaved per warp in the Gauss-Hermite filter (csrc/cgp_multi.cuh), nll-only mode.
    python profiles/scripts/time_multi.py [B ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels, mle

T, DT, XI = 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
sg = cg.SigmaPoints.gauss_hermite(4, 3)
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
for B in [int(a) for a in sys.argv[1:]] or [1000]:
    _, ys, _ = toymodels.synthetic_batch(min(B, 1000), T, DT, Xi=XI, seed=2)
    ys = torch.as_tensor(np.tile(ys, (-(-B // ys.shape[0]), 1))[:B]).to(dev)
    ref = None
    for nch in ('', '16'):
        os.environ['CGP_GH_MULTI'] = nch
        best = 1e9
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            v = mle.filter_nll('sgp_filter', (mc,), H, XI, m0, P0, DT, ys, sgps=sg)
            e1.record(); torch.cuda.synchronize()
            if it: best = min(best, e0.elapsed_time(e1))
        if ref is None: ref = v.clone()
        err = float(((v - ref).abs() / ref.abs()).max())
        cyc = best * 1e-3 * 1.965e9 / T
        print('B=%6d chirps/warp=%-8s %8.3f ms  %7.0f cycles per warp-step  %6.3f G filter steps/s  max rel diff vs plain kernel %.1e'
              % (B, nch or 'plain', best, cyc, B * T / best / 1e6, err), flush=True)
