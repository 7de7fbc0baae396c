#!/usr/bin/env python
"""Large-batch Gauss-Hermite kernel (csrc/cgp_oct.cuh, 8 lanes per chirp) against the warp-per-chirp kernels:
    python profiles/scripts/time_oct.py [B ...]
Per batch size: filter + gains and sweep (fused pair), filter only, nll only; CUDA events, best of 3 after a warm-up."""
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


def timed(fn, reps=3):
    best, out = 1e9, None
    for it in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record(); torch.cuda.synchronize()
        if it: best = min(best, e0.elapsed_time(e1))
    return best, out


for B in [int(a) for a in sys.argv[1:]] or [1000, 4000, 10000]:
    _, ys, _ = toymodels.synthetic_batch(min(B, 1000), T, DT, Xi=XI, seed=2)
    ys = torch.as_tensor(np.tile(ys, (-(-B // ys.shape[0]), 1))[:B]).to(dev)
    ref = {}
    for oct_ in ('0', '1'):
        os.environ['CGP_GH_OCT'] = oct_
        tf, f = timed(lambda: cg.sgp_filter(mc, sg, H, XI, m0, P0, DT, ys))
        ts, s = timed(lambda: cg.sgp_smoother(mc, sg, f[0], f[1], DT))
        tp, fp = timed(lambda: cg.sgp_filter(mc, sg, H, XI, m0, P0, DT, ys, smoother_gains=False))
        tn, v = timed(lambda: mle.filter_nll('sgp_filter', (mc,), H, XI, m0, P0, DT, ys, sgps=sg))
        res = dict(mf=f[0], Pf=f[1], nell=f[2][..., -1], ms=s[0], Ps=s[1], nll=v, mf_plain=fp[0])
        if oct_ == '0':
            ref = {k: x.clone() for k, x in res.items()}
            diff = ''
        else:
            diff = '  max |diff| vs warp-per-chirp: ' + ' '.join(
                '%s %.1e' % (k, float((res[k] - ref[k]).abs().max())) for k in ('mf', 'Pf', 'ms', 'Ps')) + \
                ' nll rel %.1e' % float(((res['nll'] - ref['nll']).abs() / ref['nll'].abs()).max())
        print('B=%6d oct=%s  filter+gains %8.3f ms  sweep %7.3f ms  pair %.3f G steps/s | filter only %8.3f ms (%.3f G) | nll only %8.3f ms (%.3f G)%s'
              % (B, oct_, tf, ts, B * T / (tf + ts) / 1e6, tp, B * T / tp / 1e6, tn, B * T / tn / 1e6, diff), flush=True)
        del f, s, fp, v, res
    del ref
    torch.cuda.empty_cache()
os.environ.pop('CGP_GH_OCT', None)
