"""Discrete EKF (chirp LCD model, d = 4): 16 lanes per chirp (ekf_lane_kernel, CGP_EKF_LANE=1) against one thread per chirp
(ekf_thread_kernel, CGP_EKF_LANE=0) over the batch size, T = 3141.   python profiles/scripts/time_ekf_lane.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
T, dt, Xi = 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = ev(), ev(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
base = toymodels.synthetic_batch(1000, T, dt, Xi=Xi, seed=2)[1]
for B in [int(a) for a in sys.argv[1:]] or [1, 100, 1000, 4000, 8192, 16384, 32768]:
    ys = torch.as_tensor(np.tile(base, (-(-B // 1000), 1))[:B]).to(dev)
    res, tt = {}, {}
    for lane in ('0', '1'):
        os.environ['CGP_EKF_LANE'] = lane
        res[lane] = cg.ekf(mc, H, Xi, m0, P0, dt, ys)
        tt[lane] = timed(lambda: cg.ekf(mc, H, Xi, m0, P0, dt, ys))
    d = [float((a - b).abs().max()) for a, b in zip(res['0'], res['1'])]
    print('B=%6d  thread per chirp %.3f ms   16 lanes per chirp %.3f ms   max |diff| mf %.1e Pf %.1e nell %.1e' % (B, tt['0'], tt['1'], *d), flush=True)
