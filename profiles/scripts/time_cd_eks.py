"""cd_ekf + cd_eks (chirp SDE, half-warp-per-chirp kernels) at 1000 x 3141: the two schedules of the smoother's RK4 step
(CGP_EKS_STRAIGHT=0: range branch per stage; 1: straight-line step, side chosen per step) on the bench data (frequency offset 8:
V crosses the softplus split) and on a data set that stays on the series side (offset 14)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import numpy as np, torch
import chirpgp_b200 as cg
from chirpgp_b200 import toymodels
B, T, dt, Xi = 1000, 3141, 1e-3, 0.1
dev = torch.device('cuda', 0)
drift, disp, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = ev(), ev(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for offset in (8., 14.):
    ys = torch.as_tensor(toymodels.synthetic_batch(B, T, dt, Xi=Xi, seed=2, offset=offset)[1]).to(dev)
    f = cg.cd_ekf(drift, disp, H, Xi, m0, P0, dt, ys)
    tf = timed(lambda: cg.cd_ekf(drift, disp, H, Xi, m0, P0, dt, ys))
    res = {}
    for straight in ('0', '1'):
        os.environ['CGP_EKS_STRAIGHT'] = straight
        res[straight] = cg.cd_eks(drift, disp, f[0], f[1], dt)
        ts = timed(lambda: cg.cd_eks(drift, disp, f[0], f[1], dt))
        print('offset %4.1f  V in [%.2f, %.2f]  cd_ekf %.3f ms   cd_eks straight=%s %.3f ms' % (offset, float(f[0][..., 2].min()), float(f[0][..., 2].max()), tf, straight, ts))
    print('   max |diff| between the schedules: ms %.2e  Ps %.2e' % (float((res['0'][0] - res['1'][0]).abs().max()), float((res['0'][1] - res['1'][1]).abs().max())))
