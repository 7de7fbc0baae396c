"""Allocator behaviour on the GPU box when the previous result is still alive during the next call."""
import os, sys, time, torch
import numpy as np
dev = torch.device('cuda', 0)
def t(fn, n=4):
    r = fn(); r = fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return ' '.join('%.3f' % x for x in ts)
print('A pinned 25MB :', t(lambda: torch.empty((1000, 3141), dtype=torch.float64, pin_memory=True)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import chirpgp_b200 as cg
print('B after import chirpgp_b200 :', t(lambda: torch.empty((1000, 3141), dtype=torch.float64, pin_memory=True)))
from chirpgp_b200 import toymodels
_, ys, _ = toymodels.synthetic_batch(1000, 3141, 1e-3, Xi=0.1, seed=2)
print('C after synthetic_batch :', t(lambda: torch.empty((1000, 3141), dtype=torch.float64, pin_memory=True)))
ys_host = torch.as_tensor(ys).pin_memory()
print('D after .pin_memory() :', t(lambda: torch.empty((1000, 3141), dtype=torch.float64, pin_memory=True)))
_, _, mc, m0, P0, H = cg.build_chirp_model(np.array([0.1, 0.1, 0.1, 1., 1., 7.]))
sg = cg.SigmaPoints.gauss_hermite(4, 3)
m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
print('E after model to dev :', t(lambda: torch.empty((1000, 3141), dtype=torch.float64, pin_memory=True)))
ysd = ys_host.to(dev)
r = cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ysd, readout=('freq',))
print('F after a kernel call :', t(lambda: torch.empty((1000, 3141), dtype=torch.float64, pin_memory=True)))
print('G cuda 402MB :', t(lambda: torch.empty((1000, 3141, 16), dtype=torch.float64, device=dev)))
print('H call device readout :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ysd, readout=('freq',))))
print('I call device full :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ysd)))
print('J call host readout :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ys_host, readout=('freq', 'v_var'))))
# ---- the same call under bench.py's clock sampler (nvidia-smi -lms 25 in the background)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench
s = bench.ClockSampler(0); s.start(); time.sleep(0.5)
print('K call host readout, nvidia-smi -lms 25 running :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ys_host, readout=('freq', 'v_var')), n=6))
print('L call device full, sampler running :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ysd), n=6))
print(s.stop())
print('M call host readout, sampler stopped :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ys_host, readout=('freq', 'v_var')), n=6))
print('N host full (mss,Pss) :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ys_host, readout=('mss', 'Pss')), n=4))
ys_np = ys_host.numpy()
def numpy_api():
    f = cg.sgp_filter(mc, sg, H.cpu().numpy(), 0.1, m0.cpu().numpy(), P0.cpu().numpy(), 1e-3, ys_np)
    return cg.sgp_smoother(mc, sg, f[0], f[1], 1e-3)
print('O numpy api two calls :', t(numpy_api, n=3))
print('P numpy one call full :', t(lambda: cg.sgp_filter_smoother(mc, sg, H, 0.1, m0, P0, 1e-3, ys_np), n=3))
