#!/usr/bin/env python
"""bench.py -- headline benchmark of the chirpgp_b200 hot path (contract: see the task statement / DESIGN.md).

Workload (BASELINE.json configs[1]): per GPU B = 1000 synthetic toymodel chirps, T = 3141,
dt = 1e-3, chirp model d = 4, Gauss-Hermite order 3 (81 sigma points), sgp_filter + sgp_smoother, float64.
One "step" = one filter + smoother pass over the whole batch.  metric = smoothed chirp time-steps / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); independent chirps are sharded across ranks (weak scaling:
every rank owns B = 1000 chirps), no data-path collective.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU = 1000
T = 3141
DT = 1e-3
XI = 0.1
PARAMS = [0.1, 0.1, 0.1, 1., 1., 7.]
D = 4
N_SIGMA = 81
METRIC = 'smoothed chirp time-steps/sec (batch x T), GHF+GHS'
UNIT = 'steps/s'

# dram__bytes_read.sum + dram__bytes_write.sum of one gh_duo_filter_kernel launch (ncu --set full, profiles/r1_ncu_summary.txt)
NCU_TRAFFIC_BYTES = 1_401_431_000    # 26.7 MB read + 1374.8 MB written: 552.8 MB algorithmic + 904.6 MB smoother workspace

# algorithmic bytes / flops per chirp time-step (DESIGN.md "Roofline accounting")
BYTES_FILTER = 8 + 8 * (D + D * D + 1)          # ys in, mf + Pf + nell out                      = 176
BYTES_SMOOTHER = 2 * 8 * (D + D * D)            # mf, Pf in; ms, Ps out                          = 320
BYTES_WORKSPACE = 8 * (2 * D * D + D)           # [G | mp | Pp] record the filter leaves for the sweep  = 288 (not algorithmic)
FLOPS_FILTER = 9403                             # sgp_filter, d=4, n=81, counting convention v1
FLOPS_SMOOTHER = 13668                          # sgp_smoother


def flops_per_step(variant: str, d: int = 4, n: int = 81) -> int:
    """Counting convention v1 of SURVEY 8(d) for the chirp model (dense d x d algebra, FMA = 2 flops,
    every exp/log/sin/cos/sqrt/div = 1 flop)."""
    U = 7 * d * d + 7 * d + 9
    CH = d ** 3 / 3 + d * d
    c_pt = 21
    SP = CH + n * (2 * d * d + d) + n * c_pt + 2 * n * d + 3 * n * d * d + 4 * d * d
    if variant == 'sgp_filter':
        return int(round(SP + U))
    if variant == 'sgp_smoother':
        return int(round(SP + 3 * n * d * d + 2 * d * d + CH + 2 * d ** 3 + (2 * d * d + 2 * d) + 4 * d ** 3 + 2 * d * d))
    raise ValueError(variant)


def synthetic_inputs(rank: int):
    from chirpgp_b200 import toymodels
    _, ys, _ = toymodels.synthetic_batch(B_PER_GPU, T, DT, Xi=XI, seed=2 + 1000 * rank)
    return ys


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_sample(n_chirps: int, nthreads: int = 0):
    """Times the CPU restatement of the reference path (oracle/, OpenMP over chirps) on `n_chirps` chirps of the
    same workload (more than 1000: the 1000 synthetic chirps repeated); returns (steps/s, seconds, threads)."""
    from oracle import oracle as orc
    from chirpgp_b200.quadratures import SigmaPoints
    from chirpgp_b200 import toymodels
    orc.build()
    _, ys, _ = toymodels.synthetic_batch(min(n_chirps, B_PER_GPU), T, DT, Xi=XI, seed=2)
    if n_chirps > ys.shape[0]:
        ys = np.tile(ys, (-(-n_chirps // ys.shape[0]), 1))[:n_chirps]
    spec = orc.ChirpSpec(PARAMS[0], PARAMS[1], PARAMS[3], PARAMS[4])
    m0, P0, H = orc.chirp_m0_P0_H(PARAMS[2], PARAMS[3], PARAMS[4], PARAMS[5])
    sg = SigmaPoints.gauss_hermite(D, 3)
    # all the host threads this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    threads = nthreads or len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    f = orc.sgp_filter(spec, sg, H, XI, m0, P0, DT, ys, nthreads=threads)
    orc.sgp_smoother(spec, sg, f[0], f[1], DT, nthreads=threads)
    dt = time.perf_counter() - t0
    return n_chirps * T / dt, dt, threads


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path.  JAX cannot be installed in this image
    (no wheel, no network), so the arm runs the C restatement in oracle/ ('port') on all host threads, each step a
    bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cpu_sample(max(2 * cores, 16))                              # spin up the OpenMP team
    rate, _, _ = cpu_sample(max(4 * cores, 32))                 # calibrate, then ~8 s of CPU work per timed step
    n_chirps = int(min(8000, max(2 * cores, rate * 8. / T)))
    for _ in range(args.warmup):
        cpu_sample(max(cores, 4))
    vals, secs = [], []
    for _ in range(args.steps):
        v, s, threads = cpu_sample(n_chirps)
        vals.append(v); secs.append(s)
    value = statistics.mean(vals)
    sample = '%d chirps x %d steps per timed step (GHF+GHS, d=4, 81 points), C restatement, OpenMP' % (n_chirps, T)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * statistics.mean(secs), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'configs[1]: 1000 toymodel chirps x T=3141, GHF+GHS (gauss_hermite d=4 order 3)',
                   'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'JAX not installable here: CPU arm = oracle/ C restatement of the reference algorithm (proxy)',
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import chirpgp_b200 as cg
    from chirpgp_b200 import _native

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    ys_host = torch.as_tensor(synthetic_inputs(rank)).pin_memory()
    ys = ys_host.to(dev)
    _, _, m_and_cov, m0, P0, H = cg.build_chirp_model(np.array(PARAMS))
    sgps = cg.SigmaPoints.gauss_hermite(d=D, order=3)
    m0, P0, H = m0.to(dev), P0.to(dev), H.to(dev)
    mss_host = torch.empty((B_PER_GPU, T, D), dtype=torch.float64).pin_memory()
    Pss_host = torch.empty((B_PER_GPU, T, D, D), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # > 126 MB L2

    def step():
        f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys)
        return f, None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys)
        s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
        del f, s
    # ---- device-resident timing: K steps, L2 flushed between steps (flush outside the per-step events)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    marks = []
    for _ in range(args.steps):
        flush.fill_(1.)
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys)
        e1.record()
        s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
        e2.record()
        marks.append((e0, e1, e2))
        del f, s
    barrier()
    t_filter = [a.elapsed_time(b) for a, b, _ in marks]
    t_step = [a.elapsed_time(c) for a, _, c in marks]
    ms_step = statistics.mean(t_step)
    ms_filter = statistics.mean(t_filter)

    # ---- device-resident throughput with steps issued on two alternating streams (batch after batch): the filter of
    # step i+1 (latency-bound, leaves most issue slots idle) overlaps the smoother kernels of step i.  Reported as an
    # extra key; `value` stays the serial, L2-flushed figure.
    pstreams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    n_pipe = max(4, args.steps)

    def dev_step(i):
        with torch.cuda.stream(pstreams[i % 2]):
            f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys)
            cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)

    for i in range(2):
        dev_step(i)
    barrier()
    e0 = ev(); e0.record()
    for st in pstreams:
        st.wait_event(e0)
    for i in range(n_pipe):
        dev_step(i)
    e1 = ev()
    for st in pstreams:
        torch.cuda.current_stream(dev).wait_stream(st)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_pipe = e0.elapsed_time(e1) / n_pipe

    # ---- end-to-end through the public API: every step copies its inputs from pinned host memory, filters, smooths and
    # copies the smoothed means / covariances back to pinned host memory.  Steps are issued on two alternating CUDA
    # streams (what a user who processes batch after batch does), so the device->host copy of step i overlaps the
    # filter of step i+1; all copies of all steps are inside the timed region.
    n_e2e = max(4, min(args.steps, 10))
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    outs = [(mss_host, Pss_host),
            (torch.empty((B_PER_GPU, T, D), dtype=torch.float64).pin_memory(),
             torch.empty((B_PER_GPU, T, D, D), dtype=torch.float64).pin_memory())]

    def e2e_step(i):
        st = streams[i % 2]
        with torch.cuda.stream(st):
            ys_d = ys_host.to(dev, non_blocking=True)
            f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys_d)
            s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
            outs[i % 2][0].copy_(s[0], non_blocking=True)
            outs[i % 2][1].copy_(s[1], non_blocking=True)

    for i in range(2):
        e2e_step(i)
    barrier()
    flush.fill_(1.)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e0 = ev(); e0.record()
    for st in streams:
        st.wait_event(e0)
    for i in range(n_e2e):
        e2e_step(i)
    e1 = ev()
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_e2e = e0.elapsed_time(e1) / n_e2e
    ms_e2e_wall = (time.perf_counter() - t0) * 1e3 / n_e2e
    clocks = sampler.stop()              # sampled through the three timed regions (serial, two-stream, end-to-end)
    # single-step latency (one stream, no overlap) for reference; first pass warms the default stream's allocator pool
    for _ in range(2):
        flush.fill_(1.)
        torch.cuda.synchronize(dev)
        e0, e1 = ev(), ev()
        e0.record()
        ys_d = ys_host.to(dev, non_blocking=True)
        f = cg.sgp_filter(m_and_cov, sgps, H, XI, m0, P0, DT, ys_d)
        s = cg.sgp_smoother(m_and_cov, sgps, f[0], f[1], DT)
        mss_host.copy_(s[0], non_blocking=True)
        Pss_host.copy_(s[1], non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        ms_e2e_single = e0.elapsed_time(e1)
        del f, s, ys_d

    # ---- FP64 peak (DFMA-only kernel) on this GPU
    L = _native.lib()
    blocks, iters = 148 * 8, 20000
    out = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    best = 0.
    for i in range(4):
        e0, e1 = ev(), ev()
        e0.record()
        fl = L.cgp_bench_dfma(C.c_void_p(out.data_ptr()), blocks, iters, stream)
        e1.record()
        torch.cuda.synchronize(dev)
        if i > 0:
            best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp64_peak = best

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_step, ms_filter, ms_e2e, ms_pipe], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_filter, ms_e2e, ms_pipe = [float(x) for x in t.tolist()]
    n_steps_total = world * B_PER_GPU * T
    value = n_steps_total / (ms_step * 1e-3)
    e2e_value = n_steps_total / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get('hbm_gbs', 6650.))
        hbm_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback'
        filt_bytes = B_PER_GPU * T * BYTES_FILTER
        filt_flops = B_PER_GPU * T * flops_per_step('sgp_filter')
        ach_gbs = filt_bytes / (ms_filter * 1e-3) / 1e9
        ach_tf = filt_flops / (ms_filter * 1e-3) / 1e12
        step_flops = B_PER_GPU * T * (flops_per_step('sgp_filter') + flops_per_step('sgp_smoother'))
        # CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            cpu_sample(max(2 * cores, 16))                      # spin up the OpenMP team
            rate, _, _ = cpu_sample(max(4 * cores, 32))         # calibrate, then ~12 s of CPU work
            n = int(min(8000, max(2 * cores, rate * 12. / T)))
            v, sec, threads = cpu_sample(n)
            cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                   'sample': '%d chirps x %d steps, GHF+GHS, oracle/ C restatement with OpenMP (%.1f s)' % (n, T, sec)}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': 'configs[1]: %d toymodel chirps per GPU x T=%d, dt=1e-3, chirp model d=4, '
                                   'sgp_filter + sgp_smoother with gauss_hermite(d=4, order=3) (81 points)' % (B_PER_GPU, T),
                       'batch_per_gpu': B_PER_GPU, 'T': T, 'parallelism': 'chirps sharded x%d, no collective' % world,
                       'l2': 'flushed between timed steps (256 MiB write)'},
            'clocks': clocks,
            'value_two_streams': {'value': n_steps_total / (ms_pipe * 1e-3), 'unit': UNIT, 'ms_per_step': ms_pipe,
                                  'what': 'same passes, device-resident, issued on two alternating streams (no L2 flush)'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': B_PER_GPU * T * 8,
                    'd2h_bytes_per_step': B_PER_GPU * T * 8 * (D + D * D), 'steps': n_e2e,
                    'ms_per_step_wall_clock': ms_e2e_wall, 'ms_single_step_latency': ms_e2e_single,
                    'd2h_gb_per_s': B_PER_GPU * T * 8 * (D + D * D) / (ms_e2e * 1e-3) / 1e9,
                    'bound': 'PCIe device->host copy of the 503 MB result (kernels hidden behind the copy of the previous step)',
                    'what': 'pinned host ys -> cg.sgp_filter -> cg.sgp_smoother -> pinned host (mss, Pss); steps issued on '
                            'two alternating streams so the D2H copy of one step overlaps the filter of the next'},
            'gpu_launches': args.steps * 2,
            'kernels_per_step': ['gh_duo_filter_kernel (sgp_filter + smoother gains)', 'smoother_sweep_lane4_kernel (sgp_smoother)'],
            'roofline': {'bound': 'hbm', 'kernel': 'gh_duo_filter_kernel (sgp_filter + smoother gains, producer/consumer warp pair per chirp)',
                         'achieved': ach_gbs,
                         'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gbs / hbm_peak, 'traffic': NCU_TRAFFIC_BYTES,
                         'peak_source': hbm_src, 'kernel_ms': ms_filter,
                         'algorithmic_bytes_per_step': BYTES_FILTER, 'workspace_bytes_per_step': BYTES_WORKSPACE,
                         'note': 'latency-bound, not bandwidth-bound: see roofline_fp64 and DESIGN.md section 4'},
            'roofline_fp64': {'bound': 'fp64', 'kernel': 'gh_duo_filter_kernel', 'achieved': ach_tf, 'peak': fp64_peak,
                              'unit': 'TFLOP/s', 'frac': ach_tf / fp64_peak if fp64_peak else None,
                              'flops_per_step': flops_per_step('sgp_filter'),
                              'peak_source': 'DFMA-only kernel measured in this run',
                              'whole_step_tflops': step_flops / (ms_step * 1e-3) / 1e12},
            # neither of the two throughput rooflines binds at 1000 chirps per GPU: the kernel time is T x (cycles one
            # producer warp needs per step); floor = the pure dependency latency of one step
            'roofline_chain': {'bound': 'dependency-chain latency of one filter step', 'kernel': 'gh_duo_filter_kernel',
                               'achieved_cycles_per_step': ms_filter * 1e-3 * (clocks.get('sm_mhz') or 1965.) * 1e6 / T,
                               'floor_cycles_per_step': 840, 'scheduled_cycles_per_step': 1394,
                               'frac': 840. / (ms_filter * 1e-3 * (clocks.get('sm_mhz') or 1965.) * 1e6 / T),
                               'source': 'profiles/sass_dyn.py (static schedule), profiles/microbench/fp64_latency.cu'},
            'cpu_baseline': cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
