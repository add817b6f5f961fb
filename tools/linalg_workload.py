"""Moment-matching / update / cavity kernels on synthetic state of a given shape
(config 2/4/5 of BASELINE.json): used under ncu and, with `time`, to report the
achieved HBM bandwidth of each kernel with CUDA events.

    python tools/linalg_workload.py K d n [time]
"""
import os
import sys
import json

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'ep-stan_b200'))
import torch
from epstan import _lib

K, d, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
timing = len(sys.argv) > 4
torch.cuda.set_device(0)
ctx = _lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
ctx.init_state(K, d)
rng = np.random.RandomState(0)
A = rng.standard_normal((d, d))
Q0 = A @ A.T / d + np.eye(d)
ctx.upload(_lib.Q0, Q0)
ctx.upload(_lib.R0, rng.standard_normal(d))
Qi = np.empty((d, d, K), order='F')
for k in range(K):
    B = rng.standard_normal((d, d))
    Qi[:, :, k] = 0.02 * (B @ B.T / d + 0.3 * np.eye(d))
ctx.upload(_lib.QI, Qi)
ctx.upload(_lib.Q, Q0 + Qi.sum(axis=2))
draws = rng.standard_normal((K, d, n)) * 0.3 + rng.standard_normal((K, d, 1))
ctx.set_draws(draws, n)
lib, h = ctx._lib, ctx._h


def run_all():
    ctx.moments(n, 'sample')
    ctx.moments(n, 'olse')
    ctx.update_partial(0.5 / K)
    ctx.update_finish()
    ctx.cavity(proposal=True)


run_all()
if timing:
    import ctypes as C
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0}

    def timed(fn, reps=20):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        ts = []
        for _ in range(reps):
            flush.zero_()                      # evict L2 between repetitions
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)) * 1e-3

    nullp = C.POINTER(C.c_int32)()
    nulli = C.POINTER(C.c_int)()
    # launch-only entry points (flags are read back by the wrappers; time the async launches via the C ABI)
    res = {}
    res['moments_sample'] = (timed(lambda: ctx.moments(n, 'sample')), 8.0 * K * (n * d + d * d + d))
    res['moments_olse'] = (timed(lambda: ctx.moments(n, 'olse')), 8.0 * K * (n * d + 2 * d * d + d))
    res['update_partial'] = (timed(lambda: ctx.update_partial(0.5 / K)), 8.0 * 3 * K * (d * d + d))
    res['cavity'] = (timed(lambda: ctx.cavity(proposal=True)), 8.0 * K * (2 * d * d + 2 * d + d * d + d))
    print('shape K=%d d=%d n=%d  (HBM peak %.0f GB/s measured)' % (K, d, n, peaks['hbm_gbs']))
    # algorithmic fp64 flop per site: moments 2 n d^2 + 7/3 d^3, cavity 1/3 d^3 + 2 d^2 (x2: FMA = 2 flop)
    flops = {'moments_sample': K * (2.0 * n * d * d + 7.0 / 3.0 * d ** 3), 'moments_olse': K * (2.0 * n * d * d + 7.0 / 3.0 * d ** 3),
             'cavity': K * (2.0 / 3.0 * d ** 3 + 4.0 * d * d), 'update_partial': 2.0 * K * (d * d + d)}
    for name, (t, byt) in res.items():
        print('  %-16s %9.1f us   algorithmic %8.2f MB   %7.1f GB/s   %5.1f %% of HBM peak   %6.2f fp64 TFLOP/s'
              % (name, t * 1e6, byt / 1e6, byt / t / 1e9, 100 * byt / t / 1e9 / peaks['hbm_gbs'], flops[name] / t / 1e12))
