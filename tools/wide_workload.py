"""Config-5-shaped sampler workload (wide tensor-core pass vs the fp32 SIMT pass), sampler only.

    python tools/wide_workload.py [sites=148] [n_k=200000] [D=199] [chains=32] [iter=6] [depth=6] [tc|simt|both]

Prints, per path: kernel seconds, gradient evaluations (per chain), evaluations of a site's chain batch
(= streams of the site's design matrix), algorithmic HBM GB/s (n_k (2 D + 4) bytes per stream)
and TFLOP/s (4 n_k D per chain evaluation).  The same cavities / seeds for both paths; the draws of the
two paths are compared statistically (same distribution, different rounding).
"""
import os
import sys
import time

sys.path.insert(0, '.')
sys.path.insert(0, 'ep-stan_b200')
import numpy as np
import bench
from epstan import _lib

arg = sys.argv[1:]
K = int(arg[0]) if len(arg) > 0 else 148
n_k = int(arg[1]) if len(arg) > 1 else 200000
D = int(arg[2]) if len(arg) > 2 else 199
C = int(arg[3]) if len(arg) > 3 else 32
it = int(arg[4]) if len(arg) > 4 else 6
depth = int(arg[5]) if len(arg) > 5 else 6
which = arg[6] if len(arg) > 6 else 'tc'
model = 'm1b'
d = D + 1

t0 = time.perf_counter()
X, y, prior = bench.build_problem(model, K, n_k, D, 0, K, 'synth')
print('data %.1f s (%.1f GB fp64 on the host)' % (time.perf_counter() - t0, X.nbytes / 1e9), flush=True)
ctx = _lib.Context(0)
ctx.init_state(K, d)
t0 = time.perf_counter()
ctx.upload_sites(_lib.MODEL_IDS[model], D, np.arange(K + 1, dtype=np.int64) * n_k, X, y, None, None)
print('upload %.1f s' % (time.perf_counter() - t0), flush=True)
del X
# cavity = the prior (first EP iteration)
ctx.upload(_lib.CAVQ, np.asfortranarray(np.repeat(prior['Q'][:, :, None], K, axis=2)))
ctx.upload(_lib.CAVM, np.asfortranarray(np.zeros((d, K))))
seeds = np.arange(1, K + 1) * 7919
res = {}
for path in (['tc', 'simt'] if which == 'both' else [which]):
    ctx.set_option('use_tc', 1 if path == 'tc' else 0)
    for rep in range(2):                       # first call: allocation + cold caches
        msteps, mrhat, nleap, secs = ctx.tilted_sample(seeds, C, it, None, max_treedepth=depth)
    n = C * (it - it // 2)
    dr = ctx.get_draws(n)
    streams = nleap.sum() / C                  # lower bound: chains of a site advance in lock-step
    gbs = streams * n_k * (2.0 * D + 4.0) / secs / 1e9
    tf = nleap.sum() * 4.0 * n_k * D / secs / 1e12
    print('%s: %.3f s, %d leapfrogs, %.0f streams of X_k (%.2f ms each with %d sites concurrent), %.0f GB/s algorithmic, '
          '%.1f TFLOP/s, mean step %.4g, finite %s'
          % (path, secs, nleap.sum(), streams, secs / (streams / K) * 1e3 * min(K, 148) / K, min(K, 148), gbs, tf,
             float(np.mean(msteps)), bool(np.all(np.isfinite(dr)))), flush=True)
    res[path] = dr
if len(res) == 2:
    a, b = res['tc'], res['simt']
    sd = np.sqrt(0.5 * (a.var(axis=2) + b.var(axis=2)))
    z = np.abs(a.mean(axis=2) - b.mean(axis=2)) / np.maximum(sd, 1e-12)
    print('tc vs simt: max |mean difference| / sd over sites x parameters = %.3f (n = %d draws)' % (z.max(), a.shape[2]))
ctx.close()
