"""Diagnostic: first-iteration tilted sampling of every site of a bench workload; dumps the per-site
analytics and the draws of the worst sites (gpurun_out/diag_sites.npz)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'ep-stan_b200'))
import bench
import epstan.method as method
from epstan import _lib

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
K = int(sys.argv[2]) if len(sys.argv) > 2 else None
model, K0, n_k, D, chains, siter = bench.WORKLOADS[wl]
K = K or K0
X, y, prior = bench.simulate_problem(model, K, n_k, D)
m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior,
                  chains=chains, iter=siter, df0=bench.default_df0(K), df_select='snr')
ctx = m._shard.ctx
seeds = np.random.RandomState(1234).randint(0, 2 ** 31 - 1, size=K)
msteps, mrhat, nleap, secs = ctx.tilted_sample(seeds, chains, siter, None, 0)
n = chains * (siter - siter // 2)
oks, n_ok = ctx.moments(n, 'sample')
order = np.argsort(-np.nan_to_num(mrhat, nan=1e9))
print('secs', secs, 'n_ok', n_ok, 'rhat quantiles', np.nanpercentile(mrhat, [50, 90, 99, 100]))
print('worst sites', order[:10], mrhat[order[:10]], msteps[order[:10]], nleap[order[:10]])
dr = ctx.get_draws(n)
out = dict(mrhat=mrhat, msteps=msteps, nleap=nleap, worst=order[:6], seeds=seeds)
for k in order[:6]:
    out['draws_%d' % k] = dr[k]
    out['X_%d' % k] = X[k * n_k:(k + 1) * n_k]
    out['y_%d' % k] = y[k * n_k:(k + 1) * n_k]
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'diag_sites.npz'), **out)
