"""Diagnostic: run a few EP iterations of a bench workload and dump what the sites rejected for a large
split-Rhat look like (per-chain means / sds of phi, cavity spectrum, data summary)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'ep-stan_b200'))
import bench
import epstan.method as method
from epstan import _lib

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model, K, n_k, D, chains, siter = bench.WORKLOADS[wl]
X, y, prior = bench.simulate_problem(model, K, n_k, D)
m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior,
                  chains=chains, iter=siter, df0=bench.default_df0(K), df_select='snr', rhat_max=2.0)
ctx = m._shard.ctx
info = m.run(niter, verbose=False, seed=1234)
print('info', info, 'rejected per iteration', m.history['rejected'], 'df', m.history['df'])
rh = np.array([w.last_mrhat for w in m.workers])
ms = np.array([w.last_msteps for w in m.workers])
nl = np.array([w.last_n_leapfrog for w in m.workers])
order = np.argsort(-np.nan_to_num(rh, nan=1e9))
print('rhat quantiles 50/90/99/100', np.nanpercentile(rh, [50, 90, 99, 100]))
print('worst sites', order[:8], rh[order[:8]], 'msteps', ms[order[:8]], 'nleap', nl[order[:8]], 'median nleap', np.median(nl))
n = chains * (siter - siter // 2)
per = n // chains
dr = ctx.get_draws(n)          # [K][d][n]
d = dr.shape[1]
out = {}
for k in list(order[:4]) + [int(np.argsort(rh)[K // 2])]:
    x = dr[k].reshape(d, chains, per)
    cm, cs = x.mean(axis=2), x.std(axis=2)
    spread = (cm.max(axis=1) - cm.min(axis=1)) / np.maximum(cs.mean(axis=1), 1e-12)
    top = np.argsort(-spread)[:5]
    yk = y[k * n_k:(k + 1) * n_k]
    Xk = X[k * n_k:(k + 1) * n_k]
    ev = np.linalg.eigvalsh(m._cavQ[:, :, k])
    print('--- site %d: rhat %.2f mstep %.4f nleap %d  ybar %.3f  X sd range %.2f..%.2f  cavity eig %.3g..%.3g' % (
        k, rh[k], ms[k], nl[k], yk.mean(), Xk.std(axis=0).min(), Xk.std(axis=0).max(), ev[0], ev[-1]))
    for i in top:
        print('   phi[%d]: chain means %s  chain sds %s  cavity mean %.3f sd %.3f' % (
            i, np.round(cm[i], 3), np.round(cs[i], 3), m._cavm[i, k], 1 / np.sqrt(m._cavQ[i, i, k])))
    out['draws_%d' % k] = dr[k]
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'diag_sites2.npz'), rh=rh, ms=ms, nl=nl, **out)
