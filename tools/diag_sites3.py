"""Diagnostic: what limits the step size of the slowest sites (adapted metric / step size per chain)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'ep-stan_b200'))
import bench
import epstan.method as method
model, K, n_k, D, chains, siter = bench.WORKLOADS['cfg4']
niter = int(sys.argv[1]) if len(sys.argv) > 1 else 6
X, y, prior = bench.simulate_problem(model, K, n_k, D)
m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior,
                  chains=chains, iter=siter, df0=bench.default_df0(K), df_select='snr', rhat_max=2.0)
ctx = m._shard.ctx
m.run(niter, verbose=False, seed=1234)
nl = np.array([w.last_n_leapfrog for w in m.workers]); rh = np.array([w.last_mrhat for w in m.workers]); ms = np.array([w.last_msteps for w in m.workers])
order = np.argsort(-nl)
print('nleap quantiles 10/50/90/99/100', np.percentile(nl, [10, 50, 90, 99, 100]))
print('sites with nleap > 600k:', np.sum(nl > 600e3), ' > 400k:', np.sum(nl > 400e3), ' > 200k:', np.sum(nl > 200e3))
n = chains * (siter - siter // 2); per = n // chains
dr = ctx.get_draws(n); d = dr.shape[1]; P = 128
for k in list(order[:4]) + [int(order[K // 2])]:
    minv, eps = ctx.get_adapt(k, chains, P)
    x = dr[k].reshape(d, chains, per)
    sd_draw = x.std(axis=2).mean(axis=1)
    msd = np.sqrt(minv[:, :100])
    ratio_phi = sd_draw / msd[:, :d].mean(axis=0)
    Xk = X[k * n_k:(k + 1) * n_k]; yk = y[k * n_k:(k + 1) * n_k]
    print('--- site %d nleap %d rhat %.3f mstep %.4f ybar %.3f' % (k, nl[k], rh[k], ms[k], yk.mean()))
    print('   eps per chain', eps)
    print('   metric sd phi: min %.2e med %.2e max %.2e | latents (eta, etb): min %.2e med %.2e max %.2e' % (
        msd[:, :d].min(), np.median(msd[:, :d]), msd[:, :d].max(), msd[:, d:100].min(), np.median(msd[:, d:100]), msd[:, d:100].max()))
    print('   draw sd / metric sd over phi dims: min %.3f med %.3f max %.3f; draw sd phi min %.2e med %.2e' % (
        ratio_phi.min(), np.median(ratio_phi), ratio_phi.max(), sd_draw.min(), np.median(sd_draw)))
    print('   smallest metric sd dims (chain 0):', np.argsort(msd[0])[:6], np.sort(msd[0])[:6])
