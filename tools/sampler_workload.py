"""Two EP iterations of a bench workload on a chosen number of sites (ncu / EPGPU_TRACE driver).

    python tools/sampler_workload.py cfg4 148 40      # workload, sites, sampling iterations per chain
"""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'ep-stan_b200')
import numpy as np, bench
import epstan.method as method
model, K, n_k, D, chains, siter = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
K = int(sys.argv[2]) if len(sys.argv) > 2 else K
siter = int(sys.argv[3]) if len(sys.argv) > 3 else 40
X, y, prior = bench.build_problem(model, K, n_k, D, 0, K)
m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior, chains=chains, iter=siter, df0=bench.default_df0(K))
res = m.run(2, verbose=False, seed=1, return_analytics=True)
print(res[0], res[2][0], m.n_leapfrog_total)
