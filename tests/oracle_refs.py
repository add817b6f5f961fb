"""Oracle NUTS / oracle EP reference results for the GPU sampler tests.

The fp64 NumPy oracle takes 40-90 s per case; its (deterministic, seeded) results are
cached in tests/golden/nuts_ref.npz so that the GPU box spends its time on the GPU.
`python tests/oracle_refs.py` regenerates the file; `summary()` / `ep_reference()` fall back
to computing when a key is missing, and tests/test_oracle_nuts.py re-derives one entry.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import density as dens        # noqa: E402
from oracle import ep_linalg as orc       # noqa: E402
from oracle import nuts                   # noqa: E402
import synth                              # noqa: E402

PATH = os.path.join(HERE, 'golden', 'nuts_ref.npz')

# (model, J, n, D) of test_gpu_sampler.py::test_sampler_vs_oracle_nuts (site seed 21)
SAMPLER_CASES = [('m1b', 1, 300, 4), ('m3b', 1, 400, 3), ('m4b', 2, 300, 3), ('m1b', 5, 250, 6),
                 ('m2b', 3, 300, 4), ('m5b', 1, 300, 3)]
EP_MODELS = ['m1b', 'm4b']
_cache = None


def _load():
    global _cache
    if _cache is None:
        _cache = dict(np.load(PATH)) if os.path.exists(PATH) else {}
    return _cache


def compute_summary(model, J, n, D):
    """8 chains x 2500 iterations of the oracle NUTS on the synthetic site; moments of phi."""
    site = synth.make_site(model, n, D, J, seed=21)
    td = synth.oracle_density(model, site)
    ref = nuts.sample(lambda q: tuple(v[0] for v in td.lp_grad(q[None])), td.p, chains=8,
                      n_iter=2500, n_warmup=500, seed=3)
    per_chain = [c[:, :td.d] for c in ref['per_chain']]
    draws = ref['draws'][:, :td.d]
    ess, mcse = nuts.ess_mcse(per_chain)
    return dict(mean=draws.mean(axis=0), var=draws.var(axis=0, ddof=1), ess=ess, mcse=mcse,
                stepsize=np.float64(ref['stepsize']), evals_per_draw=np.float64(ref['n_grad'] / (8 * 2500.0)))


def summary(model, J, n, D):
    key = 'nuts_%s_%d_%d_%d_' % (model, J, n, D)
    c = _load()
    if key + 'mean' in c:
        return {k: c[key + k] for k in ('mean', 'var', 'ess', 'mcse', 'stepsize', 'evals_per_draw')}
    return compute_summary(model, J, n, D)


def ep_problem(model, K, n_k, D, seed):
    """synthetic K-site problem of test_full_ep_vs_oracle_ep"""
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((K * n_k, D)) * 0.8
    beta = rng.standard_normal(D) * 0.7
    alpha = 0.6 * rng.standard_normal(K)
    bk = np.repeat(beta[None], K, axis=0) + (0.0 if model == 'm1b' else 0.3 * rng.standard_normal((K, D)))
    k_ind = np.repeat(np.arange(K), n_k)
    f = alpha[k_ind] + np.einsum('nd,nd->n', X, bk[k_ind])
    y = (rng.uniform(size=K * n_k) < 1 / (1 + np.exp(-f))).astype(np.int64)
    d = dens.dphi(model, D)
    prior = {'Q': np.eye(d) / 1.5 ** 2, 'r': np.zeros(d)}
    return X, y, prior, d


def compute_ep_reference(model):
    """oracle EP (oracle NUTS per site) of test_full_ep_vs_oracle_ep: final mean and covariance."""
    K, n_k, D, C, siter, niter = 4, 150, 3, 4, 400, 6
    X, y, prior, d = ep_problem(model, K, n_k, D, seed=9)
    st = orc.EPState(prior['Q'], prior['r'], K)
    dens_k = [dens.TiltedDensity(model, X[k * n_k:(k + 1) * n_k], y[k * n_k:(k + 1) * n_k],
                                 np.zeros(d), np.eye(d)) for k in range(K)]

    def draw_fn(it, k, cav_m, cav_P):
        td = dens_k[k]
        td.mu, td.Omega = cav_m, cav_P
        res = nuts.sample(lambda q: tuple(v[0] for v in td.lp_grad(q[None])), td.p, chains=C,
                          n_iter=siter, seed=1000 * it + k)
        return res['draws'][:, :d]

    oinfo, oms, oSs = orc.run_ep(st, draw_fn, niter, lambda i: 0.6)
    return dict(info=np.int64(oinfo), m=oms[-1], S=oSs[-1])


def ep_reference(model):
    key = 'ep_%s_' % model
    c = _load()
    if key + 'm' in c:
        return {k: c[key + k] for k in ('info', 'm', 'S')}
    return compute_ep_reference(model)


def main():
    out = {}
    for case in SAMPLER_CASES:
        s = compute_summary(*case)
        for k, v in s.items():
            out['nuts_%s_%d_%d_%d_' % case + k] = v
        print(case, 'stepsize %.3f' % s['stepsize'], 'evals/draw %.1f' % s['evals_per_draw'])
    for model in EP_MODELS:
        r = compute_ep_reference(model)
        for k, v in r.items():
            out['ep_%s_' % model + k] = v
        print('ep', model, r['info'], np.round(r['m'], 3))
    np.savez_compressed(PATH, **out)
    print('wrote', PATH, len(out), 'arrays')


if __name__ == '__main__':
    main()
