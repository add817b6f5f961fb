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
                 ('m2b', 3, 300, 4), ('m5b', 1, 300, 3), ('m1b', 1, 600, 70),
                 ('m1b', 1, 2000, 19), ('m3b', 1, 5000, 49)]       # BASELINE site shapes: 4 min / 55 min of CPU
#                  (tools: the chains of one case can run in parallel processes, RandomState([seed, chain]))
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


# ---- second cache: oracle posteriors used by the mix_pred / fit.py result tests -----------------------
PATH2 = os.path.join(HERE, 'golden', 'nuts_ref2.npz')
_cache2 = None
PARAM_STATS_CASES = [('m1b', 1, 300, 4), ('m4b', 3, 240, 3), ('m2b', 2, 300, 3)]


def _cached2(key, fn):
    """dict of arrays under `key`: from tests/golden/nuts_ref2.npz, else computed (slow, on the calling host)"""
    global _cache2
    if _cache2 is None:
        _cache2 = dict(np.load(PATH2)) if os.path.exists(PATH2) else {}
    pre = key + '/'
    hit = {k[len(pre):]: v for k, v in _cache2.items() if k.startswith(pre)}
    return hit if hit else fn()


def _transformed(model, q, d, J, D):
    """alpha, beta of the Stan programs from draws q = [phi | eta | etb] (m1b.stan:30-36 ...), as one matrix"""
    four = model in ('m4b', 'm5b')
    ia, ib = (1, 2 + D) if four else (0, 1)
    alpha = q[:, d:d + J] * np.exp(q[:, [ia]]) + (q[:, [0]] if four else 0.0)
    if model == 'm1b':
        return alpha
    if model == 'm2b':
        return np.concatenate([alpha, q[:, d + J:d + J + D] * np.exp(q[:, [ib]])], axis=1)
    etb = q[:, d + J:].reshape(-1, J, D)
    beta = etb * np.exp(q[:, None, ib:ib + D]) + (q[:, None, 2:2 + D] if four else 0.0)
    return np.concatenate([alpha, beta.reshape(len(q), -1)], axis=1)


def compute_param_stats_ref(model, J, n, D):
    site = synth.make_site(model, n, D, J, seed=23)
    td = dens.TiltedDensity(model, site['X'], site['y'], site['mu'], site['Omega'], j_ind=site['j_ind'], J=J)
    q = nuts.sample(lambda v: tuple(w[0] for w in td.lp_grad(v[None])), td.p, chains=8, n_iter=1500, seed=5)['draws']
    T = _transformed(model, q, site['d'], J, D)
    return dict(mean=T.mean(axis=0), sd=T.std(axis=0))


def param_stats_ref(model, J, n, D):
    return _cached2('pstats_%s_%d_%d_%d' % (model, J, n, D), lambda: compute_param_stats_ref(model, J, n, D))


def _fit_problem(model, J, D, npg, seed_data=100):
    import importlib
    exp = os.path.join(os.path.dirname(HERE), 'ep-stan_b200', 'experiment')
    if exp not in sys.path:
        sys.path.insert(0, exp)
    mdl = importlib.import_module('models.' + model).model(J, D, npg)
    data = mdl.simulate_data(Sigma_x='rand', rng=seed_data)
    _, _, Q0, r0 = mdl.get_prior()
    return data, Q0, r0


def compute_fit_posterior(model, J, D, npg, chains=8, n_iter=1500, seed=3):
    """full-data posterior of the fit.py problem (simulate_data seed 100) by the oracle NUTS:
    moments of phi and of alpha"""
    data, Q0, r0 = _fit_problem(model, J, D, npg)
    td = dens.TiltedDensity(model, data.X, data.y, np.linalg.solve(Q0, r0), Q0, j_ind=data.j_ind, J=J)
    q = nuts.sample(lambda v: tuple(w[0] for w in td.lp_grad(v[None])), td.p, chains=chains, n_iter=n_iter, seed=seed)['draws']
    d = td.d
    four = model in ('m4b', 'm5b')
    alpha = q[:, d:d + J] * np.exp(q[:, [1 if four else 0]]) + (q[:, [0]] if four else 0.0)
    return dict(phi_m=q[:, :d].mean(axis=0), phi_S=np.cov(q[:, :d].T), alpha_m=alpha.mean(axis=0), alpha_sd=alpha.std(axis=0))


def fit_posterior(model, J, D, npg):
    return _cached2('fitpost_%s_%d_%d_%d' % (model, J, D, npg), lambda: compute_fit_posterior(model, J, D, npg))


def main2():
    out = {}
    for case in PARAM_STATS_CASES:
        for k, v in compute_param_stats_ref(*case).items():
            out['pstats_%s_%d_%d_%d/' % case + k] = v
        print('pstats', case)
    for case in (('m1b', 6, 2, 40), ('m1b', 8, 3, 40), ('m4b', 8, 3, 40)):
        for k, v in compute_fit_posterior(*case).items():
            out['fitpost_%s_%d_%d_%d/' % case + k] = v
        print('fitpost', case)
    np.savez_compressed(PATH2, **out)
    print('wrote', PATH2, len(out), 'arrays')


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
    if len(sys.argv) > 1 and sys.argv[1] == '2':
        main2()
    else:
        main()
