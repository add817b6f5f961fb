"""Regenerates, from the fixed seeds, the INPUTS that oracle/make_golden.py fed
to the reference (only the reference's outputs are stored in tests/golden)."""

import numpy as np
from scipy import linalg
from scipy.stats import multivariate_normal

from oracle import fakes

MAX_UINT = 2 ** 31 - 1   # pystan.constants.MAX_UINT (reference method.py:40)

LINALG_CASES = (('d6', 11, 6), ('d20', 12, 20), ('d50', 13, 50))
WORKER_CASES = (('d6', 21, 6, 4, 50), ('d20', 22, 20, 8, 200), ('d50', 23, 50, 8, 200))
CV_CASES = (('n60d4', 41, 60, 4), ('n800d20', 42, 800, 20))


def linalg_case(seed, d):
    rng = np.random.RandomState(seed)
    S = fakes.random_spd(rng, d)
    m = rng.standard_normal(d)
    U = linalg.cholesky(S, lower=False)
    Ug = U + np.tril(rng.standard_normal((d, d)), -1)
    P = fakes.random_spd(rng, d, scale=3.0)
    return dict(S=S, m=m, Ug=Ug, P=P, n=10 * d)


def worker_case(seed, d, C, it):
    rng = np.random.RandomState(seed)
    K = 3
    Qs, rs = fakes.gaussian_site_factors(seed + 100, K, d)
    Q = fakes.random_spd(rng, d, scale=4.0) + Qs.sum(axis=2) * 0.5
    r = rng.standard_normal(d)
    n = C * (it - it // 2)
    return dict(K=K, Qs=Qs, rs=rs, Q=Q, r=r, n=n,
                Qi=[0.5 * Qs[:, :, k] for k in range(K)],
                ri=[0.5 * rs[:, k] for k in range(K)],
                seeds=[seed * 7 + k for k in range(K)])


def stan_seed(worker_seed):
    """Worker.tilted: method.py:342-346."""
    return int(np.random.RandomState(worker_seed).randint(0, MAX_UINT))


def run_seeds(seed, niter, K):
    """Master.run: method.py:956-960."""
    return np.random.RandomState(seed).randint(0, MAX_UINT, size=(niter, K))


def cv_case(seed, n, d):
    rng = np.random.RandomState(seed)
    S1 = fakes.random_spd(rng, d)
    m1 = rng.standard_normal(d)
    S2 = S1 + 0.1 * fakes.random_spd(rng, d)
    m2 = m1 + 0.1 * rng.standard_normal(d)
    samp = m1 + rng.standard_normal((n, d)) @ linalg.cholesky(S1, lower=False)
    lp = multivariate_normal(mean=m1, cov=S1).logpdf(samp)
    return dict(S1=S1, m1=m1, S2=S2, m2=m2, samp=samp, lp=lp, m3=m1 + 5.0)


# Master.run scenarios of make_golden.gen_master: tag -> description
def master_scenarios():
    from oracle.ep_linalg import default_df0
    rng = np.random.RandomState(32)
    priorB = {'Q': fakes.random_spd(rng, 8), 'r': rng.standard_normal(8)}
    return {
        'runA': dict(K=5, d=6, fseed=31, chains=4, iter=100, niter=8, seed=5,
                     df0=None, prec_estim='sample', skip=0, prior=None),
        'runB': dict(K=6, d=8, fseed=33, chains=4, iter=120, niter=10, seed=6,
                     df0=default_df0(6), prec_estim='olse', skip=2, prior=priorB),
        'runC': dict(K=8, d=10, fseed=34, chains=2, iter=40, niter=6, seed=7,
                     df0=1.0, prec_estim='sample', skip=0, prior=None),
        'runC2': dict(resume='runC', niter=3, seed=8),
        'runD': dict(K=4, d=5, fseed=35, chains=4, iter=100, niter=2, seed=9,
                     df0=0.5, prec_estim='sample', skip=0, prior=None, improper=True,
                     constant=(0, 1)),
        'runE': dict(K=4, d=5, fseed=36, chains=4, iter=100, niter=2, seed=10,
                     df0=1.0, prec_estim='sample', skip=0, prior=None, inflate=1e4),
        'runF': dict(K=3, d=4, fseed=37, chains=4, iter=60, niter=2, seed=11,
                     df0=None, prec_estim='sample', skip=0, prior=None, constant=True),
    }
