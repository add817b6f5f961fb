"""Statistical pin of oracle/nuts.py on targets with known moments."""
import numpy as np

from oracle import nuts


def test_nuts_gaussian_moments():
    rng = np.random.RandomState(5)
    dim = 6
    A = rng.standard_normal((dim, dim))
    S = A @ A.T / dim + 0.3 * np.eye(dim)
    P = np.linalg.inv(S)
    m = rng.standard_normal(dim)

    def lp_grad(q):
        dq = q - m
        return -0.5 * dq @ P @ dq, -(P @ dq)

    res = nuts.sample(lp_grad, dim, chains=4, n_iter=1200, n_warmup=300, seed=1)
    x = res['draws']
    ess, mcse = nuts.ess_mcse(res['per_chain'])
    assert np.all(np.abs(x.mean(axis=0) - m) < 5 * mcse + 1e-3)
    assert np.max(np.abs(np.cov(x.T) - S)) < 0.25 * np.max(np.diag(S))
    assert np.all(nuts.split_rhat(res['per_chain']) < 1.05)
    assert 0.05 < res['stepsize'] < 2.0 and res['n_divergent'] == 0


def test_adaptation_schedule():
    w = nuts.VarWindows(100, 2)
    assert (w.init_buffer, w.term_buffer, w.base_window) == (15, 10, 75)
    ends = [i for i in range(100) if w.learn(np.array([float(i), 0.0])) is not None]
    assert ends == [89]
    w = nuts.VarWindows(1000, 1)
    ends = [i for i in range(1000) if w.learn(np.array([float(i % 7)])) is not None]
    assert ends == [99, 149, 249, 449, 949]
    w = nuts.VarWindows(10, 1)
    assert all(w.learn(np.zeros(1)) is None for _ in range(10))


def test_cached_oracle_references_are_current():
    """tests/golden/nuts_ref.npz (oracle NUTS / oracle EP results used by the GPU sampler tests) holds
    every case, and re-running the oracle reproduces the stored numbers (same seeds, same code)."""
    import oracle_refs
    cache = oracle_refs._load()
    for case in oracle_refs.SAMPLER_CASES:
        for k in ('mean', 'var', 'ess', 'mcse', 'stepsize', 'evals_per_draw'):
            assert 'nuts_%s_%d_%d_%d_' % case + k in cache
    for model in oracle_refs.EP_MODELS:
        assert 'ep_%s_m' % model in cache and int(cache['ep_%s_info' % model]) == 0
    case = oracle_refs.SAMPLER_CASES[0]
    fresh = oracle_refs.compute_summary(*case)
    stored = oracle_refs.summary(*case)
    # same seeds: normally identical; a different BLAS build may change low-order bits and with them the
    # trajectories, so the check is statistical
    assert np.all(np.abs(fresh['mean'] - stored['mean']) < 4.0 * np.sqrt(2.0) * stored['mcse'])
    assert np.all(np.abs(fresh['var'] / stored['var'] - 1.0) < 0.3)
    assert abs(fresh['stepsize'] / stored['stepsize'] - 1.0) < 0.2
    assert abs(fresh['evals_per_draw'] / stored['evals_per_draw'] - 1.0) < 0.3
