"""Oracle references for posterior parity (check (c)) on BASELINE shapes.

    python tests/oracle_refs_ep.py cfg3            # oracle EP: ~20 min on 8 cores
    python tests/oracle_refs_ep.py cfg3 8 target   # + the full-data posterior by the oracle NUTS (128 000 rows: hours)
    python tests/oracle_refs_ep.py cfg4s           # config-4 subset (first 16 sites, 5 EP iterations): ~1 h on 8 cores

For a workload this caches in tests/golden/ep_ref_<tag>.npz
  * the oracle EP run (oracle NUTS per site: oracle/nuts.py on oracle/density.py, oracle moment matching and
    updates: oracle/ep_linalg.py) on the SAME data the GPU test uses (bench.simulate_problem, seed 100), with
    the same seed-independent settings (chains, iterations, damping rule), and
  * optionally (argument `target`) the full-data posterior of phi by the oracle NUTS ("target", as
    experiment/fit.py --run_target does it with one multi-group Stan program).  It takes hours on the
    128 000-row problem; the GPU test uses the GPU full-data sampler for that comparison instead, which
    tests/test_gpu_experiment.py::test_fit_results_against_oracle_posterior pins to the oracle on a problem
    the oracle finishes in seconds.
Both are fp64 CPU runs; the GPU results must agree with them within the KL tolerances stated in
tests/test_gpu_parity_ep.py.  PyStan itself is not installable here: parity of the sampler is
pinned on this restatement, the moment/update path on the reference itself.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import density as dens        # noqa: E402
from oracle import ep_linalg as orc       # noqa: E402
from oracle import nuts                   # noqa: E402

# tag: (model, K_total simulated, K used, n_k, D, chains, siter, EP iterations)
CASES = {
    'cfg3': ('m1b', 64, 64, 2000, 19, 8, 200, 12),
    'cfg4s': ('m3b', 1024, 16, 5000, 49, 4, 200, 5),
}


def problem(tag):
    import bench
    model, Ktot, K, n_k, D, C, siter, niter = CASES[tag]
    X, y, prior = bench.simulate_problem(model, Ktot, n_k, D)
    return X[:K * n_k], y[:K * n_k], prior


_G = {}


def _site_job(args):
    tag, it, k, cav_m, cav_P = args
    model, Ktot, K, n_k, D, C, siter, niter = CASES[tag]
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    X, y = _G['X'], _G['y']
    td = dens.TiltedDensity(model, X[k * n_k:(k + 1) * n_k], y[k * n_k:(k + 1) * n_k], cav_m, cav_P)
    res = nuts.sample(lambda q: tuple(v[0] for v in td.lp_grad(q[None])), td.p, chains=C, n_iter=siter,
                      seed=100000 * it + k)
    return res['draws'][:, :td.d], res['n_grad'], res['stepsize']


def oracle_ep(tag, procs):
    model, Ktot, K, n_k, D, C, siter, niter = CASES[tag]
    X, y, prior = problem(tag)
    _G['X'], _G['y'] = X, y
    d = dens.dphi(model, D)
    st = orc.EPState(np.asarray(prior['Q'], dtype=np.float64), np.asarray(prior['r'], dtype=np.float64), K)
    df0 = orc.default_df0(K)
    floor = min(1.0 / K, 0.2)
    ms, Ss, dfs = [], [], []
    with mp.get_context('fork').Pool(procs) as pool:
        for it in range(1, niter + 1):
            t0 = time.time()
            st.iter = it
            jobs = [(tag, it, k, st.cav_m[:, k].copy(), st.cav_Q[:, :, k].copy()) for k in range(K)]
            out = pool.map(_site_job, jobs, chunksize=1)
            oks = np.zeros(K, dtype=bool)
            for k, (draws, ng, eps) in enumerate(out):
                oks[k], st.dQi[:, :, k], st.dri[:, k] = orc.tilted_moments(draws, st.Q, st.r, 'sample')
            T2, S2, n_ok = orc.snr_stats(st, oks)
            df = orc.snr_damping(T2, S2, n_ok, d, df0(it), floor, 3.0)
            info, dfu, m, S = st.update(lambda i: df)
            assert info == 0, info
            ms.append(m)
            Ss.append(S)
            dfs.append(dfu)
            print('[%s] EP iter %d: df %.4f, grad evals %.3g, mean eps %.3f, %.0f s' % (
                tag, it, dfu, sum(o[1] for o in out), np.mean([o[2] for o in out]), time.time() - t0), flush=True)
    return dict(ep_m=np.array(ms), ep_S=np.array(Ss), ep_df=np.array(dfs))


def _target_chain(args):
    tag, c = args
    model, Ktot, K, n_k, D, C, siter, niter = CASES[tag]
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    X, y, prior = _G['X'], _G['y'], _G['prior']
    Q0 = np.asarray(prior['Q'], dtype=np.float64)
    m0 = np.linalg.solve(Q0, np.asarray(prior['r'], dtype=np.float64))
    j_ind = np.repeat(np.arange(K), n_k)
    td = dens.TiltedDensity(model, X, y, m0, Q0, j_ind=j_ind, J=K)
    res = nuts.sample(lambda q: tuple(v[0] for v in td.lp_grad(q[None])), td.p, chains=1, n_iter=1000,
                      n_warmup=500, seed=777 + c)
    return res['draws'][:, :td.d]


def oracle_target(tag, procs):
    X, y, prior = problem(tag)
    _G['X'], _G['y'], _G['prior'] = X, y, prior
    with mp.get_context('fork').Pool(min(procs, 8)) as pool:
        chains = pool.map(_target_chain, [(tag, c) for c in range(8)])
    draws = np.concatenate(chains, axis=0)
    means = np.array([c.mean(axis=0) for c in chains])
    print('[%s] target: %d draws, between-chain sd of the mean / posterior sd: max %.3f' % (
        tag, draws.shape[0], np.max(means.std(axis=0) / draws.std(axis=0))), flush=True)
    return dict(tgt_m=draws.mean(axis=0), tgt_S=np.cov(draws.T), tgt_chain_means=means)


def main():
    tag = sys.argv[1]
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    out = {}
    if len(sys.argv) > 3 and sys.argv[3] == 'target':
        out.update(oracle_target(tag, procs))
    out.update(oracle_ep(tag, procs))
    path = os.path.join(HERE, 'golden', 'ep_ref_%s.npz' % tag)
    np.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    main()
