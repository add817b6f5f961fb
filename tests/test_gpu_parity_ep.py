"""Posterior parity, north_star check (c), on BASELINE shapes (VERDICT r1 "next" item 2).

The GPU EP run (batched NUTS on the tcgen05 pass + fp64 moment matching / updates) against
  (i)  the oracle EP run on the SAME data with the same seed-independent settings and the same damping factors
       (oracle NUTS per site, oracle moment matching and updates; tests/oracle_refs_ep.py, cached in
       tests/golden/ep_ref_<tag>.npz): config 3 in full (K = 64, 12 iterations), the first 16 sites of config 4
       (5 iterations; there the statement is relative to the seed-to-seed spread of the GPU run itself, see the test);
  (ii) opt-in (EPGPU_SLOW_TESTS=1, 7 minutes on one SM): the full-data posterior of phi ("target", what fit.py
       --run_target does): all groups in one multi-group site whose cavity is the prior, sampled by the same GPU NUTS
       (8 x 500 draws).  That path is pinned to the fp64 oracle NUTS by
       tests/test_gpu_experiment.py::test_fit_results_against_oracle_posterior on a problem the oracle finishes in
       seconds (the oracle needs hours on the 128 000 rows of config 3).
Tolerances are KL divergences between the Gaussian approximations, stated per assertion: both EP runs carry
Monte Carlo noise from C x 100 draws per site and iteration.
"""
import os

import numpy as np
import pytest

from oracle import ep_linalg as orc
import oracle_refs_ep as refs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _ref(tag):
    path = os.path.join(GOLD, 'ep_ref_%s.npz' % tag)
    if not os.path.exists(path):
        pytest.skip('no cached oracle reference %s (python tests/oracle_refs_ep.py %s)' % (path, tag))
    return np.load(path)


def _gpu_ep(tag, seed=4321):
    import epstan.method as method
    model, Ktot, K, n_k, D, C, siter, niter = refs.CASES[tag]
    X, y, prior = refs.problem(tag)
    # Both runs take the damping factors the oracle run selected (cached `ep_df`): after a handful of iterations
    # the state is dominated by the first, large steps, so two runs whose noisy selection rule picked different
    # factors are not comparable (measured on cfg4s: KL 18.6 between runs that chose 0.35 / 0.0625 at iteration 2).
    # The selection rule itself is covered by tests/test_damping.py.
    dfs = np.asarray(_ref(tag)['ep_df'], dtype=np.float64)
    m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior,
                      chains=C, iter=siter, df0=lambda i: float(dfs[min(i, len(dfs)) - 1]))
    info, (ms, Ss), (st, mst, mrh, oth) = m.run(niter, verbose=False, seed=seed, return_analytics=True)
    assert info == 0
    return m, ms, Ss, mrh


def _gpu_target(tag):
    """full-data posterior of phi on the GPU: one multi-group site, cavity = prior"""
    from epstan.method import Worker
    model, Ktot, K, n_k, D, C, siter, niter = refs.CASES[tag]
    X, y, prior = refs.problem(tag)
    d = prior['Q'].shape[0]
    w = Worker(0, 'experiment/models/%s' % model, d, X, y,
               A={'J': K, 'j_ind': np.repeat(np.arange(K), n_k) + 1}, chains=8, iter=1000)
    assert w.cavity(np.asfortranarray(prior['Q']), np.asarray(prior['r'], dtype=np.float64),
                    np.zeros((d, d), order='F'), np.zeros(d))
    w.tilted(np.zeros((d, d), order='F'), np.zeros(d), save_samples=('phi',), seed=99)
    samp = w.saved_samp['phi']
    assert w.last_mrhat < 1.1, w.last_mrhat
    return samp.mean(axis=0), np.cov(samp, rowvar=False)


def test_cfg3_ep_vs_oracle_ep():
    """BASELINE configs[2]: m1b_sg, K=64, n_k=2000, D=19 (d=20), 8 chains x 200, 12 EP iterations."""
    ref = _ref('cfg3')
    m, ms, Ss, mrh = _gpu_ep('cfg3')
    # same algorithm and damping factors, different sampler implementation and seeds
    kl_ep = orc.kl_mvn(ref['ep_m'][-1], ref['ep_S'][-1], ms[-1], Ss[-1])
    sd = np.sqrt(np.diag(ref['ep_S'][-1]))
    z = np.abs(ms[-1] - ref['ep_m'][-1]) / sd
    print('cfg3: KL(oracle EP || GPU EP) %.4f  max |mean diff| / sd %.3f  max Rhat %s  df %s' % (
        kl_ep, z.max(), np.round(mrh, 3), np.round(m.history['df'], 4)))
    # d = 20; both runs carry the Monte Carlo noise of 800 draws per site and iteration, and the state after 12
    # iterations is dominated by the first step (damping 0.5): measured 0.52 between the two runs; 1.0 nat is a mean
    # shift of 0.3 posterior sd in every dimension
    assert kl_ep < 1.0, kl_ep
    assert z.max() < 1.0
    assert np.all(mrh[3:] < 1.2), mrh
    # the run has settled: successive global approximations differ by less than the noise of one iteration
    step = [orc.kl_mvn(ms[i], Ss[i], ms[i - 1], Ss[i - 1]) for i in range(1, len(ms))]
    assert max(step[-3:]) < 0.2, step


@pytest.mark.skipif(not os.environ.get('EPGPU_SLOW_TESTS'), reason='7 minutes on one SM: set EPGPU_SLOW_TESTS=1')
def test_cfg3_ep_vs_full_data_posterior():
    """(ii) the EP approximation against the full-data posterior of phi (one multi-group site of 128 000 rows,
    8 x 500 draws by the GPU sampler: one CTA, ~7 min)."""
    ref = _ref('cfg3')
    m, ms, Ss, mrh = _gpu_ep('cfg3')
    tm, tS = _gpu_target('cfg3')
    kl_tgt = orc.kl_mvn(tm, tS, ms[-1], Ss[-1])
    kl_tgt_oracle = orc.kl_mvn(tm, tS, ref['ep_m'][-1], ref['ep_S'][-1])
    z = np.abs(ms[-1] - tm) / np.sqrt(np.diag(tS))
    print('cfg3: KL(target || GPU EP) %.4f  KL(target || oracle EP) %.4f  max |mean - target| / sd %.3f' % (
        kl_tgt, kl_tgt_oracle, z.max()))
    assert kl_tgt < max(1.0, 2.0 * kl_tgt_oracle), (kl_tgt, kl_tgt_oracle)
    assert z.max() < 1.0


def test_cfg4_subset_ep_vs_oracle_ep():
    """First 16 sites of BASELINE configs[3]: m3b_sg, n_k=5000, D=49 (d=50, 100 sampled parameters per site),
    4 chains x 200, 5 EP iterations with the damping factors of the oracle run (0.5, 0.35, 0.0625 x 3).

    After five iterations the state is dominated by the first two, large steps, whose tilted moments come from
    chains that have not mixed yet (max split-Rhat 1.5 / 1.16 in iterations 1 / 2, in the fp64 oracle as on the
    GPU: 100 warm-up iterations from U(-2,2) on a 100-dimensional funnel): two runs of the SAME sampler with
    different seeds differ by KL ~ 15 nat here.  The parity statement that can be made at this cost (the oracle
    run takes an hour) is therefore relative: the oracle run is no further from a GPU run than a second GPU run
    with another seed is."""
    ref = _ref('cfg4s')
    m, ms, Ss, mrh = _gpu_ep('cfg4s')
    m2, ms2, Ss2, mrh2 = _gpu_ep('cfg4s', seed=8765)
    sym = lambda a, A, b, B: 0.5 * (orc.kl_mvn(a, A, b, B) + orc.kl_mvn(b, B, a, A))
    kl_ep = min(sym(ref['ep_m'][-1], ref['ep_S'][-1], ms[-1], Ss[-1]), sym(ref['ep_m'][-1], ref['ep_S'][-1], ms2[-1], Ss2[-1]))
    kl_gg = sym(ms[-1], Ss[-1], ms2[-1], Ss2[-1])
    sd = np.sqrt(np.diag(ref['ep_S'][-1]))
    z = np.minimum(np.abs(ms[-1] - ref['ep_m'][-1]), np.abs(ms2[-1] - ref['ep_m'][-1])) / sd
    zg = np.abs(ms[-1] - ms2[-1]) / sd
    print('cfg4s: sym. KL(oracle EP, GPU EP) %.3f  sym. KL(GPU EP seed 1, seed 2) %.3f  max |mean diff| / sd: oracle-GPU %.3f, '
          'GPU-GPU %.3f  max Rhat %s / %s  df %s' % (kl_ep, kl_gg, z.max(), zg.max(), np.round(mrh, 3), np.round(mrh2, 3),
                                                     np.round(m.history['df'], 4)))
    assert np.allclose(m.history['df'], ref['ep_df'], rtol=1e-9)          # no positive-definiteness retries on either side
    assert kl_ep < 2.0 * kl_gg + 1.5, (kl_ep, kl_gg)
    assert z.max() < 2.0 * zg.max() + 0.5, (z.max(), zg.max())
    assert np.all(mrh[2:] < 1.2), mrh
