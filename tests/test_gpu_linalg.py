"""GPU parity of the moment-matching / cavity / update path (SURVEY 8a rows
a5-a13) through the drop-in ``epstan`` API, against

  * tests/golden/*.npz  -- outputs of the unmodified reference, and
  * oracle/ep_linalg.py -- the pinned CPU restatement, on larger seeded inputs.

Tolerance: <= 1e-10 relative in fp64 (north_star (a)).
"""

import numpy as np
import pytest

from conftest import relerr
import golden_inputs as gi
from oracle import ep_linalg as orc
from oracle import fakes

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope='module')
def ep():
    import epstan.method
    import epstan.util
    return epstan


def test_kat_dpotri(ep, golden):
    """reference test_scipy.py:22-31"""
    out, _ = ep.util.invert_normal_params(np.asfortranarray(np.eye(6) * 2.0), cho_form=True)
    assert np.allclose(out, np.eye(6) * 0.25, atol=1e-12)


@pytest.mark.parametrize('tag,seed,d', gi.LINALG_CASES)
def test_invert_and_olse(ep, golden, tag, seed, d):
    g = golden['linalg']
    c = gi.linalg_case(seed, d)
    u = ep.util
    Q, r = u.invert_normal_params(c['S'], c['m'])
    assert Q.flags['FARRAY']
    assert relerr(Q, g['inv_%s_Q' % tag]) < TOL
    assert relerr(r, g['inv_%s_r' % tag]) < TOL
    assert np.array_equal(Q, Q.T)                      # copy_triu_to_tril semantics
    Q2, r2 = u.invert_normal_params(np.asfortranarray(c['Ug']), c['m'], cho_form=True)
    assert relerr(Q2, g['inv_%s_choform_Q' % tag]) < TOL
    assert relerr(r2, g['inv_%s_choform_r' % tag]) < TOL
    # in-place / explicit output forms
    A = np.asfortranarray(c['S'].copy())
    b = c['m'].copy()
    oA, ob = u.invert_normal_params(A, b, out_A='in-place', out_b='in-place')
    assert oA is A and ob is b and relerr(A, g['inv_%s_Q' % tag]) < TOL
    assert relerr(u.olse(np.asfortranarray(c['S']), c['n'], P=np.asfortranarray(c['P'])), g['olse_%s_P' % tag]) < TOL
    assert relerr(u.olse(np.asfortranarray(c['S']), c['n']), g['olse_%s_naive' % tag]) < TOL
    with pytest.raises(u.LinAlgError):
        u.invert_normal_params(c['S'] - 10 * np.eye(d))


@pytest.mark.parametrize('tag,seed,d,C,it', gi.WORKER_CASES)
def test_worker_cavity_and_tilted(ep, golden, tag, seed, d, C, it):
    """The calls of make_golden.gen_worker, replayed on the GPU Worker."""
    g = golden['worker']
    c = gi.worker_case(seed, d, C, it)
    Worker = ep.method.Worker
    Q = np.asfortranarray(c['Q'])
    for k in range(c['K']):
        for mode in ('sample', 'olse'):
            model = fakes.FakeModel(c['Qs'], c['rs'])
            model.quiet = True
            w = Worker(k, model, d, np.zeros((4, 2)), np.zeros(4), A={'site_id': k},
                       chains=C, iter=it, prec_estim=mode)
            with pytest.raises(RuntimeError):
                w.tilted(np.zeros((d, d), order='F'), np.zeros(d))
            assert w.cavity(Q, c['r'], np.asfortranarray(c['Qi'][k]), c['ri'][k])
            assert relerr(w.Mat, g['wrk_%s_k%d_cavQ' % (tag, k)]) < TOL
            assert relerr(w.vec, g['wrk_%s_k%d_cavm' % (tag, k)]) < TOL
            dQi = np.zeros((d, d), order='F')
            dri = np.zeros(d)
            assert w.tilted(dQi, dri, seed=c['seeds'][k])
            assert model.seeds[0][1] == int(g['wrk_%s_k%d_stanseed' % (tag, k)])
            assert relerr(dQi, g['wrk_%s_k%d_%s_dQi' % (tag, k, mode)]) < TOL
            assert relerr(dri, g['wrk_%s_k%d_%s_dri' % (tag, k, mode)]) < TOL
            assert w.phase == 2 and w.iteration == 1 and w.nsamp == c['n']
    w = Worker(0, fakes.FakeModel(c['Qs'], c['rs']), d, np.zeros((4, 2)), np.zeros(4))
    assert w.cavity(Q, c['r'], np.asfortranarray(c['Q'] + np.eye(d)), c['r']) is False
    assert w.phase == 0


def build_master(ep, sc):
    Qs, rs = fakes.gaussian_site_factors(sc['fseed'], sc['K'], sc['d'])
    model = fakes.FakeModel(Qs, rs, inflate=sc.get('inflate'), constant=sc.get('constant', False))
    model.quiet = True
    K, d = sc['K'], sc['d']
    kw = dict(chains=sc['chains'], iter=sc['iter'])
    if sc['prior'] is not None:
        kw['prior'] = sc['prior']
    if sc['df0'] is not None:
        kw['df0'] = sc['df0']
    if sc['prec_estim'] != 'sample':
        kw.update(prec_estim=sc['prec_estim'], prec_estim_skip=sc['skip'])
    m = ep.method.Master(model, np.zeros((2 * K, 2)), np.zeros(2 * K), site_sizes=np.full(K, 2),
                         dphi=d, A_k={'site_id': list(range(K))}, **kw)
    if sc.get('improper'):
        m.iter = 1
        m.Qi[:, :, 0] = 6.0 * np.eye(d)
        m.Qi[:, :, 1] = -4.0 * np.eye(d)
        m.Q[:] = m.Q0 + m.Qi.sum(axis=2)
        for k, w in enumerate(m.workers):
            if m._shard.k_begin <= k < m._shard.k_end:      # (multi-rank runs: local sites only)
                w.cavity(m.Q, m.r, m.Qi[:, :, k], m.ri[:, k])
            w.phase = 1
    return m


@pytest.mark.parametrize('tag', ['runA', 'runB', 'runC', 'runD', 'runE', 'runF'])
def test_master_run_vs_reference(ep, golden, tag):
    """Master.run against the exact Gaussian sampler: every INFO_* branch."""
    g = golden['master']
    scs = gi.master_scenarios()
    sc = scs[tag]
    m = build_master(ep, sc)
    info, (ms, Ss) = m.run(sc['niter'], verbose=False, seed=sc['seed'])
    assert info == int(g[tag + '_info'])
    assert relerr(ms, g[tag + '_m']) < TOL
    assert relerr(Ss, g[tag + '_S']) < TOL
    if info == 0:
        assert relerr(m.Q, g[tag + '_Q']) < TOL
        assert relerr(m.r, g[tag + '_r']) < TOL
        assert relerr(m.Qi, g[tag + '_Qi']) < TOL
        assert relerr(m.ri, g[tag + '_ri']) < TOL
        assert m.Qi.flags['F_CONTIGUOUS'] and m.Qi.shape == (sc['d'], sc['d'], sc['K'])
    if tag == 'runC':
        sc2 = scs['runC2']
        res = m.run(sc2['niter'], verbose=False, seed=sc2['seed'], return_analytics=True)
        assert res[0] == int(g['runC2_info'])
        assert relerr(res[1][0], g['runC2_m']) < TOL
        assert relerr(res[1][1], g['runC2_S']) < TOL
        assert len(res[2]) == 4 and res[2][0].shape == (sc2['niter'],)
        assert m.iter == sc['niter'] + sc2['niter']


def test_master_argument_errors(ep):
    Master = ep.method.Master
    X, y = np.zeros((6, 2)), np.zeros(6)
    model = fakes.FakeModel(*fakes.gaussian_site_factors(1, 3, 2))
    with pytest.raises(TypeError):
        Master(model, X, y, site_sizes=[2, 2, 2], dphi=2, bogus=1)
    with pytest.raises(ValueError):
        Master(model, X, y, site_sizes=[2, 2, 2])                  # neither prior nor dphi
    with pytest.raises(ValueError):
        Master(model, X, y, site_sizes=[6], dphi=2)                # K < 2
    with pytest.raises(ValueError):
        Master(model, X, y, site_sizes=[2, 2, 1], dphi=2)          # sizes do not cover X
    with pytest.raises(ValueError):
        Master(model, X, y, site_sizes=[2, 2, 2], dphi=2, df0=1.5)
    with pytest.raises(NotImplementedError):
        Master(model, X, y, dphi=2)
    with pytest.raises(ValueError):
        Master(model, X, y, site_sizes=[2, 2, 2], prior={'Q': -np.eye(2), 'r': np.zeros(2)})
    assert Master(model, X, y, site_sizes=[2, 2, 2], dphi=2, A_k={'site_id': [0, 1, 2]}).run(0, verbose=False) \
        == (0, (None, None))


@pytest.mark.parametrize('K,d,n', [(64, 20, 800), (1024, 50, 800), (16, 200, 3200), (5, 7, 30), (3, 1, 10),
                                   (6, 120, 600), (5, 150, 640), (4, 199, 802)])  # DMMA Gram: one CTA (10 tiles), two CTAs (15, 28 tiles)
def test_moments_and_update_full_size(ep, K, d, n):
    """BASELINE config shapes (2, 4, 5) and ragged small ones, driven through the
    batched C-ABI calls; every site compared with the oracle."""
    from epstan import _lib
    rng = np.random.RandomState(1000 + K + d)
    ctx = _lib.Context(0)
    ctx.init_state(K, d)
    Q0 = fakes.random_spd(rng, d)
    r0 = rng.standard_normal(d)
    st = orc.EPState(Q0, r0, K)
    st.Qi = np.stack([0.05 * fakes.random_spd(rng, d) for _ in range(K)], axis=2)
    st.ri = 0.05 * rng.standard_normal((d, K))
    st.Q = Q0 + st.Qi.sum(axis=2)
    st.r = r0 + st.ri.sum(axis=1)
    # per-site draws with site-specific mean / scale, chain-like offset structure
    scale = np.exp(0.3 * rng.standard_normal((K, d, 1)))
    draws = rng.standard_normal((K, d, n)) * scale * 0.3 + 3.0 * rng.standard_normal((K, d, 1))
    ctx.upload(_lib.Q0, Q0)
    ctx.upload(_lib.R0, r0)
    ctx.upload(_lib.Q, st.Q)
    ctx.upload(_lib.R, st.r)
    ctx.upload(_lib.QI, np.asfortranarray(st.Qi))
    ctx.upload(_lib.RI, np.asfortranarray(st.ri))
    ctx.set_draws(draws, n)
    # (olse with a prior is degenerate for d == 1: its denominator f2*f2p - trSP^2 is 0)
    for mode in (('sample', 'olse') if d > 1 else ('sample',)):
        oks, n_ok = ctx.moments(n, mode)
        dQi = ctx.download(_lib.DQI, np.empty((d, d, K), order='F'))
        dri = ctx.download(_lib.DRI, np.empty((d, K), order='F'))
        assert oks.all() and n_ok == K
        worst = 0.0
        for k in range(K):
            ok, oQ, orr = orc.tilted_moments(draws[k].T, st.Q, st.r, mode)
            assert ok
            st.dQi[:, :, k], st.dri[:, k] = oQ, orr
            worst = max(worst, relerr(dQi[:, :, k], oQ), relerr(dri[:, k], orr))
        assert worst < TOL, (mode, worst)
    # damped update attempts + cavities + global moments (a10-a12): a large and a small df
    for df in (0.9, min(0.3, 1.0 / K)):
        ctx.update_partial(df)
        pd = ctx.update_finish()
        Qi2 = st.Qi + df * st.dQi
        ri2 = st.ri + df * st.dri
        Qn = Q0 + Qi2.sum(axis=2)
        rn = r0 + ri2.sum(axis=1)
        try:
            orc._chol_upper(Qn)
            opd = True
        except orc.NotPosDef:
            opd = False
        assert pd == opd
        assert relerr(ctx.download(_lib.QI2, np.empty((d, d, K), order='F')), Qi2) < 1e-14
        assert relerr(ctx.download(_lib.Q, np.empty((d, d), order='F')), Qn) < 1e-13
        if not pd:
            continue
        flags, all_ok = ctx.cavity(proposal=True)
        cavQ = ctx.download(_lib.CAVQ, np.empty((d, d, K), order='F'))
        cavm = ctx.download(_lib.CAVM, np.empty((d, K), order='F'))
        for k in range(K):
            ok, P, mu = orc.cavity(Qn, rn, Qi2[:, :, k], ri2[:, k])
            assert ok == flags[k]
            assert relerr(cavQ[:, :, k], P) < TOL
            if ok:
                assert relerr(cavm[:, k], mu) < TOL
        m = np.empty(d)
        S = np.empty((d, d), order='F')
        ctx.global_moments(m, S)
        oS, om = orc.invert_normal_params(Qn, rn)
        assert relerr(m, om) < TOL and relerr(S, oS) < TOL
    ctx.close()


@pytest.mark.parametrize('d', [1, 2, 5, 50, 200])
def test_force_pd_min_eig(ep, d):
    """a12 forcing branch: lambda_min by tridiagonalisation + bisection."""
    from epstan import _lib
    rng = np.random.RandomState(d)
    K = 6
    ctx = _lib.Context(0)
    ctx.init_state(K, d)
    Qi2 = np.empty((d, d, K), order='F')
    for k in range(K):
        A = rng.standard_normal((d, d))
        Qi2[:, :, k] = (A + A.T) * 0.5 + (k - 2.5) * np.eye(d) * np.sqrt(d)
    Qi = np.asfortranarray(rng.standard_normal((d, d, K)))
    ctx.upload(_lib.QI2, Qi2)
    ctx.upload(_lib.QI, Qi)
    forced, lam = ctx.force_pd(1e-5, 0.5)
    out = ctx.download(_lib.QI, np.empty((d, d, K), order='F'))
    for k in range(K):
        ref = orc.min_eig(Qi2[:, :, k])
        scale = np.abs(np.linalg.eigvalsh(Qi2[:, :, k])).max()
        assert abs(lam[k] - ref) < 1e-12 * max(scale, 1.0)
        assert forced[k] == (ref < 1e-5)
        exp = Qi[:, :, k].copy()
        if forced[k]:
            exp[np.arange(d), np.arange(d)] += 0.5 - ref
        assert relerr(out[:, :, k], exp) < 1e-11
    ctx.close()


@pytest.mark.parametrize('K,d', [(8, 6), (64, 20), (256, 50)])
def test_damp_sweep(ep, K, d):
    """a13: find_damp.py sweep over 31 damping values."""
    from epstan import _lib
    rng = np.random.RandomState(7 * K + d)
    ctx = _lib.Context(0)
    ctx.init_state(K, d)
    Q0 = fakes.random_spd(rng, d)
    r0 = rng.standard_normal(d)
    st = orc.EPState(Q0, r0, K)
    st.Qi = np.stack([0.1 * fakes.random_spd(rng, d) for _ in range(K)], axis=2)
    st.ri = 0.1 * rng.standard_normal((d, K))
    # deltas large enough that big damping factors break some cavities
    st.dQi = np.stack([(A + A.T) * 0.2 for A in rng.standard_normal((K, d, d))], axis=2) / np.sqrt(d)
    st.dri = 0.1 * rng.standard_normal((d, K))
    for aid, arr in ((_lib.Q0, Q0), (_lib.R0, r0), (_lib.QI, st.Qi), (_lib.RI, st.ri),
                     (_lib.DQI, st.dQi), (_lib.DRI, st.dri)):
        ctx.upload(aid, np.asfortranarray(arr))
    dfs = np.linspace(0, 1, 33)[1:-1]
    S_t = fakes.random_spd(rng, d, scale=0.1)
    m_t = 0.1 * rng.standard_normal(d)
    mse, kl = ctx.damp_sweep(dfs, m_t, S_t)
    omse, okl = orc.damp_sweep(st, dfs, m_t, S_t)
    assert np.array_equal(np.isnan(mse), np.isnan(omse))
    assert np.array_equal(np.isnan(kl), np.isnan(okl))
    good = ~np.isnan(omse)
    assert good.any()
    assert relerr(mse[good], omse[good]) < TOL
    assert relerr(kl[good], okl[good]) < TOL
    ctx.close()


@pytest.mark.parametrize('tag,seed,n,d', gi.CV_CASES)
def test_cv_moments_vs_reference(ep, golden, tag, seed, n, d):
    """a8: util.cv_moments (control-variate moments) against the reference's outputs."""
    g = golden['cv']
    c = gi.cv_case(seed, n, d)
    Q2, r2 = orc.invert_normal_params(c['S2'], c['m2'])
    # north_star check (a): 1e-10 relative in fp64 (the d2 x d2 Gram system has cond ~4e4 at d = 20)
    tol = 1e-10
    for mcv in (True, False):
        t = '%s_%s' % (tag, 'multi' if mcv else 'single')
        S, m, used = ep.util.cv_moments(c['samp'].copy(), c['lp'], Q2, r2, multiple_cv=mcv)
        assert used == bool(g['cv_%s_used' % t])
        assert relerr(m, g['cv_%s_m' % t]) < tol
        assert relerr(S, g['cv_%s_S' % t]) < tol
        assert np.array_equal(S, S.T)
        # ret_a=True: the coefficient arrays of the reference (same shapes; the linear solve behind them has a
        # condition number of ~4e4 at d = 20, hence 1e-8 on the coefficients themselves)
        S2_, m2_, used2, a_S, a_m = ep.util.cv_moments(c['samp'].copy(), c['lp'], Q2, r2, multiple_cv=mcv, ret_a=True)
        assert used2 == used and np.array_equal(S2_, S) and np.array_equal(m2_, m)
        assert a_S.shape == g['cv_%s_aS' % t].shape and relerr(a_S, g['cv_%s_aS' % t]) < 1e-8
        assert a_m.shape == g['cv_%s_am' % t].shape and relerr(a_m, g['cv_%s_am' % t]) < 1e-8
    Q3, r3 = orc.invert_normal_params(c['S2'], c['m3'])
    S, m, used = ep.util.cv_moments(c['samp'].copy(), c['lp'], Q3, r3)
    assert ep.util.cv_moments(c['samp'].copy(), c['lp'], Q3, r3, ret_a=True)[3:] == (0, 0)
    assert used is False
    assert relerr(m, g['cv_%s_fallback_m' % tag]) < TOL
    assert relerr(S, g['cv_%s_fallback_S' % tag]) < TOL


def test_cv_moments_batched_config2(ep):
    """config 2 shape: K=64 sites, n=800, d=20 in one batched call vs the oracle."""
    from epstan import _lib
    K, n, d = 64, 800, 20
    rng = np.random.RandomState(77)
    draws = np.empty((K, d, n)); lps = np.empty((K, n)); Qt = np.empty((K, d, d)); rt = np.empty((K, d))
    cases = []
    from scipy.stats import multivariate_normal
    for k in range(K):
        S1 = fakes.random_spd(rng, d); m1 = rng.standard_normal(d)
        S2 = S1 + 0.1 * fakes.random_spd(rng, d); m2 = m1 + 0.1 * rng.standard_normal(d)
        samp = m1 + rng.standard_normal((n, d)) @ np.linalg.cholesky(S1).T
        lp = multivariate_normal(mean=m1, cov=S1).logpdf(samp)
        Q2, r2 = orc.invert_normal_params(S2, m2)
        draws[k], lps[k], Qt[k], rt[k] = samp.T, lp, Q2, r2
        cases.append((samp, lp, Q2, r2))
    ctx = _lib.Context(0)
    for mcv in (True, False):
        S, m, used = ctx.cv_moments(draws, lps, Qt, rt, multiple_cv=mcv, regulate_a=0.9, max_a=5.0)
        for k in (0, 17, 63):
            oS, om, oused = orc.cv_moments(*cases[k], multiple_cv=mcv, regulate_a=0.9, max_a=5.0)
            assert bool(used[k]) == oused
            assert relerr(m[k], om) < 1e-10 and relerr(S[k], oS) < 1e-10
    ctx.close()


@pytest.mark.parametrize('K,d,n', [(7, 5, 30), (64, 20, 800), (3, 50, 401)])
def test_mix_phi_sums(ep, K, d, n):
    """f4: the pooled sums behind Master.mix_phi (method.py:1280-1298) from the device draw buffer."""
    from epstan import _lib
    rng = np.random.RandomState(5)
    draws = rng.standard_normal((K, d, n)) * 0.3 + rng.standard_normal((K, d, 1))
    ctx = _lib.Context(0)
    ctx.init_state(K, d)
    ctx.set_draws(draws, n)
    sums = ctx.mix_phi_sums(n)
    means = draws.mean(axis=2)
    sS = np.zeros((d, d)); sMM = np.zeros((d, d))
    for k in range(K):
        xc = draws[k] - means[k][:, None]
        sS += xc @ xc.T
        sMM += np.outer(means[k], means[k])
    assert relerr(sums[:d], means.sum(axis=0)) < 1e-12
    assert relerr(sums[d:d + d * d].reshape(d, d, order='F'), sS) < 1e-11
    assert relerr(sums[d + d * d:].reshape(d, d, order='F'), sMM) < 1e-12
    ctx.close()
