"""GPU tilted densities and the batched NUTS sampler.

The sampling half has no pinned reference output (PyStan is absent: "parity
unpinned"); the checks are
  (1) log-density / gradient vs the fp64 oracle (oracle/density.py), fp32 tolerance;
  (2) sampler moments vs a long run of the fp64 oracle NUTS within 4 x MCSE
      (north_star check (b) with the oracle standing in for PyStan);
  (3) a full EP run vs the same EP run driven by the oracle sampler: KL between
      the final Gaussian approximations below a stated tolerance (check (c)).
"""

import numpy as np
import pytest

from oracle import density as dens
from oracle import ep_linalg as orc
from oracle import nuts
import synth
import oracle_refs

pytestmark = pytest.mark.gpu


def make_ctx(model, sites, use_tc=None):
    from epstan import _lib
    K = len(sites)
    d = sites[0]['d']
    D = sites[0]['X'].shape[1]
    ctx = _lib.Context(0)
    ctx.init_state(K, d)
    k_lim = np.concatenate(([0], np.cumsum([s['X'].shape[0] for s in sites])))
    X = np.concatenate([s['X'] for s in sites])
    y = np.concatenate([s['y'] for s in sites])
    multi = any(s['J'] > 1 for s in sites)
    ctx.upload_sites(_lib.MODEL_IDS[model], D, k_lim, X, y,
                     np.concatenate([s['j_ind'] for s in sites]) if multi else None,
                     [s['J'] for s in sites] if multi else None)
    ctx.upload(_lib.CAVQ, np.asfortranarray(np.stack([s['Omega'] for s in sites], axis=2)))
    ctx.upload(_lib.CAVM, np.asfortranarray(np.stack([s['mu'] for s in sites], axis=1)))
    if use_tc is not None:
        ctx.set_option('use_tc', use_tc)
    return ctx


@pytest.mark.parametrize('model', dens.MODELS)
@pytest.mark.parametrize('J,n,D', [(1, 37, 3), (1, 700, 19), (4, 90, 5), (3, 1300, 49)])
def test_logdensity_parity(model, J, n, D):
    sites = [synth.make_site(model, n, D, J, seed=11), synth.make_site(model, n + 13, D, J, seed=12)]
    ctx = make_ctx(model, sites, use_tc=0)          # the fp32 SIMT pass
    rng = np.random.RandomState(1)
    for k, site in enumerate(sites):
        td = synth.oracle_density(model, site)
        assert ctx.num_params(k) == td.p
        q = 0.4 * rng.standard_normal((40, td.p))       # > 32: exercises batching
        lp, grad = ctx.logdensity(k, q)
        olp, ograd = td.lp_grad(q)
        # fp32 contractions over n rows: a few 1e-6 relative per term
        assert np.max(np.abs(lp - olp) / np.maximum(1.0, np.abs(olp))) < 5e-5
        assert np.max(np.abs(grad - ograd)) < 2e-4 * max(1.0, np.max(np.abs(ograd)))
    ctx.close()


@pytest.mark.parametrize('model', dens.MODELS)
@pytest.mark.parametrize('n,D', [(37, 3), (128, 8), (700, 19), (5000, 49), (1300, 63)])
def test_logdensity_parity_tensor_core(model, n, D):
    """tcgen05/TMA likelihood pass: X is stored in bf16, centred per site (a data quantisation
    of the deviations from the column means; the coefficients stay fp32-accurate through the
    hi/lo split), so the check is against the fp64 oracle evaluated on the SAME stored inputs.
    The gradient carries bf16 rounding of the residuals E (~2^-9 relative)."""
    sites = [synth.make_site(model, n, D, 1, seed=31), synth.make_site(model, n + 77, D, 1, seed=32)]
    ctx = make_ctx(model, sites, use_tc=1)
    rng = np.random.RandomState(2)
    for k, site in enumerate(sites):
        site_q = dict(site, X=synth.tc_stored_X(site['X']))
        td = synth.oracle_density(model, site_q)
        q = 0.4 * rng.standard_normal((21, td.p))       # > 16: exercises batching
        lp, grad = ctx.logdensity(k, q)
        olp, ograd = td.lp_grad(q)
        assert np.max(np.abs(lp - olp) / np.maximum(1.0, np.abs(olp))) < 5e-5
        assert np.max(np.abs(grad - ograd)) < 6e-3 * max(1.0, np.max(np.abs(ograd)))
    ctx.close()


@pytest.mark.parametrize('model,n,D', [('m1b', 300, 64), ('m1b', 1300, 127), ('m3b', 700, 130),
                                       ('m1b', 2100, 199), ('m3b', 900, 199), ('m4b', 500, 99),
                                       ('m2b', 640, 90), ('m2b', 37, 255)])
def test_logdensity_parity_wide_tensor_core(model, n, D):
    """Wide tcgen05/TMA pass (csrc/epg_lik_tcw.cuh; config 5: D+1 up to 256 in 64-column sub-tiles whose GEMM1
    partial products accumulate in TMEM, 32 chains): same check as above, against the fp64 oracle on the
    stored (centred, bf16) inputs.  Shapes cover 2, 3 and 4 sub-tiles, a last sub-tile with a single
    K-step (D+1 = 65), the full 256 columns and ragged last row tiles."""
    sites = [synth.make_site(model, n, D, 1, seed=61), synth.make_site(model, n + 77, D, 1, seed=62)]
    for s in sites:                                  # (keep the cavity term moderate at d = 200)
        s['Omega'] = s['Omega'] / max(1.0, np.abs(s['Omega']).max()) + np.eye(s['d'])
    ctx = make_ctx(model, sites, use_tc=1)
    rng = np.random.RandomState(5)
    for k, site in enumerate(sites):
        site_q = dict(site, X=synth.tc_stored_X(site['X']))
        td = synth.oracle_density(model, site_q)
        q = (0.4 / np.sqrt(max(D, 16) / 16.0)) * rng.standard_normal((40, td.p))     # > 32: exercises batching
        lp, grad = ctx.logdensity(k, q)
        olp, ograd = td.lp_grad(q)
        assert np.max(np.abs(lp - olp) / np.maximum(1.0, np.abs(olp))) < 5e-5
        assert np.max(np.abs(grad - ograd)) < 6e-3 * max(1.0, np.max(np.abs(ograd)))
    ctx.close()


def test_tensor_core_pass_keeps_shifted_inputs():
    """Inputs like 35.5 +- 0.02 (the simulators shift the inputs of groups with extreme intercepts,
    common.py:132-317): plain bf16 storage (8 significant bits) would erase the within-site variation;
    the centred copy keeps it -- parity with the fp64 oracle on the TRUE inputs, not only the stored ones."""
    model, n, D = 'm3b', 2000, 49
    site = synth.make_site(model, n, D, 1, seed=51)
    rng = np.random.RandomState(4)
    site['X'] = 35.5 + 0.02 * rng.standard_normal((n, D))
    f = 0.3 + (site['X'] - 35.5) @ (20.0 * rng.standard_normal(D))
    site['y'] = (rng.uniform(size=n) < 1 / (1 + np.exp(-f))).astype(np.int64)
    ctx = make_ctx(model, [site, site], use_tc=1)
    td = synth.oracle_density(model, site)
    td_stored = synth.oracle_density(model, dict(site, X=synth.tc_stored_X(site['X'])))
    td_plain = synth.oracle_density(model, dict(site, X=synth.bf16_round(site['X'])))
    q = 0.05 * rng.standard_normal((8, td.p))
    q[:, td.d + 1:] = 10.0 * rng.standard_normal((8, D))      # slopes that resolve the 0.02 spread
    q[:, 1:td.d] = 0.1 * rng.standard_normal((8, D))
    # keep the mean part of the predictor moderate: sum(beta) ~ 0
    beta = q[:, td.d + 1:] * np.exp(q[:, 1:td.d])
    q[:, td.d + 1] -= beta.sum(axis=1) / np.exp(q[:, 1])
    lp, grad = ctx.logdensity(0, q)
    olp, ograd = td.lp_grad(q)
    slp, sgrad = td_stored.lp_grad(q)
    plp, _ = td_plain.lp_grad(q)
    assert np.max(np.abs(lp - slp) / np.maximum(1.0, np.abs(slp))) < 5e-5
    assert np.max(np.abs(lp - olp) / np.maximum(1.0, np.abs(olp))) < 2e-3       # true inputs: bf16 of the deviations
    assert np.max(np.abs(plp - olp) / np.maximum(1.0, np.abs(olp))) > 10 * np.max(np.abs(lp - olp) / np.maximum(1.0, np.abs(olp)))
    assert np.max(np.abs(grad - sgrad)) < 6e-3 * max(1.0, np.max(np.abs(sgrad)))
    ctx.close()


def _moment_check(draws_gpu, per_chain_gpu, ref):
    """4 x MCSE agreement of means; variances within 4 x their MC error.  `ref`: summary of the
    oracle run (mean, var, ess, mcse of phi; tests/oracle_refs.py)."""
    ess_g, mcse_g = nuts.ess_mcse(per_chain_gpu)
    tol = 4.0 * np.sqrt(mcse_g ** 2 + ref['mcse'] ** 2)
    dm = np.abs(draws_gpu.mean(axis=0) - ref['mean'])
    assert np.all(dm < tol), (dm / tol).max()
    vg = draws_gpu.var(axis=0, ddof=1)
    rel = np.abs(vg / ref['var'] - 1.0)
    # batch-means ESS is optimistic for variances (they mix slower than means): halve it
    tolv = 4.0 * np.sqrt(2.0 / np.maximum(0.5 * ess_g, 10) + 2.0 / np.maximum(0.5 * ref['ess'], 10))
    assert np.all(rel < tolv), (rel / tolv).max()


@pytest.mark.parametrize('model,J,n,D,C', [('m1b', 1, 300, 4, 8), ('m3b', 1, 400, 3, 8),
                                           ('m4b', 2, 300, 3, 4), ('m1b', 5, 250, 6, 16),
                                           ('m2b', 3, 300, 4, 8), ('m5b', 1, 300, 3, 8),
                                           # wide tensor-core pass: > 16 chains (one sub-tile), D+1 > 64 (two)
                                           ('m1b', 1, 300, 4, 24), ('m1b', 1, 600, 70, 32),
                                           # BASELINE site shapes (configs 3 and 4) on the tensor-core pass
                                           ('m1b', 1, 2000, 19, 8), ('m3b', 1, 5000, 49, 4)])
def test_sampler_vs_oracle_nuts(model, J, n, D, C):
    site = synth.make_site(model, n, D, J, seed=21)
    NS = 6
    ctx = make_ctx(model, [site] * NS)           # identical sites: independent replicas
    iters, warm = 1000, 400
    seeds = [101 * (k + 1) for k in range(NS)]
    msteps, mrhat, nleap, secs = ctx.tilted_sample(seeds, C, iters, warm)
    per = iters - warm
    dr = ctx.get_draws(C * per)                  # (NS, d, n)
    assert np.all(np.isfinite(dr)) and np.all(msteps > 0) and np.all(nleap > 0)
    # the m3b/m4b tilted densities are funnels: an occasional chain lingers in the
    # neck (the fp64 oracle shows the same), so judge the replicas collectively
    assert np.median(mrhat) < 1.1, mrhat
    # fp64 oracle NUTS, 8 chains x 2500 iterations on the same site (cached: tests/oracle_refs.py)
    ref_phi = oracle_refs.summary(model, J, n, D)
    # step size and work per draw track the fp64 sampler
    assert 0.6 < np.median(msteps) / ref_phi['stepsize'] < 1.6
    assert 0.5 < np.median(nleap) / (C * iters) / ref_phi['evals_per_draw'] < 2.0
    failures = 0
    for k in range(NS):
        x = dr[k].T
        try:
            _moment_check(x, [x[c * per:(c + 1) * per] for c in range(C)], ref_phi)
        except AssertionError:
            failures += 1
    assert failures <= 1, failures
    # different seeds give different draws; the same seed reproduces them bit for bit
    assert not np.array_equal(dr[0], dr[1])
    ctx.tilted_sample(seeds, C, iters, warm)
    assert np.array_equal(ctx.get_draws(C * per), dr)
    ctx.close()


@pytest.mark.parametrize('model,D,C,NS', [('m1b', 6, 8, 7), ('m3b', 5, 4, 10), ('m4b', 19, 3, 5)])
def test_pingpong_kernel_matches_one_site_per_cta(model, D, C, NS):
    """The two-sites-per-CTA kernel (used when there are more sites than SMs) runs the same arithmetic and
    the same counter-based random streams as the one-site-per-CTA kernel: identical draws and analytics,
    for ragged site sizes and any assignment of sites to CTAs."""
    sites = [synth.make_site(model, 130 + 97 * k, D, 1, seed=40 + k) for k in range(NS)]
    ctx = make_ctx(model, sites, use_tc=1)
    seeds = [17 * (k + 3) for k in range(NS)]
    iters, warm = 120, 60
    out = {}
    for mode in (0, 2):
        ctx.set_option('pingpong', mode)
        res = ctx.tilted_sample(seeds, C, iters, warm)
        out[mode] = (ctx.get_draws(C * (iters - warm)).copy(),) + tuple(np.array(r) for r in res[:3])
    assert np.all(np.isfinite(out[0][0]))
    for u, v in zip(out[0], out[2]):
        assert np.array_equal(u, v)
    # a sub-range of the sites (k0 > 0) goes through the queue as well
    ctx.set_option('pingpong', 2)
    ctx.tilted_sample(seeds[2:], C, iters, warm, k0=2, k1=NS)
    assert np.array_equal(ctx.get_draws(C * (iters - warm))[2:], out[0][0][2:])
    ctx.close()


def test_init_prev_and_zero_init():
    site = synth.make_site('m1b', 200, 3, 1, seed=5)
    ctx = make_ctx('m1b', [site])
    ctx.tilted_sample([7], 4, 60, 30, init_mode=1)
    a = ctx.get_draws(4 * 30).copy()
    ctx.tilted_sample([8], 4, 60, 30, init_mode=2)         # continue from the last draws
    b = ctx.get_draws(4 * 30)
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b)) and not np.array_equal(a, b)
    ctx.close()


def test_reinit_sites_overrides_init_prev():
    """epg_reinit_sites: a marked site's chains start afresh (around the cavity mean) in the next init_prev run;
    the other sites continue from their last draws; the mark is consumed."""
    sites = [synth.make_site('m1b', 200, 3, 1, seed=5 + k) for k in range(3)]
    n = 4 * 30

    def run(mark, second_mode=2):
        ctx = make_ctx('m1b', sites)
        ctx.tilted_sample([7, 8, 9], 4, 60, 30, init_mode=0)
        if mark is not None:
            ctx.reinit_sites([mark])
        ctx.tilted_sample([17, 18, 19], 4, 60, 30, init_mode=second_mode)
        b = ctx.get_draws(n).copy()
        ctx.tilted_sample([27, 28, 29], 4, 60, 30, init_mode=2)
        c = ctx.get_draws(n).copy()
        ctx.close()
        return b, c

    b0, c0 = run(None)
    b1, c1 = run(1)
    assert np.array_equal(b0[0], b1[0]) and np.array_equal(b0[2], b1[2])       # unmarked sites: unchanged
    assert not np.array_equal(b0[1], b1[1])                                    # marked site: a fresh start
    # ... from which the sampler reaches the same distribution
    assert np.all(np.abs(b0[1].mean(axis=1) - b1[1].mean(axis=1)) < 5 * b0[1].std(axis=1) / np.sqrt(n / 10))
    assert np.array_equal(c0[0], c1[0]) and not np.array_equal(c0[1], c1[1])   # third run: init_prev everywhere again
    assert np.all(np.isfinite(b1)) and np.all(np.isfinite(c1))


@pytest.mark.parametrize('model,J,n,D', [('m1b', 1, 300, 4), ('m4b', 3, 240, 3), ('m2b', 2, 300, 3)])
def test_param_stats_for_mix_pred(model, J, n, D):
    """f4: the on-device moments of the transformed site parameters (epg_get_param_stats, what Master.mix_pred
    pools): the phi slots against the retained draws themselves, alpha / beta against the fp64 oracle NUTS
    (transformed per draw as the Stan programs do) within 5 Monte Carlo standard errors."""
    site = synth.make_site(model, n, D, J, seed=23)
    ctx = make_ctx(model, [site])
    ctx.set_option('param_stats', 1)
    C, iters, warm = 8, 700, 300
    ctx.tilted_sample([11], C, iters, warm)
    per = iters - warm
    ntot = C * per
    mean, ssd = ctx.param_stats()
    d = site['d']
    p = ctx.num_params(0)
    dr = ctx.get_draws(ntot)[0]                                        # (d, n)
    assert np.max(np.abs(mean[0, :d] - dr.mean(axis=1))) < 1e-5 * max(1.0, np.max(np.abs(dr)))
    ssd_ref = ((dr - dr.mean(axis=1, keepdims=True)) ** 2).sum(axis=1)
    assert np.max(np.abs(ssd[0, :d] - ssd_ref) / ssd_ref) < 2e-3        # (fp32 running sums)
    # alpha, beta: oracle NUTS (8 x 1500) on the same site, transformed as in the Stan programs
    # (cached: tests/oracle_refs.py, tests/golden/nuts_ref2.npz)
    ref = oracle_refs.param_stats_ref(model, J, n, D)
    om, osd = ref['mean'], ref['sd']
    assert om.shape[0] == p - d
    gm = mean[0, d:p]
    gsd = np.sqrt(ssd[0, d:p] / (ntot - 1))
    mcse = osd * np.sqrt(1.0 / 400 + 1.0 / 400)                        # ~400 effective draws on either side
    assert np.all(np.abs(gm - om) < 5 * mcse), (gm - om) / mcse
    assert np.all(np.abs(np.log(gsd / osd)) < 0.35), gsd / osd
    ctx.close()


_ep_problem = oracle_refs.ep_problem


@pytest.mark.parametrize('model', ['m1b', 'm4b'])
def test_full_ep_vs_oracle_ep(model):
    """check (c): same data, seed-independent settings, GPU EP vs oracle EP."""
    import epstan.method as method
    K, n_k, D, C, siter, niter = 4, 150, 3, 4, 400, 6
    X, y, prior, d = _ep_problem(model, K, n_k, D, seed=9)
    m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior,
                      chains=C, iter=siter, df0=0.6)
    info, (ms, Ss), (stimes, msteps, mrhats, other) = m.run(niter, verbose=False, seed=1, return_analytics=True)
    assert info == 0
    assert np.all(np.isfinite(ms)) and np.all(stimes > 0) and np.all(msteps > 0) and np.all(mrhats < 3.0)
    assert m.Qi.shape == (d, d, K) and np.allclose(m.Q, m.Q0 + m.Qi.sum(axis=2), rtol=1e-12, atol=1e-12)

    # oracle EP on the same problem: oracle NUTS (4 x 400 per site and iteration) + oracle moment
    # matching / updates, 6 iterations, damping 0.6 (cached: tests/oracle_refs.py)
    oref = oracle_refs.ep_reference(model)
    assert int(oref['info']) == 0
    oms, oSs = [oref['m']], [oref['S']]
    kl = orc.kl_mvn(oms[-1], oSs[-1], ms[-1], Ss[-1])
    # both runs carry Monte-Carlo noise from 4 x 200 draws per site per iteration
    assert kl < 0.25, kl
    sd = np.sqrt(np.diag(oSs[-1]))
    assert np.all(np.abs(ms[-1] - oms[-1]) < 0.6 * sd)


@pytest.mark.parametrize('use_tc', [0, 1])
def test_config5_shape_large_d_many_chains(use_tc):
    """config-5-like shape (D+1 = 200 > 64, 32 chains) on the fp32 SIMT pass and on the wide tensor-core pass:
    density parity at d=200 and a short adaptive run that must produce finite, well-mixed draws whose
    moments can be matched."""
    model, n, D, C = 'm1b', 1500, 199, 32
    site = synth.make_site(model, n, D, 1, seed=41)
    site['Omega'] = site['Omega'] + 4.0 * np.eye(D + 1)
    ctx = make_ctx(model, [site, site], use_tc=use_tc)
    # (the tensor-core pass holds the centred inputs in bf16: compare on the stored values)
    td = synth.oracle_density(model, dict(site, X=synth.tc_stored_X(site['X'])) if use_tc else site)
    rng = np.random.RandomState(3)
    q = 0.1 * rng.standard_normal((5, td.p))
    lp, grad = ctx.logdensity(0, q)
    olp, ograd = td.lp_grad(q)
    assert np.max(np.abs(lp - olp) / np.maximum(1.0, np.abs(olp))) < 5e-5
    assert np.max(np.abs(grad - ograd)) < (6e-3 if use_tc else 5e-4) * max(1.0, np.max(np.abs(ograd)))
    msteps, mrhat, nleap, secs = ctx.tilted_sample([5, 6], C, 120, 60)
    dr = ctx.get_draws(C * 60)
    assert dr.shape == (2, D + 1, C * 60) and np.all(np.isfinite(dr))
    assert np.all(msteps > 0) and np.all(mrhat < 1.5) and np.all(nleap > 0)
    oks, n_ok = ctx.moments(C * 60, 'sample')
    assert oks.all()
    if use_tc:
        # the two passes sample the same distribution (the SIMT pass on the unrounded inputs)
        ctx.set_option('use_tc', 0)
        ctx.tilted_sample([5, 6], C, 120, 60)
        dr0 = ctx.get_draws(C * 60)
        sd = np.sqrt(0.5 * (dr.var(axis=2) + dr0.var(axis=2)))
        z = np.abs(dr.mean(axis=2) - dr0.mean(axis=2)) / sd
        assert z.max() < 0.35, z.max()            # 1920 autocorrelated draws each: MCSE of a mean ~ 0.05 sd
    ctx.close()


def test_control_max_treedepth():
    """`control` (PyStan's own `sampling` keyword; an extension, the reference never sets it): the tree-depth
    cap bounds the leapfrogs per transition; unknown controls raise."""
    import epstan.method as method
    K, n_k, D, C, siter = 3, 150, 3, 4, 60
    X, y, prior, d = _ep_problem('m1b', K, n_k, D, seed=9)
    kw = dict(site_sizes=np.full(K, n_k), prior=prior, chains=C, iter=siter, df0=0.5)
    m = method.Master('experiment/models/m1b_sg', X, y, control={'max_treedepth': 2}, **kw)
    assert m.run(1, verbose=False, seed=1, calc_moments=False) == 0
    # at most 2^2 - 1 leapfrogs per transition (+ the step-size heuristic's evaluations at the start)
    assert all(0 < w.last_n_leapfrog <= C * (siter * 3 + 60) for w in m.workers)
    m2 = method.Master('experiment/models/m1b_sg', X, y, **kw)
    assert m2.run(1, verbose=False, seed=1, calc_moments=False) == 0
    assert sum(w.last_n_leapfrog for w in m2.workers) > sum(w.last_n_leapfrog for w in m.workers)
    with pytest.raises(ValueError):
        method.Master('experiment/models/m1b_sg', X, y, control={'stepsize': 0.1}, **kw)
