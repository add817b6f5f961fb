"""Generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference's epstan package is copied to a scratch directory, its Cython
module compiled there, and imported next to a two-line ``pystan`` stub (the
reference imports PyStan at module top: epstan/util.py:34, method.py:40) and an
empty ``matplotlib`` stub (experiment/find_damp.py:7).  Inputs are derived from
fixed ``np.random.RandomState`` seeds so only the seeds and the reference's
OUTPUTS are stored; tests regenerate the inputs.

TEST INFRASTRUCTURE ONLY.
"""

import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np
from scipy import linalg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('EP_REFERENCE', '/root/reference')
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)

from oracle import fakes  # noqa: E402


def import_reference():
    scratch = tempfile.mkdtemp(prefix='epref_')
    shutil.copytree(os.path.join(REF, 'epstan'), os.path.join(scratch, 'epstan'))
    shutil.copy(os.path.join(REF, 'setup.py'), scratch)
    subprocess.run([sys.executable, 'setup.py', 'build_ext', '--inplace'],
                   cwd=scratch, check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    os.makedirs(os.path.join(scratch, 'pystan'))
    with open(os.path.join(scratch, 'pystan', '__init__.py'), 'w') as f:
        f.write('class StanModel: pass\nfrom . import constants\n')
    with open(os.path.join(scratch, 'pystan', 'constants.py'), 'w') as f:
        f.write('MAX_UINT = 2**31 - 1\n')
    os.makedirs(os.path.join(scratch, 'matplotlib'))
    open(os.path.join(scratch, 'matplotlib', '__init__.py'), 'w').close()
    open(os.path.join(scratch, 'matplotlib', 'pyplot.py'), 'w').close()
    sys.path.insert(0, scratch)
    sys.path.insert(0, os.path.join(REF, 'experiment'))
    import warnings
    warnings.simplefilter('ignore')
    import epstan.util
    import epstan.method
    # shims for library drift only; the reference files are untouched
    epstan.method.timer = time.perf_counter
    _eigvalsh = linalg.eigvalsh

    def eigvalsh_compat(a, *args, eigvals=None, **kw):
        if eigvals is not None:
            kw['subset_by_index'] = eigvals
        return _eigvalsh(a, *args, **kw)
    epstan.method.linalg.eigvalsh = eigvalsh_compat
    return epstan


FakeFit, FakeModel = fakes.FakeFit, fakes.FakeModel


def gen_linalg(ep, out):
    u = ep.util
    # KAT of the reference's own test_scipy.py:22-31
    A = np.asfortranarray(np.eye(6) * 2.0)
    u.dpotri_routine(A, overwrite_c=True)
    out['kat_dpotri_2I6'] = A

    for tag, seed, d in (('d6', 11, 6), ('d20', 12, 20), ('d50', 13, 50)):
        rng = np.random.RandomState(seed)
        S = fakes.random_spd(rng, d)
        m = rng.standard_normal(d)
        Q, r = u.invert_normal_params(S, m)
        out['inv_%s_Q' % tag], out['inv_%s_r' % tag] = Q, r
        U = linalg.cholesky(S, lower=False)
        # cho_form with garbage below the diagonal (LAPACK ignores it)
        Ug = np.asfortranarray(U + np.tril(rng.standard_normal((d, d)), -1))
        Q2, r2 = u.invert_normal_params(Ug, m, cho_form=True)
        out['inv_%s_choform_Q' % tag], out['inv_%s_choform_r' % tag] = Q2, r2
        P = fakes.random_spd(rng, d, scale=3.0)
        n = 10 * d
        out['olse_%s_P' % tag] = u.olse(np.asfortranarray(S), n, P=np.asfortranarray(P))
        out['olse_%s_naive' % tag] = u.olse(np.asfortranarray(S), n)


def gen_worker(ep, out):
    """Worker.cavity + both moment branches of Worker.tilted (method.py:267-475)."""
    Worker = ep.method.Worker
    for tag, seed, d, C, it in (('d6', 21, 6, 4, 50), ('d20', 22, 20, 8, 200),
                                ('d50', 23, 50, 8, 200)):
        rng = np.random.RandomState(seed)
        K = 3
        Qs, rs = fakes.gaussian_site_factors(seed + 100, K, d)
        Q = np.asfortranarray(fakes.random_spd(rng, d, scale=4.0) + Qs.sum(axis=2) * 0.5)
        r = rng.standard_normal(d)
        for k in range(K):
            Qi = np.asfortranarray(0.5 * Qs[:, :, k])
            ri = 0.5 * rs[:, k]
            for mode in ('sample', 'olse'):
                model = FakeModel(Qs, rs)
                w = Worker(k, model, d, np.zeros((4, 2)), np.zeros(4),
                           A={'site_id': k}, chains=C, iter=it, prec_estim=mode)
                ok = w.cavity(Q, r, Qi, ri)
                assert ok
                out['wrk_%s_k%d_cavQ' % (tag, k)] = w.Mat.copy()
                out['wrk_%s_k%d_cavm' % (tag, k)] = w.vec.copy()
                dQi = np.zeros((d, d), order='F')
                dri = np.zeros(d)
                ok = w.tilted(dQi, dri, seed=seed * 7 + k)
                assert ok
                out['wrk_%s_k%d_%s_dQi' % (tag, k, mode)] = dQi
                out['wrk_%s_k%d_%s_dri' % (tag, k, mode)] = dri
                out['wrk_%s_k%d_stanseed' % (tag, k)] = np.array(model.seeds[0][1])
        # a non-PD cavity
        w = Worker(0, FakeModel(Qs, rs), d, np.zeros((4, 2)), np.zeros(4))
        out['wrk_%s_cav_notpd' % tag] = np.array(
            w.cavity(Q, r, np.asfortranarray(Q + np.eye(d)), r))


def gen_master(ep, out):
    """Master.run end to end against the exact Gaussian sampler (method.py:899-1247)."""
    Master = ep.method.Master
    import fit as ref_fit

    def build(K, d, seed, model_kw=None, **kw):
        Qs, rs = fakes.gaussian_site_factors(seed, K, d)
        model = FakeModel(Qs, rs, **(model_kw or {}))
        X = np.zeros((2 * K, 2))
        y = np.zeros(2 * K)
        m = Master(model, X, y, site_sizes=np.full(K, 2), dphi=d,
                   A_k={'site_id': list(range(K))}, **kw)
        return m, model

    def store(tag, res, master):
        info, (ms, Ss) = res[0], res[1]
        out[tag + '_info'] = np.array(info)
        out[tag + '_m'] = ms
        out[tag + '_S'] = Ss
        out[tag + '_Q'] = master.Q.copy()
        out[tag + '_r'] = master.r.copy()
        out[tag + '_Qi'] = master.Qi.copy()
        out[tag + '_ri'] = master.ri.copy()

    # A: plain run, default damping 1/K
    m, _ = build(5, 6, 31, chains=4, iter=100)
    store('runA', m.run(8, verbose=False, seed=5), m)
    # B: fit.py damping schedule, olse with two skipped iterations, prior given
    rng = np.random.RandomState(32)
    prior = {'Q': fakes.random_spd(rng, 8), 'r': rng.standard_normal(8)}
    m, _ = build(6, 8, 33, chains=4, iter=120, prior=prior, prec_estim='olse',
                 prec_estim_skip=2, df0=ref_fit.default_df0(6))
    store('runB', m.run(10, verbose=False, seed=6), m)
    # C: aggressive damping + noisy moments -> decay branch is exercised
    m, _ = build(8, 10, 34, chains=2, iter=40, df0=1.0)
    store('runC', m.run(6, verbose=False, seed=7), m)
    # C2: resume (state continuity, SURVEY 5 "checkpoint/resume")
    store('runC2', m.run(3, verbose=False, seed=8), m)
    # D: improper site -> df decays below the treshold -> forced pos.def.
    m, _ = build(4, 5, 35, model_kw={'constant': (0, 1)}, chains=4, iter=100, df0=0.5)
    m.iter = 1
    m.Qi[:, :, 0] = 6.0 * np.eye(5)
    m.Qi[:, :, 1] = -4.0 * np.eye(5)
    m.Q[:] = m.Q0 + m.Qi.sum(axis=2)
    for k, w in enumerate(m.workers):
        w.cavity(m.Q, m.r, m.Qi[:, :, k], m.ri[:, k])
        w.phase = 1          # site 0's cavity is improper; sampler ignores that
    store('runD', m.run(2, verbose=False, seed=9), m)
    # E: first-iteration global failure -> INFO_INVALID_PRIOR
    m, _ = build(4, 5, 36, model_kw={'inflate': 1e4}, chains=4, iter=100, df0=1.0)
    store('runE', m.run(2, verbose=False, seed=10), m)
    # F: degenerate draws -> every site fails -> INFO_ALL_SITES_FAIL
    m, _ = build(3, 4, 37, model_kw={'constant': True}, chains=4, iter=60)
    store('runF', m.run(2, verbose=False, seed=11), m)


def gen_cv(ep, out):
    """cv_moments on the recipe of epstan/test_cv.py:28-40,149-173 and config 2."""
    from scipy.stats import multivariate_normal
    u = ep.util
    for tag, seed, n, d in (('n60d4', 41, 60, 4), ('n800d20', 42, 800, 20)):
        rng = np.random.RandomState(seed)
        S1 = fakes.random_spd(rng, d)
        m1 = rng.standard_normal(d)
        S2 = S1 + 0.1 * fakes.random_spd(rng, d)
        m2 = m1 + 0.1 * rng.standard_normal(d)
        samp = m1 + rng.standard_normal((n, d)) @ linalg.cholesky(S1, lower=False)
        lp = multivariate_normal(mean=m1, cov=S1).logpdf(samp)
        Q2, r2 = u.invert_normal_params(S2, m2)
        for mcv in (True, False):
            S_hat, m_hat, used = u.cv_moments(samp.copy(), lp, Q2, r2, multiple_cv=mcv)
            t = '%s_%s' % (tag, 'multi' if mcv else 'single')
            out['cv_%s_S' % t], out['cv_%s_m' % t] = S_hat, m_hat
            out['cv_%s_used' % t] = np.array(used)
            _, _, _, a_S, a_m = u.cv_moments(samp.copy(), lp, Q2, r2, multiple_cv=mcv, ret_a=True)
            out['cv_%s_aS' % t], out['cv_%s_am' % t] = np.array(a_S), np.array(a_m)
        # treshold fallback: control variate far from the sample
        m3 = m1 + 5.0
        Q3, r3 = u.invert_normal_params(S2, m3)
        S_hat, m_hat, used = u.cv_moments(samp.copy(), lp, Q3, r3)
        out['cv_%s_fallback_S' % tag], out['cv_%s_fallback_m' % tag] = S_hat, m_hat
        out['cv_%s_fallback_used' % tag] = np.array(used)


def mix_case(seed=61):
    """Synthetic per-site draws for Master.mix_phi / mix_pred (tests regenerate them from the seed)."""
    rng = np.random.RandomState(seed)
    K, d, D = 3, 4, 2
    Jk = [2, 1, 3]
    n = [40, 50, 60]
    sites = []
    for k in range(K):
        sites.append(dict(phi=rng.standard_normal((45, d)) * 0.3 + rng.standard_normal(d),   # (one n: the device draw buffer)
                          beta=rng.standard_normal((n[k], D)) + 0.5 * k,
                          alpha=rng.standard_normal((n[k], Jk[k])) * 0.5 + k,
                          gamma=rng.standard_normal((n[k], 2)) + k))
    starts = np.concatenate(([0], np.cumsum(Jk)))
    smap_alpha = [slice(int(starts[k]), int(starts[k + 1])) for k in range(K)]
    smap_gamma = [np.array([0, 1]), np.array([1, 2]), np.array([2, 3])]        # overlapping: indices 1, 2 twice
    return sites, Jk, smap_alpha, smap_gamma


def gen_mix(ep, out):
    """The unmodified reference's Master.mix_phi / mix_pred (method.py:1250-1478) on mocked workers (the
    methods only read `workers[k].saved_samp` / `workers[k].fit`, which the reference never fills itself)."""
    import types
    sites, Jk, smap_alpha, smap_gamma = mix_case()

    class Fit(object):
        def __init__(self, samples):
            self.samples = samples
            self.model_pars = list(samples.keys())
            self.par_dims = [list(samples[p].shape[1:]) for p in self.model_pars]

        def extract(self, pars):
            return {pars: self.samples[pars].copy()}
    workers = [types.SimpleNamespace(fit=Fit({p: s[p] for p in ('beta', 'alpha', 'gamma')}),
                                     saved_samp={'phi': np.asfortranarray(s['phi'].copy())}) for s in sites]
    dummy = types.SimpleNamespace(iter=1, K=len(sites), dphi=sites[0]['phi'].shape[1], workers=workers)
    S, m = ep.method.Master.mix_phi(dummy)
    out['mix_phi_S'], out['mix_phi_m'] = S, m
    means, vars_ = ep.method.Master.mix_pred(dummy, ['beta', 'alpha', 'gamma'], [None, smap_alpha, smap_gamma],
                                             [None, (sum(Jk),), (4,)])
    for name, mm, vv in zip(('beta', 'alpha', 'gamma'), means, vars_):
        out['mix_%s_m' % name], out['mix_%s_v' % name] = mm, vv
    # experiment/fit.py:763-815 `_create_pmaps`: which entries of a parameter each site fills, recorded as the
    # flat indices every site's map selects
    import fit as ref_fit
    for tag, J, K, Ns in (('KeqJ', 6, 6, None), ('KltJ', 6, 3, [2, 1, 3])):
        phiers, shapes = (0, None, 0), ((J,), (2,), (J, 2))
        pmaps = ref_fit._create_pmaps(phiers, J, K, Ns)
        for ip, (pm, shp) in enumerate(zip(pmaps, shapes)):
            if pm is None:
                continue
            arr = np.arange(int(np.prod(shp))).reshape(shp)
            sel = [np.atleast_1d(arr[pm[k]]).ravel() for k in range(K)]
            out['pmap_%s_%d_idx' % (tag, ip)] = np.concatenate(sel)
            out['pmap_%s_%d_len' % (tag, ip)] = np.array([len(v) for v in sel])


def gen_misc(ep, out):
    import fit as ref_fit
    import find_damp as ref_fd
    for K in (2, 4, 32, 64):
        f = ref_fit.default_df0(K)
        out['df0_K%d' % K] = np.array([f(i) for i in range(1, 41)])
    rng = np.random.RandomState(51)
    for d in (3, 20):
        S0, S1 = fakes.random_spd(rng, d), fakes.random_spd(rng, d)
        m0, m1 = rng.standard_normal(d), rng.standard_normal(d)
        out['kl_d%d' % d] = np.array(ref_fd.kl_mvn(m0, S0, m1, S1))
    for tag, J, K, Nj in (('const', 64, 4, 20), ('const32', 64, 32, 20),
                          ('ragged', 40, 7, np.random.RandomState(52).randint(5, 40, size=40)),
                          ('ragged2', 33, 32, np.random.RandomState(53).randint(1, 9, size=33))):
        Nk, Nj_k, j_ind_k = ep.util.distribute_groups(J, K, Nj)
        out['dg_%s_Nk' % tag], out['dg_%s_Njk' % tag] = Nk, Nj_k
        out['dg_%s_jind' % tag] = j_ind_k


def gen_models(ep, out):
    """experiment/models/m{1..5}b.py simulators and priors (small shapes)."""
    import importlib
    for name in ('m1b', 'm2b', 'm3b', 'm4b', 'm5b'):
        mod = importlib.import_module('models.' + name)
        for tag, kw, npg in (('corr', dict(Sigma_x='rand'), 5), ('iid', dict(), [3, 8])):
            mdl = mod.model(6, 3, npg)
            dat = mdl.simulate_data(rng=100, **kw)
            key = 'mdl_%s_%s_' % (name, tag)
            out[key + 'X'], out[key + 'y'], out[key + 'Nj'] = dat.X, dat.y, dat.Nj
            out[key + 'phi'] = dat.true_values['phi']
            out[key + 'unc'] = np.append(dat.calc_uncertainty()[0], dat.calc_uncertainty()[1])
        S0, m0, Q0, r0 = mod.model(6, 3, 5).get_prior()
        out['mdl_%s_Q0' % name], out['mdl_%s_r0' % name] = Q0, r0


def main():
    ep = import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in (('linalg', gen_linalg), ('worker', gen_worker),
                     ('master', gen_master), ('cv', gen_cv), ('mix', gen_mix), ('misc', gen_misc),
                     ('models', gen_models)):
        out = {}
        fn(ep, out)
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **out)
        print('%-8s %3d arrays  %7.1f kB' % (name, len(out), os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
