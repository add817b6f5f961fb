"""Damping-factor sweep for the distributed EP algorithm, B200 edition of the
reference's experiment/find_damp.py.

    $ python find_damp.py <model_name> [K] [iters]

Per EP iteration, after the tilted step of all sites, N_DAMP damping values in
(0,1) are tried: for each the global approximation is rebuilt, it and all K
cavities must be pos.def., and the result is scored against a saved target
(results/target_<model>.npz, written by ``fit.py <model> --run_target 1``) by
MSE of the mean and KL(target || approximation)  (reference find_damp.py:144-174,
kl_mvn :32-51).  The whole sweep is ONE batched device call (31 x K Cholesky
factorisations) instead of 31 x K LAPACK calls in Python loops.  Then the
preselected ``fit.default_df0`` is applied with the reference's decay rule
(:186-236).  Results: results/find_damp_K<K>.npz with the reference's keys
(``lls*`` need the target draws and are left NaN unless target_samp exists).
"""

import os
import sys

import numpy as np
from scipy import stats

CUR_PATH = os.path.dirname(os.path.abspath(__file__))
PARENT_PATH = os.path.abspath(os.path.join(CUR_PATH, os.pardir))
RES_PATH = os.path.join(CUR_PATH, 'results')
for _p in (CUR_PATH, PARENT_PATH):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import fit                                   # noqa: E402
from epstan.method import Master             # noqa: E402
from epstan import _lib                      # noqa: E402

CHAINS = 8
SITER = 200
N_DAMP = 31


def kl_mvn(m0, S0, m1, S1, sum_log_diag_cho_S0=None):
    """KL(N(m0,S0) || N(m1,S1)) (reference find_damp.py:32-51); host helper for
    single evaluations -- the sweep computes the same quantity on the device."""
    L1 = np.linalg.cholesky(S1)
    if sum_log_diag_cho_S0 is None:
        sum_log_diag_cho_S0 = np.sum(np.log(np.diag(np.linalg.cholesky(S0))))
    dm = m1 - m0
    sol = np.linalg.solve(S1, S0)
    return (0.5 * (np.trace(sol) + dm.dot(np.linalg.solve(S1, dm)) - len(m0))
            - sum_log_diag_cho_S0 + np.sum(np.log(np.diag(L1))))


def main(model_name, K=None, iters=None, target=None, conf_overrides=None):
    if target is None:
        tf = np.load(os.path.join(RES_PATH, 'target_{}.npz'.format(model_name)), allow_pickle=True)
        m_target, S_target, tconf = tf['m_target'], tf['S_target'], tf['conf'][()]
        tf.close()
    else:
        m_target, S_target, tconf = target
    samp_path = os.path.join(RES_PATH, 'target_samp_{}.npz'.format(model_name))
    samp_target = np.load(samp_path)['samp_target'] if os.path.exists(samp_path) else None
    J, D = tconf['J'], tconf['D']
    K = J if K is None else K
    iters = fit.EP_DEFAULT_ITERS_TO_RUN(K) if iters is None else iters
    kw = dict(J=J, D=D, K=K, chains=CHAINS, siter=SITER, save_true=False)
    kw.update(conf_overrides or {})
    master = fit.main(model_name, fit.configurations(**kw), ret_master=True)
    ctx = master._shard.ctx
    df0 = fit.default_df0(K)
    decay = Master.DEFAULT_KWARGS['df_decay']
    treshold = Master.DEFAULT_KWARGS['df_treshold']

    damps = np.linspace(0, 1, N_DAMP + 2)[1:-1]
    mses = np.full((iters, N_DAMP), np.nan)
    lls = np.full((iters, N_DAMP), np.nan)
    kls = np.full((iters, N_DAMP), np.nan)
    damps_selected = np.full(iters, np.nan)
    mses_selected = np.full(iters + 1, np.nan)
    lls_selected = np.full(iters + 1, np.nan)
    kls_selected = np.full(iters + 1, np.nan)

    def score(m, S):
        ll = np.nan
        if samp_target is not None:
            ll = np.sum(stats.multivariate_normal.logpdf(samp_target, mean=m, cov=S))
        return np.mean((m - m_target) ** 2), ll, kl_mvn(m_target, S_target, m, S)

    init_S, init_m = master.cur_approx()
    mses_selected[0], lls_selected[0], kls_selected[0] = score(init_m, init_S.T)

    master._push_state(with_cavity=True)
    rng = np.random.RandomState()
    d = master.dphi
    m_buf, S_buf = np.empty(d), np.empty((d, d), order='F')
    for iter_ind in range(iters):
        curiter = iter_ind + 1
        print("Iteration {}/{}".format(curiter, iters))
        oks = master._tilted_all(rng.randint(0, 2 ** 31 - 1, size=master.K), None)
        if not np.all(oks):
            print("    Tilted fails at {}".format(int(np.nonzero(~oks)[0][0]) + 1))
            break
        # all N_DAMP candidates in one device call
        mses[iter_ind], kls[iter_ind] = ctx.damp_sweep(damps, m_target, S_target)
        # preselected damping with the decay rule
        df = df0(curiter)
        while True:
            damps_selected[iter_ind] = df
            ctx.update_partial(df)
            ok = ctx.update_finish()
            if ok:
                ctx.global_moments(m_buf, S_buf)
                ok = ctx.cavity(proposal=True)[1]
            if ok:
                mses_selected[iter_ind + 1], lls_selected[iter_ind + 1], kls_selected[iter_ind + 1] = \
                    score(m_buf, S_buf.T)
                break
            print('    lowering-df')
            df *= decay
            damps_selected[iter_ind] = df
            if df < treshold:
                print('    df_threshold reached')
                break
        ctx.accept()
        for w in master.workers:
            w.phase = 1
    master.sync_host()
    os.makedirs(RES_PATH, exist_ok=True)
    out = dict(damps=damps, mses=mses, lls=lls, kls=kls, damps_selected=damps_selected,
               mses_selected=mses_selected, lls_selected=lls_selected, kls_selected=kls_selected)
    np.savez(os.path.join(RES_PATH, 'find_damp_K{}.npz'.format(K)), **out)
    return out


if __name__ == '__main__':
    kwargs = {}
    if len(sys.argv) > 2:
        kwargs['K'] = int(sys.argv[2])
    if len(sys.argv) > 3:
        kwargs['iters'] = int(sys.argv[3])
    main(sys.argv[1], **kwargs)
