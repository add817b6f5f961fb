"""A simulated experiment for the distributed EP algorithm (arXiv:1412.4869),
B200 edition of the reference's experiment/fit.py.

Execute with:
    $ python fit.py <model_name> [<optional arguments>]

Same command line as the reference (fit.py:972-1004): all 26 option names of
CONFS with the defaults of CONF_DEFAULT, e.g.
    $ python fit.py m1b --run_ep 1 --K 4
Group ``<model_name>`` is a simulated model in ./models (m1b, m2b, m3b, m4b, m5b).  The
EP branch (``--run_ep``) is the reference's (fit.py:279-459) on the GPU Master.
``--run_full`` / ``--run_target`` sample the full-data posterior with the same
built-in NUTS sampler (one site holding every group, cavity = prior);
``--mix`` pools the last tilted draws (Master.mix_phi / mix_pred); ``--run_consensus`` pools per-site draws on one GPU.

Results go to ./results with the reference's file names and npz keys
(res_d_<model>.npz: m_s_ep, S_s_ep, time_s_ep, mstepsize_s_ep, mrhat_s_ep;
true_vals_<model>.npz; res_f_<model>.npz; target_<model>.npz).
"""

import argparse
import os
import sys

import numpy as np

CUR_PATH = os.path.dirname(os.path.abspath(__file__))
PARENT_PATH = os.path.abspath(os.path.join(CUR_PATH, os.pardir))
RES_PATH = os.path.join(CUR_PATH, 'results')
MOD_PATH = os.path.join(CUR_PATH, 'models')
for _p in (CUR_PATH, PARENT_PATH):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from epstan.method import Master, Worker          # noqa: E402
from epstan.util import distribute_groups         # noqa: E402

CONFS = [
    'J', 'D', 'npg', 'cor_input',
    'run_all', 'run_ep', 'run_full', 'run_consensus', 'run_target',
    'iter', 'siter', 'target_siter', 'chains',
    'K', 'damp', 'mix', 'prec_estim',
    'seed_data', 'seed_ep', 'seed_full', 'seed_cons', 'seed_target',
    'id', 'save_true', 'save_res', 'save_target_samp',
]

CONF_DEFAULT = dict(
    J=64, D=16, K=32, npg=20, cor_input=True,
    run_all=False, run_ep=False, run_full=False, run_consensus=False, run_target=False,
    iter=None, siter=200, target_siter=10000, chains=4,
    damp=None, mix=False, prec_estim='sample',
    seed_data=100, seed_ep=1, seed_full=2, seed_cons=3, seed_target=4,
    id=None, save_true=True, save_res=True, save_target_samp=False,
)

FULL_ITERS = [50, 100, 200, 300, 400, 600, 800, 1000, 1200, 1600, 2000, 3200]
CONS_ITERS = [50, 100, 500, 1000, 2000, 4000]


def EP_DEFAULT_ITERS_TO_RUN(K):
    return int(max(4 * K, 20))


DAMP_DECAY_AT_K = 0.9


def DAMP_START(K):
    return 0.5


def DAMP_END(K):
    return min(1 / K, 0.2)


def default_df0(K):
    """Default damping factor function: exponential decay from DAMP_START to
    DAMP_END, 90 % of the way at iteration K (reference fit.py:180-186)."""
    rate = -np.log(1 - DAMP_DECAY_AT_K) / (K - 1)
    lo = DAMP_END(K)
    span = DAMP_START(K) - lo
    return lambda curiter: span * np.exp(-rate * (curiter - 1)) + lo


class configurations(object):
    """Configuration container for the function main."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            if k not in CONF_DEFAULT:
                raise ValueError("Invalid option `{}`".format(k))
            setattr(self, k, v)
        for k, v in CONF_DEFAULT.items():
            if k not in kwargs:
                setattr(self, k, v)

    def __str__(self):
        return '\n'.join('{!s} = {!r}'.format(opt, self.__dict__[opt]) for opt in CONFS if opt in self.__dict__)

    __repr__ = __str__


def _res_file(stem, model_name, conf):
    os.makedirs(RES_PATH, exist_ok=True)
    name = '{}_{}_{}.npz'.format(stem, model_name, conf.id) if conf.id else '{}_{}.npz'.format(stem, model_name)
    return os.path.join(RES_PATH, name)


def _full_posterior_draws(model_name, data, prior, chains, siter, seed):
    """Draws of phi from the full-data posterior: one 'site' that holds every
    group, whose cavity is the prior, sampled by the built-in GPU NUTS."""
    dphi = prior['Q'].shape[0]
    w = Worker(0, os.path.join(MOD_PATH, model_name), dphi, data.X, data.y,
               A={'J': data.J, 'j_ind': data.j_ind + 1}, chains=chains, iter=siter)
    zero_Q = np.zeros((dphi, dphi), order='F')
    if not w.cavity(np.asfortranarray(prior['Q']), prior['r'], zero_Q, np.zeros(dphi)):
        raise RuntimeError("prior is not pos.def.")
    dQ = np.zeros((dphi, dphi), order='F')
    dr = np.zeros(dphi)
    w.tilted(dQ, dr, save_samples=('phi',), seed=seed)
    return w.saved_samp['phi'], w


def _param_maps(phiers, J, K, Nj_k):
    """Which entries of each inferred parameter a site contributes to (reference fit.py:763-815,
    `_create_pmaps`): None for a shared parameter; for a parameter with a hierarchical dimension `ih`, per
    site an index into that dimension -- the site's own group (K == J) or its block of merged groups."""
    maps = []
    for ih in phiers:
        if ih is None:
            maps.append(None)
        elif K == J:
            maps.append(np.arange(K) if ih == 0 else
                        [tuple(k if a == ih else slice(None) for a in range(ih + 1)) for k in range(K)])
        else:
            starts = np.concatenate(([0], np.cumsum(Nj_k)))
            blocks = [slice(int(starts[k]), int(starts[k + 1])) for k in range(K)]
            maps.append(blocks if ih == 0 else
                        [tuple(b if a == ih else slice(None) for a in range(ih + 1)) for b in blocks])
    return maps


def _site_master(model_name, data, J, K, options):
    """Master over K sites built from the J simulated groups (reference fit.py:300-345)."""
    if K < 2:
        raise ValueError("K should be at least 2.")
    elif K < J:
        Nk, Nj_k, j_ind_k = distribute_groups(J, K, data.Nj)
        return Master(os.path.join(MOD_PATH, model_name), data.X, data.y,
                      A_k={'J': Nj_k}, A_n={'j_ind': j_ind_k + 1}, site_sizes=Nk, **options)
    elif K == J:
        return Master(os.path.join(MOD_PATH, model_name + '_sg'), data.X, data.y,
                      site_sizes=data.Nj, **options)
    elif K <= data.N:
        raise NotImplementedError("Splitting the groups not implemented.")
    raise ValueError("K cant be greater than number of samples")


def _consensus_mc(model_name, data, prior, J, K, conf, iters_list):
    """Consensus MC baseline (reference fit.py:539-675): every site is sampled on its own against
    the fractionated prior N(Q0/K, r0/K) and the draws of all sites are pooled.  The sites are the
    ones of the EP run, so the batched GPU sampler is used as is: the 'cavity' of every site is set
    to the fractionated prior."""
    from scipy import linalg
    from epstan import _lib
    from epstan.method import MAX_UINT
    master = _site_master(model_name, data, J, K, dict(prior=prior, chains=conf.chains, iter=conf.siter))
    sh = master._shard
    if sh.comm.size > 1:
        raise NotImplementedError("consensus MC runs on one GPU (it pools the draws of all sites)")
    ctx, d = sh.ctx, master.dphi
    Qk, rk = prior['Q'] / K, prior['r'] / K
    mk = linalg.cho_solve(linalg.cho_factor(Qk), rk)
    ctx.upload(_lib.CAVQ, np.asfortranarray(np.repeat(Qk[:, :, None], K, axis=2)))
    ctx.upload(_lib.CAVM, np.asfortranarray(np.repeat(mk[:, None], K, axis=1)))
    seeds = np.random.RandomState(seed=conf.seed_cons).randint(0, MAX_UINT, size=K)   # constant over the runs
    n_it = len(iters_list)
    out = dict(m_s_cons=np.full((n_it, d), np.nan), S_s_cons=np.full((n_it, d, d), np.nan),
               time_s_cons=np.full(n_it, np.nan), mstepsize_s_cons=np.full(n_it, np.nan),
               mrhat_s_cons=np.full(n_it, np.nan))
    for i, iters in enumerate(iters_list):
        iters = int(iters)
        print('  iter {}: {}'.format(i + 1, iters))
        msteps, mrhat, _, secs = ctx.tilted_sample(seeds, conf.chains, iters)
        n = conf.chains * (iters - iters // 2)
        samp = ctx.get_draws(n).transpose(0, 2, 1).reshape(-1, d)        # all sites' draws of phi
        out['m_s_cons'][i] = samp.mean(axis=0)
        samp = samp - out['m_s_cons'][i]
        out['S_s_cons'][i] = samp.T.dot(samp) / (samp.shape[0] - 1)
        out['time_s_cons'][i] = secs
        out['mstepsize_s_cons'][i] = np.mean(msteps)
        out['mrhat_s_cons'][i] = np.max(mrhat)
    return out


def main(model_name, conf, ret_master=False):
    """Fit requested model with given configurations (reference fit.py:210-760).
    ``ret_master`` returns the epstan Master before running it."""
    if not isinstance(conf, configurations):
        raise ValueError("Invalid arg. `conf`, use class fit.configurations")
    print("Configurations:")
    print('    ' + str(conf).replace('\n', '\n    '))
    J, D, K = conf.J, conf.D, conf.K

    model_module = getattr(__import__('models.' + model_name), model_name)
    model = model_module.model(J, D, conf.npg)
    data = model.simulate_data(Sigma_x='rand', rng=conf.seed_data) if conf.cor_input \
        else model.simulate_data(rng=conf.seed_data)
    uncertainty_global, uncertainty_group = data.calc_uncertainty()
    S0, m0, Q0, r0 = model.get_prior()
    prior = {'Q': Q0, 'r': r0}
    pnames, pshapes, phiers = model.get_param_definitions()

    if conf.save_true:
        np.savez(_res_file('true_vals', model_name, conf), J=J, D=D, npg=conf.npg, seed=conf.seed_data,
                 pnames=pnames, uncertainty_global=uncertainty_global, uncertainty_group=uncertainty_group,
                 X_param=data.X_param, **data.true_values)
        print("True values saved into results")

    # ------------------------------------------------------------ distributed EP
    if conf.run_ep or conf.run_all or ret_master:
        print("Distributed method")
        iters_to_run = EP_DEFAULT_ITERS_TO_RUN(K) if conf.iter is None else conf.iter
        df0 = default_df0(K) if conf.damp is None else conf.damp
        epstan_options = dict(prior=prior, prec_estim=conf.prec_estim, df0=df0, init_site=None,
                              chains=conf.chains, iter=conf.siter, warmup=None, thin=1)
        epstan_master = _site_master(model_name, data, J, K, epstan_options)
        pmaps = _param_maps(phiers, J, K, None if K == J else distribute_groups(J, K, data.Nj)[1])
        if ret_master:
            print("Returning epstan.Master")
            return epstan_master

        S_ep_init, m_ep_init = epstan_master.cur_approx()
        print("Run distributed EP algorithm for {} iterations.".format(iters_to_run))
        info, (m_s_ep, S_s_ep), (time_s_ep, mstepsize_s_ep, mrhat_s_ep, othertimes) = epstan_master.run(
            iters_to_run, return_analytics=True, save_last_param=pnames if conf.mix else None, seed=conf.seed_ep)
        time_s_ep = np.insert(time_s_ep.cumsum(), 0, 0.0)
        S_s_ep = np.concatenate((S_ep_init[None, :, :], S_s_ep), axis=0)
        m_s_ep = np.concatenate((m_ep_init[None, :], m_s_ep), axis=0)
        mstepsize_s_ep = np.insert(mstepsize_s_ep, 0, np.nan)
        mrhat_s_ep = np.insert(mrhat_s_ep, 0, np.nan)
        if info:
            if conf.save_res:
                np.savez(_res_file('res_d', model_name, conf), conf=conf.__dict__, m_s_ep=m_s_ep, S_s_ep=S_s_ep,
                         time_s_ep=time_s_ep, mstepsize_s_ep=mstepsize_s_ep, mrhat_s_ep=mrhat_s_ep,
                         othertimes=othertimes, last_iter=epstan_master.iter)
                print("Uncomplete distributed model results saved.")
            raise RuntimeError('epstan algorithm failed with error code: {}'.format(info))
        extra = {}
        if conf.mix:
            # final approximation by mixing the last samples of all the sites (reference fit.py:408-420)
            print("Form the final approximation by mixing the last samples from all the sites.")
            S_mix, m_mix = epstan_master.mix_phi()
            pms, pvars = epstan_master.mix_pred(list(pnames), pmaps, list(pshapes))
            extra = dict(othertimes=othertimes, m_phi_ep=m_mix, S_phi_ep=S_mix)
            for name, pm, pv in zip(pnames, pms, pvars):
                extra['m_' + name + '_ep'] = pm
                extra['v_' + name + '_ep'] = pv
        if conf.save_res:
            np.savez(_res_file('res_d', model_name, conf), conf=conf.__dict__, m_s_ep=m_s_ep, S_s_ep=S_s_ep,
                     time_s_ep=time_s_ep, mstepsize_s_ep=mstepsize_s_ep, mrhat_s_ep=mrhat_s_ep, **extra)
            print("Distributed model results saved.")
        del epstan_master
        print("Done with distributed method")

    # ------------------------------------------------------- full model sampling
    if conf.run_full or conf.run_all:
        print("Full model")
        m_s, S_s, t_s = [], [], []
        for siter in FULL_ITERS:
            samp, w = _full_posterior_draws(model_name, data, prior, conf.chains, siter, conf.seed_full)
            m_s.append(samp.mean(axis=0))
            S_s.append(np.cov(samp, rowvar=False))
            t_s.append(w.last_time)
        if conf.save_res:
            np.savez(_res_file('res_f', model_name, conf), conf=conf.__dict__, m_s_full=np.array(m_s),
                     S_s_full=np.array(S_s), time_s_full=np.array(t_s))
            print("Full model results saved.")

    # --------------------------------------------------------------- consensus MC
    if conf.run_consensus or conf.run_all:
        print("Consensus MC")
        iters_list = list(CONS_ITERS) + ([CONS_ITERS[-1] * 1.7] if K == J else [])   # "run additionally a bit longer"
        res_c = _consensus_mc(model_name, data, prior, J, K, conf, iters_list)
        if conf.save_res:
            np.savez(_res_file('res_c', model_name, conf), conf=conf.__dict__, **res_c)
            print("Consensus MC results saved.")
        print("Done with consensus MC")

    # --------------------------------------------------------- target approximation
    if conf.run_target or conf.run_all:
        print("Target approximation")
        samp, w = _full_posterior_draws(model_name, data, prior, conf.chains, conf.target_siter, conf.seed_target)
        m_target = samp.mean(axis=0)
        S_target = np.cov(samp, rowvar=False)
        if conf.save_res:
            np.savez(_res_file('target', model_name, conf), conf=conf.__dict__, m_target=m_target,
                     S_target=S_target)
            if conf.save_target_samp:
                np.savez(_res_file('target_samp', model_name, conf), samp_target=samp)
            print("Target results saved.")


# ==============================================================================
# Command line argument parsing (reference fit.py:856-1004)
# ==============================================================================

def _parse_bool(arg):
    up = str(arg).upper()
    if up == 'TRUE'[:len(up)] or up == '1':
        return True
    if up == 'FALSE'[:len(up)] or up == '0':
        return False
    raise ValueError("Invalid boolean option")


def _parse_positive_int(arg):
    if arg.isalnum() and int(arg) > 0:
        return int(arg)
    raise ValueError("Invalid integer option")


def _parse_nonnegative_int(arg):
    if arg.isalnum():
        return int(arg)
    raise ValueError("Invalid integer option")


def _parse_damp(arg):
    f = float(arg)
    if f <= 0.0 or f > 1.0:
        raise ValueError("Invalid damp option")
    return f


_B, _P, _N = dict(type=_parse_bool, metavar='B'), dict(type=_parse_positive_int, metavar='P'), \
    dict(type=_parse_nonnegative_int, metavar='N')
CONF_CUSTOMS = dict(
    J=_P, D=_P, K=_P, npg=dict(nargs='+', type=_parse_positive_int, metavar='P'), cor_input=_B,
    run_all=_B, run_ep=_B, run_full=_B, run_consensus=_B, run_target=_B,
    iter=_P, siter=_P, target_siter=_P, chains=_P,
    damp=dict(type=_parse_damp, metavar='F'), mix=_B, prec_estim=dict(metavar='S'),
    seed_data=_N, seed_ep=_N, seed_full=_N, seed_cons=_N, seed_target=_N,
    id=dict(metavar='S'), save_true=_B, save_res=_B, save_target_samp=_B,
)

CONF_HELP = dict(
    J='number of hierarchical groups', D='number of inputs', K='number of sites',
    npg='number of observations per group (constant or min max)', cor_input='correlated input variable',
    run_all='run all the methods', run_ep='run the distributed EP method', run_full='run the full model method',
    run_consensus='run consensus MC method', run_target='run target approximation',
    iter='number of distributed EP iterations', siter='sampler iterations in each major iteration',
    target_siter='sampler iterations for the target approximation', chains='number of chains used in sampling',
    damp='damping factor constant', mix='mix last iteration samples',
    prec_estim='estimate method for tilted distribution precision matrix: sample or olse',
    seed_data='seed for data simulation', seed_ep='seed for distributed EP sampling',
    seed_full='seed for full sampling', seed_cons='seed for consensus sampling',
    seed_target='seed for target sampling', id='optional id appended to the end of the result files',
    save_true='save true values', save_res='save results', save_target_samp='save target approximation samples',
)


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split('\n\n', 1)[0],
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('model_name', help="name of the model")
    for opt in CONFS:
        parser.add_argument('--' + opt, default=CONF_DEFAULT[opt],
                            help='{}, default {}'.format(CONF_HELP[opt], CONF_DEFAULT[opt]), **CONF_CUSTOMS[opt])
    args = vars(parser.parse_args(argv))
    model_name = args.pop('model_name')
    if isinstance(args['npg'], list):
        if len(args['npg']) == 1:
            args['npg'] = args['npg'][0]
        elif len(args['npg']) > 2:
            raise ValueError("Invalid arg `npg`, provide one or two elements")
    return model_name, configurations(**args)


if __name__ == '__main__':
    main(*parse_args())
