"""experiment/fit.py and find_damp.py on the GPU (reference config 1 shape,
scaled down so the test finishes in seconds)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EXP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ep-stan_b200', 'experiment')


@pytest.fixture(scope='module')
def fit():
    if EXP not in sys.path:
        sys.path.insert(0, EXP)
    import fit as fit_mod
    return fit_mod


def test_fit_cli_options(fit):
    name, conf = fit.parse_args(['m1b', '--run_ep', '1', '--K', '4', '--npg', '10', '30', '--damp', '0.5'])
    assert name == 'm1b' and conf.run_ep is True and conf.K == 4 and conf.npg == [10, 30] and conf.damp == 0.5
    assert conf.J == 64 and conf.D == 16 and conf.siter == 200 and conf.chains == 4 and conf.seed_data == 100
    with pytest.raises(ValueError):
        fit.configurations(bogus=1)
    assert abs(fit.default_df0(32)(1) - 0.5) < 1e-15 and fit.EP_DEFAULT_ITERS_TO_RUN(4) == 20


@pytest.mark.parametrize('model,K', [('m1b', 4), ('m4b', 8), ('m3b', 2), ('m2b', 4), ('m5b', 8)])
def test_fit_main_ep(fit, tmp_path, monkeypatch, model, K):
    """K < J (grouped sites, `<model>.stan` density) and K == J (`_sg` density)."""
    monkeypatch.setattr(fit, 'RES_PATH', str(tmp_path))
    conf = fit.configurations(J=8, D=3, K=K, npg=25, run_ep=True, iter=3, siter=100, chains=4)
    fit.main(model, conf)
    res = np.load(os.path.join(str(tmp_path), 'res_d_%s.npz' % model), allow_pickle=True)
    d = {'m1b': 4, 'm2b': 2, 'm3b': 4, 'm4b': 8, 'm5b': 8}[model]
    assert res['m_s_ep'].shape == (4, d) and res['S_s_ep'].shape == (4, d, d)
    assert res['time_s_ep'].shape == (4,) and np.isnan(res['mrhat_s_ep'][0])
    assert np.all(np.isfinite(res['m_s_ep'])) and np.all(np.diff(res['time_s_ep']) > 0)
    tv = np.load(os.path.join(str(tmp_path), 'true_vals_%s.npz' % model), allow_pickle=True)
    assert tv['phi'].shape == (d,)


def test_target_and_find_damp(fit, tmp_path, monkeypatch):
    monkeypatch.setattr(fit, 'RES_PATH', str(tmp_path))
    import find_damp
    monkeypatch.setattr(find_damp, 'RES_PATH', str(tmp_path))
    monkeypatch.setattr(find_damp, 'SITER', 100)
    conf = fit.configurations(J=6, D=2, K=6, npg=30, run_target=True, target_siter=600, save_true=False)
    fit.main('m1b', conf)
    tgt = np.load(os.path.join(str(tmp_path), 'target_m1b.npz'), allow_pickle=True)
    assert tgt['m_target'].shape == (3,) and np.all(np.linalg.eigvalsh(tgt['S_target']) > 0)
    out = find_damp.main('m1b', K=6, iters=3, conf_overrides=dict(npg=30))
    assert out['kls'].shape == (3, 31) and np.isfinite(out['kls']).any()
    assert np.all(np.isfinite(out['kls_selected'])) and np.all(np.isfinite(out['damps_selected']))
    # EP moves the approximation towards the full-data target
    assert out['kls_selected'][-1] < out['kls_selected'][0]


@pytest.mark.parametrize('K', [6, 3])
def test_fit_consensus_mc(fit, tmp_path, monkeypatch, K):
    """fit.py --run_consensus (reference fit.py:539-675): per-site sampling against the fractionated
    prior, pooled moments, `res_c_*.npz` schema; K == J appends the longer run."""
    monkeypatch.setattr(fit, 'RES_PATH', str(tmp_path))
    monkeypatch.setattr(fit, 'CONS_ITERS', [40, 80])
    conf = fit.configurations(J=6, D=2, K=K, npg=30, run_consensus=True, chains=4)
    fit.main('m1b', conf)
    res = np.load(os.path.join(str(tmp_path), 'res_c_m1b.npz'), allow_pickle=True)
    n_it = 3 if K == 6 else 2
    assert res['m_s_cons'].shape == (n_it, 3) and res['S_s_cons'].shape == (n_it, 3, 3)
    assert np.all(np.isfinite(res['m_s_cons'])) and np.all(res['time_s_cons'] > 0)
    assert np.all(res['mstepsize_s_cons'] > 0) and np.all(res['mrhat_s_cons'] > 0.9)
    for S in res['S_s_cons']:
        assert np.all(np.linalg.eigvalsh(S) > 0)
