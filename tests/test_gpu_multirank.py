"""Real-GPU, multi-rank (NCCL) runs of the EP loop against the single-rank result.

Needs >= 2 visible GPUs (skipped otherwise; the gloo tests in test_distributed_gloo.py cover the host
logic on CPU).  Sampling is deterministic per site, so sharding only changes the order of the site sum:
results must agree to ~1e-9."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _launch(tmp_path, nproc, mode, K, port, env_extra=None):
    out = str(tmp_path / ('mr_%d_%s.npz' % (nproc, mode)))
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(HERE, 'mr_worker.py'), out, mode, str(K)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(out)


@pytest.mark.parametrize('mode', ['current', 'side', 'private'])
def test_nccl_ranks_match_single_rank(tmp_path, mode):
    n = _ngpu()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    import mr_worker
    K = 10
    m, info, ms, Ss = mr_worker.run(K)
    assert info == 0
    for nproc in sorted(set([2, min(n, 4)])):
        res = _launch(tmp_path, nproc, mode, K, 29600 + nproc)
        assert int(res['info']) == 0 and int(res['size']) == nproc
        for a, b in ((res['ms'], ms), (res['Ss'], Ss), (res['Q'], m.Q), (res['Qi'], m.Qi), (res['cavm'], m._cavm)):
            assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b))), (mode, nproc)


def test_nccl_selection_and_uneven_shards(tmp_path):
    """automatic damping selection (its own all-reduce) with K not divisible by the rank count"""
    n = _ngpu()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    import mr_worker
    K = 11
    m, info, ms, Ss = mr_worker.run(K, df_select='snr')
    nproc = min(n, 4)
    res = _launch(tmp_path, nproc, 'current', K, 29650, {'MR_SELECT': '1'})
    assert int(res['info']) == info == 0
    assert np.allclose(res['df'], m.history['df'], rtol=1e-9)
    assert np.max(np.abs(res['ms'] - ms)) <= 1e-9 * max(1.0, np.max(np.abs(ms)))
    assert np.max(np.abs(res['Ss'] - Ss)) <= 1e-9 * max(1.0, np.max(np.abs(Ss)))
