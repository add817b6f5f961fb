"""The N>1 path on CPU: two gloo ranks, each owning half of the sites, must
reproduce the single-process reference results (golden Master.run scenarios).
The device context is replaced by the oracle-backed double in
tests/fake_backend.py -- what is under test is the host logic of the sharded
Master: contiguous site shards, ONE all-reduce of [sum Qi2 | sum ri2] per update
attempt, consensus on the pos.def. flags, gathering of the mirrors."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.join(ROOT, 'tests')


def _worker(rank, world, port, tag, ret):
    for p in (ROOT, os.path.join(ROOT, 'ep-stan_b200'), TESTS):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import epstan.method as method
        import fake_backend
        import golden_inputs as gi
        from test_gpu_linalg import build_master
        method.Master._context_factory = staticmethod(lambda dev, stream: fake_backend.OracleContext(dev, stream))

        class _Ep(object):
            pass
        ep = _Ep()
        ep.method = method
        scs = gi.master_scenarios()
        sc = scs[tag]
        m = build_master(ep, sc)
        assert m.comm.size == world and m._shard.n_local in (sc['K'] // world, sc['K'] // world + 1)
        info, (ms, Ss) = m.run(sc['niter'], verbose=False, seed=sc['seed'])
        out = dict(info=info, m=ms, S=Ss, Q=m.Q.copy(), Qi=m.Qi.copy(), ri=m.ri.copy(),
                   k=(m._shard.k_begin, m._shard.k_end))
        if tag == 'runC':
            sc2 = scs['runC2']
            res = m.run(sc2['niter'], verbose=False, seed=sc2['seed'])
            out.update(info2=res[0], m2=res[1][0], S2=res[1][1])
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def _worker_snr(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, 'ep-stan_b200'), TESTS):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import epstan.method as method
        import fake_backend
        import test_damping as td
        from oracle import ep_linalg as orc
        method.Master._context_factory = staticmethod(lambda dev, stream: fake_backend.OracleContext(dev, stream))
        sc = td._scenario(K=9, d=5, chains=4, it=60, niter=5, seed=17, df0=orc.default_df0(9))
        m = td._master(method, sc, df_select='snr')
        info, (ms, Ss) = m.run(sc['niter'], verbose=False, seed=sc['seed'])
        S_mix, m_mix = m.mix_phi()
        ret[rank] = dict(info=info, m=ms, S=Ss, df=list(m.history['df']), n_local=m._shard.n_local,
                         S_mix=S_mix.copy(), m_mix=m_mix.copy(),
                         draws=m._shard.ctx.draws.copy(), k=(m._shard.k_begin, m._shard.k_end))
    finally:
        dist.destroy_process_group()


def test_two_and_three_rank_damping_selection():
    """df_select='snr' over 2 and 3 ranks (uneven shards): the all-reduced statistics give every rank the
    same damping as the single-process oracle replay."""
    from conftest import relerr
    sys.path.insert(0, TESTS)
    import test_damping as td
    from oracle import ep_linalg as orc
    sc = td._scenario(K=9, d=5, chains=4, it=60, niter=5, seed=17, df0=orc.default_df0(9))
    oinfo, oms, oSs, _, _ = td._oracle_run(sc, df_select='snr')
    for world in (2, 3):
        port = 29700 + (os.getpid() % 2000) + world
        ctx = mp.get_context('spawn')
        ret = ctx.Manager().dict()
        procs = [ctx.Process(target=_worker_snr, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert sorted(ret.keys()) == list(range(world))
        assert sum(ret[r]['n_local'] for r in range(world)) == 9
        for r in range(world):
            assert ret[r]['info'] == oinfo == 0
            assert relerr(ret[r]['m'], oms) < 1e-10 and relerr(ret[r]['S'], oSs) < 1e-10
            assert ret[r]['df'] == ret[0]['df']
        # Master.mix_phi over the ranks == the reference's pooling formula (method.py:1280-1298) on all the draws
        draws = np.concatenate([ret[r]['draws'][:ret[r]['k'][1] - ret[r]['k'][0]] for r in range(world)], axis=0)
        K, d, n = draws.shape
        assert K == 9
        means = draws.mean(axis=2)
        m_ref = means.mean(axis=0)
        S_ref = np.zeros((d, d))
        for k in range(K):
            xc = draws[k] - means[k][:, None]
            S_ref += xc @ xc.T + n * np.outer(means[k] - m_ref, means[k] - m_ref)
        S_ref /= K * n - 1
        for r in range(world):
            assert relerr(ret[r]['m_mix'], m_ref) < 1e-12 and relerr(ret[r]['S_mix'], S_ref) < 1e-10


@pytest.mark.parametrize('tag', ['runA', 'runC', 'runD', 'runF'])
def test_two_rank_master_matches_reference(golden, tag):
    from conftest import relerr
    world = 2
    port = 29500 + (os.getpid() % 2000) + {'runA': 0, 'runC': 1, 'runD': 2, 'runF': 3}[tag]
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, tag, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    g = golden['master']
    assert sorted(ret.keys()) == [0, 1]
    assert ret[0]['k'][1] == ret[1]['k'][0]          # contiguous shards
    for r in range(world):
        o = ret[r]
        assert o['info'] == int(g[tag + '_info'])
        assert relerr(o['m'], g[tag + '_m']) < 1e-10 and relerr(o['S'], g[tag + '_S']) < 1e-10
        if o['info'] == 0:
            # host mirrors are complete on every rank
            assert relerr(o['Q'], g[tag + '_Q']) < 1e-10
            assert relerr(o['Qi'], g[tag + '_Qi']) < 1e-10 and relerr(o['ri'], g[tag + '_ri']) < 1e-10
        if tag == 'runC':
            assert o['info2'] == int(g['runC2_info'])
            assert relerr(o['m2'], g['runC2_m']) < 1e-10 and relerr(o['S2'], g['runC2_S']) < 1e-10
