"""Host-side pieces of the drop-in package that need no GPU: the data
simulators and priors of experiment/models (same seed -> the reference's data),
distribute_groups, fit-object readers, model-name resolution, and the C ABI
surface (the library loads and exports every symbol include/epgpu.h declares)."""
import os
import re
import sys

import numpy as np
import pytest

import golden_inputs as gi
from oracle import fakes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXP = os.path.join(ROOT, 'ep-stan_b200', 'experiment')


def test_model_simulators_match_reference(golden):
    if EXP not in sys.path:
        sys.path.insert(0, EXP)
    import importlib
    g = golden['models']
    for name in ('m1b', 'm2b', 'm3b', 'm4b', 'm5b'):
        mod = importlib.import_module('models.' + name)
        for tag, kw, npg in (('corr', dict(Sigma_x='rand'), 5), ('iid', dict(), [3, 8])):
            dat = mod.model(6, 3, npg).simulate_data(rng=100, **kw)
            key = 'mdl_%s_%s_' % (name, tag)
            assert np.allclose(dat.X, g[key + 'X'], rtol=1e-12, atol=1e-13)
            assert np.array_equal(dat.y, g[key + 'y']) and np.array_equal(dat.Nj, g[key + 'Nj'])
            assert np.allclose(dat.true_values['phi'], g[key + 'phi'])
            unc = dat.calc_uncertainty()
            assert np.allclose(np.append(unc[0], unc[1]), g[key + 'unc'])
        S0, m0, Q0, r0 = mod.model(6, 3, 5).get_prior()
        assert np.allclose(Q0, g['mdl_%s_Q0' % name]) and np.allclose(r0, g['mdl_%s_r0' % name])
        assert np.allclose(S0 @ Q0, np.eye(Q0.shape[0]))


def test_distribute_groups(golden):
    from epstan.util import distribute_groups
    g = golden['misc']
    cases = (('const', 64, 4, 20), ('const32', 64, 32, 20),
             ('ragged', 40, 7, np.random.RandomState(52).randint(5, 40, size=40)),
             ('ragged2', 33, 32, np.random.RandomState(53).randint(1, 9, size=33)))
    for tag, J, K, Nj in cases:
        Nk, Njk, jind = distribute_groups(J, K, Nj)
        assert np.array_equal(Nk, g['dg_%s_Nk' % tag])
        assert np.array_equal(Njk, g['dg_%s_Njk' % tag])
        assert np.array_equal(jind, g['dg_%s_jind' % tag]) and jind.dtype == np.int32
    Nj, a, b = distribute_groups(5, 5, 3)
    assert a is None and b is None and np.array_equal(Nj, [3] * 5)
    Nk, ppg, _ = distribute_groups(3, 5, np.array([10, 4, 7]))
    assert Nk.sum() == 21 and len(Nk) == 5 and ppg.sum() == 5
    with pytest.raises(ValueError):
        distribute_groups(3, 1, 4)
    with pytest.raises(ValueError):
        distribute_groups(2, 100, 3)


def test_fit_readers_and_model_names():
    from epstan import util
    rng = np.random.RandomState(0)
    x = rng.standard_normal((12, 3))
    fitobj = fakes.FakeFit(x, chains=3, niter=8, warmup=4)
    out = util.copy_fit_samples(fitobj, 'phi')
    assert out.flags['F_CONTIGUOUS'] and np.array_equal(out, x)
    last = util.get_last_fit_sample(fitobj)
    assert len(last) == 3 and np.array_equal(last[1]['phi'], x[7])
    m = util.load_stan('/some/path/models/m4b_sg.stan')
    assert m.family == 'm4b' and m.single_group and m.dphi(16) == 34
    assert util.load_stan('m1b.pkl').dphi(19) == 20 and not util.load_stan('m3b').single_group
    with pytest.raises(ValueError):
        util.load_stan('m2a')


def test_c_abi_surface():
    """The library loads without a device and exports every declared symbol; a
    context cannot be created on a machine without a GPU (no CPU fallback)."""
    from epstan import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'epgpu.h')).read()
    declared = set(re.findall(r'\b(epg_[a-z_0-9]+)\s*\(', header))
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert hasattr(lib, name)
    assert lib.epg_version() == 1
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(_lib.EpgError):
            _lib.Context(0)


def test_c_abi_from_plain_c(tmp_path):
    """The boundary is usable from C: the header compiles as C99 (no C++ or torch types) and a C
    program that dlopens the library resolves every entry point and gets a loud failure, not a
    fallback, from epg_create on a machine without a GPU."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    from epstan import _lib
    names = sorted(name for name, _, _ in _lib.SYMBOLS)
    src = tmp_path / 'abi.c'
    src.write_text(
        '#include <dlfcn.h>\n#include <stdio.h>\n#include "epgpu.h"\n'
        'static const char* names[] = {%s};\n'
        'int main(int argc, char** argv) {\n'
        '    void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);\n'
        '    if (!h) { fprintf(stderr, "%%s\\n", dlerror()); return 2; }\n'
        '    for (unsigned i = 0; i < sizeof(names) / sizeof(names[0]); ++i)\n'
        '        if (!dlsym(h, names[i])) { fprintf(stderr, "missing %%s\\n", names[i]); return 3; }\n'
        '    int (*version)(void) = (int (*)(void))dlsym(h, "epg_version");\n'
        '    int (*create)(epg_ctx**, int, void*, int) = (int (*)(epg_ctx**, int, void*, int))dlsym(h, "epg_create");\n'
        '    epg_ctx* ctx = 0;\n'
        '    int rc = create(&ctx, 0, 0, 1);\n'
        '    printf("%%d %%d %%d\\n", version(), rc, ctx != 0);\n'
        '    return 0;\n}\n' % ', '.join('"%s"' % n for n in names))
    exe = tmp_path / 'abi'
    subprocess.run(['gcc', '-std=gnu99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src),
                    '-o', str(exe), '-ldl'], check=True)
    out = subprocess.run([str(exe), _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == '1'
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert int(out[1]) != 0 and out[2] == '0'


def test_site_partition_properties():
    """Size-independent properties of the host-side partitions (hypothesis): `distribute_groups` keeps
    every observation exactly once, in order, with within-site group indices 0..J_k-1; `Comm.shard`
    tiles the sites over the ranks in contiguous, balanced blocks."""
    from hypothesis import given, settings, strategies as st
    from epstan.util import distribute_groups
    from epstan._comm import Comm

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 30), min_size=3, max_size=40), st.data())
    def merge(Nj, data):
        J = len(Nj)
        K = data.draw(st.integers(2, J - 1))
        Nk, Nj_k, j_ind_k = distribute_groups(J, K, np.array(Nj))
        assert len(Nk) == K and Nk.sum() == sum(Nj) and Nj_k.sum() == J and np.all(Nj_k >= 1)
        assert j_ind_k.shape == (sum(Nj),)
        row, grp = 0, 0
        for k in range(K):                              # groups stay whole, adjacent and ordered
            assert Nk[k] == sum(Nj[grp:grp + Nj_k[k]])
            seg = j_ind_k[row:row + Nk[k]]
            assert np.array_equal(seg, np.repeat(np.arange(Nj_k[k]), Nj[grp:grp + Nj_k[k]]))
            row += Nk[k]
            grp += Nj_k[k]

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 30), min_size=2, max_size=12), st.data())
    def split(Nj, data):
        J, N = len(Nj), sum(Nj)
        if N <= J:
            return
        K = data.draw(st.integers(J + 1, N))
        Nk, parts, none = distribute_groups(J, K, np.array(Nj))
        assert none is None and len(Nk) == K and parts.sum() == K and np.all(Nk >= 1)
        assert np.array_equal(np.add.reduceat(Nk, np.concatenate(([0], np.cumsum(parts)[:-1]))), Nj)

    @settings(max_examples=100, deadline=None)
    @given(st.integers(1, 5000), st.integers(1, 64))
    def shards(K, size):
        edges = []
        for rank in range(size):
            c = Comm()
            c.rank, c.size = rank, size
            edges.append(c.shard(K))
        assert edges[0][0] == 0 and edges[-1][1] == K
        assert all(edges[i][1] == edges[i + 1][0] for i in range(size - 1))
        lens = [b - a for a, b in edges]
        assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)

    merge()
    split()
    shards()


def test_bench_problem_is_shard_consistent():
    """bench.py generates every site from its own seed, so a rank that builds only its shard holds the
    same rows as the single-GPU run (the N-GPU bench lines measure the same problem)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench
    K, n_k, D = 6, 40, 5
    Xf, yf, prior = bench.build_problem('m3b', K, n_k, D, 0, K)
    for size in (2, 4):
        for rank in range(size):
            base, rem = divmod(K, size)
            k0 = rank * base + min(rank, rem)
            k1 = k0 + base + (1 if rank < rem else 0)
            Xs, ys, _ = bench.build_problem('m3b', K, n_k, D, k0, k1)
            assert np.array_equal(Xs[k0 * n_k:k1 * n_k], Xf[k0 * n_k:k1 * n_k])
            assert np.array_equal(ys[k0 * n_k:k1 * n_k], yf[k0 * n_k:k1 * n_k])
            assert not Xs[:k0 * n_k].any() and not Xs[k1 * n_k:].any()
    assert prior['Q'].shape == (D + 1, D + 1) and set(np.unique(yf)) <= {0, 1}
    assert abs(bench.default_df0(32)(1) - 0.5) < 1e-15
    # the default bench data are the repository's simulators (== the reference's, golden) with seed 100:
    # every rank simulates the same full data set
    Xa, ya, pa = bench.build_problem('m3b', K, n_k, D, 0, 2, 'sim')
    Xb, yb, pb = bench.build_problem('m3b', K, n_k, D, 4, 6, 'sim')
    assert np.array_equal(Xa, Xb) and np.array_equal(ya, yb) and Xa.shape == (K * n_k, D)
    assert np.allclose(np.diag(pa['Q']), 1 / 1.5 ** 2)
    m0, S0 = np.zeros(3), np.eye(3)
    assert abs(bench.kl_mvn(m0, S0, m0, S0)) < 1e-14 and bench.kl_mvn(m0, S0, m0 + 1, 2 * S0) > 0
