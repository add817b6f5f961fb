"""CPU oracle for the ep-stan EP inner loop.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``ep-stan_b200/``)
imports this directory.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and
there only as the checker or the timed CPU baseline.

Contents
--------
ep_linalg.py   fp64 NumPy/SciPy restatement of the moment-matching / cavity /
               damped-update path (reference epstan/method.py, epstan/util.py,
               epstan/cython_util.pyx).  PINNED against the unmodified reference
               (imported in the build container) by tests/golden/*.npz, see
               make_golden.py.
density.py     fp64 NumPy restatement of the tilted log-densities and gradients
               of experiment/models/m{1,2,3,4,5}b[_sg].stan.  Pinned by finite
               differences only (Stan itself is not available) -> the sampling
               half is "parity unpinned".
nuts.py        fp64 NumPy restatement of Stan 2.17's adaptive diag_e NUTS
               (published algorithm; PyStan 2.17.0.0 is an un-vendored
               dependency of the reference and is not installed here).
               "parity unpinned": statistical checks only.
fakes.py       deterministic Gaussian sampler double shared by make_golden.py and
               the tests.
make_golden.py generates tests/golden/*.npz from the unmodified reference.
"""
