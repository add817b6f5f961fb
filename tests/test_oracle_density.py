"""Pins oracle/density.py: gradient vs central finite differences, log-density
vs an independent scalar evaluation that reads like the Stan model blocks."""
import numpy as np
import pytest

from oracle import density as dens
import synth


@pytest.mark.parametrize('model', dens.MODELS)
@pytest.mark.parametrize('J', [1, 3])
def test_density_gradient(model, J):
    site = synth.make_site(model, n=40, D=4, J=J, seed=3)
    td = synth.oracle_density(model, site)
    rng = np.random.RandomState(0)
    q = 0.5 * rng.standard_normal((3, td.p))
    lp, grad = td.lp_grad(q)
    for i in range(3):
        assert abs(lp[i] - td.lp_scalar(q[i])) < 1e-9 * max(1.0, abs(lp[i]))
        h = 1e-6
        for j in range(td.p):
            e = np.zeros(td.p)
            e[j] = h
            fd = (td.lp_grad(q[i] + e)[0][0] - td.lp_grad(q[i] - e)[0][0]) / (2 * h)
            assert abs(fd - grad[i, j]) < 1e-5 * max(1.0, abs(grad[i, j]))
    assert td.p == dens.num_params(model, 4, J) and td.d == dens.dphi(model, 4)
