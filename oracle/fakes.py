"""Deterministic stand-ins for the sampler, shared by oracle/make_golden.py and
the tests so both sides see identical draws.  TEST INFRASTRUCTURE ONLY."""

import os

import numpy as np
from scipy import linalg


def random_spd(rng, d, scale=1.0, ridge=0.5):
    """Random well-conditioned SPD matrix from a RandomState."""
    A = rng.standard_normal((d, d))
    return scale * (A @ A.T / d + ridge * np.eye(d))


def gaussian_site_factors(seed, K, d):
    """Per-site Gaussian 'likelihood' factors (Q_k, r_k), k=0..K-1."""
    rng = np.random.RandomState(seed)
    Qs = np.stack([random_spd(rng, d, scale=2.0) for _ in range(K)], axis=2)
    rs = rng.standard_normal((d, K))
    return Qs, rs


def gaussian_tilted_draws(seed, cav_mean, cav_prec, Qk, rk, n, inflate=1.0):
    """n exact draws from N(cavity) x N(site factor): precision cav_prec + Qk.

    Depends only on (seed, inputs) via RandomState(seed).standard_normal, which
    NumPy keeps bit-stable across versions.  ``inflate`` scales the draws'
    spread (used to provoke update failures).
    """
    d = cav_mean.shape[0]
    P = cav_prec + Qk
    h = cav_prec @ cav_mean + rk
    U = linalg.cholesky(P, lower=False)
    mean = linalg.cho_solve((U, False), h)
    z = np.random.RandomState(seed).standard_normal((n, d))
    # x = mean + U^-1 z  has covariance P^-1
    x = linalg.solve_triangular(U, z.T, lower=False).T
    return mean + inflate * x


class FakeFit:
    """Duck-typed PyStan fit (interface listed in SURVEY 8b)."""

    def __init__(self, draws, chains, niter, warmup):
        n, d = draws.shape
        per = niter - warmup
        self.model_pars = ['phi']
        self.par_dims = [[d]]
        self.sim = {'chains': chains, 'warmup2': [warmup] * chains, 'samples': []}
        for c in range(chains):
            ch = {'lp__': np.zeros(niter)}
            for i in range(d):
                col = np.zeros(niter)
                col[warmup:] = draws[c * per:(c + 1) * per, i]
                ch['phi[{}]'.format(i)] = col
            self.sim['samples'].append({'chains': ch})

    def get_sampler_params(self):
        return [{'stepsize__': np.full(4, 0.25)} for _ in range(self.sim['chains'])]

    def summary(self):
        return {'summary': np.ones((3, 10))}


class FakeModel:
    """Exact Gaussian 'tilted' sampler keyed on the Stan seed.

    ``constant``: True or a collection of site ids whose draws are degenerate
    (all equal) so that their moment estimate fails."""

    def __init__(self, Qs, rs, inflate=None, constant=False):
        self.Qs, self.rs = Qs, rs
        self.inflate = inflate
        self.constant = constant
        self.seeds = []

    def sampling(self, data, chains, iter, warmup, thin, init, seed, refresh):
        k = int(data['site_id'])
        if warmup is None:
            warmup = iter // 2
        n = chains * (iter - warmup)
        self.seeds.append((k, int(seed)))
        infl = 1.0 if self.inflate is None else self.inflate
        if self.constant is True or (self.constant and k in self.constant):
            # all-zero draws: mean and residuals are exactly 0, so the reference's
            # QR factor is exactly singular and dpotri reports it (no rounding luck)
            x = np.zeros((n, len(data['mu_phi'])))
        else:
            x = gaussian_tilted_draws(
                int(seed), np.array(data['mu_phi']), np.array(data['Omega_phi']),
                self.Qs[:, :, k], self.rs[:, k], n, inflate=infl)
        if not getattr(self, 'quiet', False):
            # the reference scrapes this line from fd 1 (util.py:722-723)
            os.write(1, b'Elapsed Time: 0.01 seconds (Total)\n')
        return FakeFit(x, chains, iter, warmup)
