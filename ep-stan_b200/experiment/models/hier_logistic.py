"""Hierarchical logistic-regression simulators behind m1b / m2b / m3b / m4b / m5b.

One table-driven implementation; the RNG call order follows the reference's
simulators (experiment/models/m1b.py:83-176, m2b.py:95-176, m3b.py, m4b.py:95-195,
m5b.py:98-200) so that the same `seed_data` gives the same data set.
"""

import numpy as np
from scipy.linalg import cholesky

from .common import data, calc_input_param_classification, rand_corr_vine

B_ABS_MIN_SUM = 1e-4     # keep |sum(beta)| away from zero (mean shifts divide by it)

SPEC = {
    # prior variances are those of the reference modules (m1b.py:46-51, m3b.py:45-50, m4b.py:45-59)
    'm1b': dict(dphi=lambda D: D + 1, group_slopes=False),
    'm2b': dict(dphi=lambda D: 2, group_slopes=False),
    'm3b': dict(dphi=lambda D: D + 1, group_slopes=True),
    'm4b': dict(dphi=lambda D: 2 * D + 2, group_slopes=True),
    'm5b': dict(dphi=lambda D: 2 * D + 2, group_slopes=True),
}


class HierLogistic(object):
    family = None

    def __init__(self, J, D, npg):
        self.J, self.D, self.npg = J, D, npg
        self.dphi = SPEC[self.family]['dphi'](D)

    # -- truth ---------------------------------------------------------------
    def _draw_truth(self, rng):
        J, D, fam = self.J, self.D, self.family
        if fam == 'm1b':
            sigma_a = 1.0
            beta = rng.randn(D) * 1.0
            s = beta.sum()
            while abs(s) < B_ABS_MIN_SUM:
                i = rng.randint(D)
                s -= beta[i]
                beta[i] = rng.randn() * 1.0
                s += beta[i]
            alpha_j = rng.randn(J) * sigma_a
            return alpha_j, beta, np.append(np.log(sigma_a), beta), beta
        if fam == 'm2b':
            # one slope vector for all groups, scaled by a single sigma_b (intercepts first, then slopes)
            sigma_a, sigma_b = 1.0, 1.0
            alpha_j = rng.randn(J) * sigma_a
            beta = rng.randn(D) * sigma_b
            s = beta.sum()
            while abs(s) < B_ABS_MIN_SUM:
                i = rng.randint(D)
                s -= beta[i]
                beta[i] = rng.randn() * sigma_b
                s += beta[i]
            return alpha_j, beta, np.array([np.log(sigma_a), np.log(sigma_b)]), beta
        if fam == 'm3b':
            sigma_a = 1.0
            sigma_b = np.exp(rng.randn(D) * 1.0)
            alpha_j = rng.randn(J) * sigma_a
            beta_j = rng.randn(J, D) * sigma_b
            scale = sigma_b
            phi = np.append(np.log(sigma_a), np.log(sigma_b))
        elif fam == 'm5b':
            # heavy-tailed truth: half-Cauchy scales, Laplace locations and latents
            mu_a, sigma_a = 0.1, 1.0
            sigma_b = np.abs(rng.standard_cauchy(D) * 1.0)
            mu_b = rng.laplace(size=D) * 0.0
            alpha_j = mu_a + rng.laplace(size=J) * sigma_a
            beta_j = mu_b + rng.laplace(size=(J, D)) * sigma_b
            scale = sigma_b
            phi = np.concatenate(([mu_a, np.log(sigma_a)], mu_b, np.log(sigma_b)))
        else:
            mu_a, sigma_a = 1.5, np.exp(0.4)
            mu_b = rng.rand(D) * 4.0 - 2.0
            sigma_b = np.exp(rng.rand(D) * 1.0 - 0.5)
            alpha_j = mu_a + rng.randn(J) * sigma_a
            beta_j = mu_b + rng.randn(J, D) * sigma_b
            scale = sigma_b
            phi = np.concatenate(([mu_a, np.log(sigma_a)], mu_b, np.log(sigma_b)))
        for j in range(J):
            s = beta_j[j].sum()
            while abs(s) < B_ABS_MIN_SUM:
                i = rng.randint(D)
                s -= beta_j[j, i]
                beta_j[j, i] = (0.0 if fam == 'm3b' else mu_b[i]) + rng.randn() * scale[i]
                s += beta_j[j, i]
        return alpha_j, beta_j, phi, beta_j

    def simulate_data(self, Sigma_x=None, rng=None):
        """Returns a `common.data` instance.  `Sigma_x`: None (identity), 'rand'
        (random vine correlation matrix) or an ndarray."""
        J, D, npg = self.J, self.D, self.npg
        if not isinstance(rng, np.random.RandomState):
            rng = np.random.RandomState(rng)
        seed_input_cov = rng.randint(2 ** 31 - 1)
        if isinstance(Sigma_x, str) and Sigma_x == 'rand':
            Sigma_x = rand_corr_vine(D, seed=seed_input_cov)
        if hasattr(npg, '__getitem__') and len(npg) == 2:
            Nj = rng.randint(npg[0], npg[1] + 1, size=J)
        else:
            Nj = npg * np.ones(J, dtype=np.int64)
        N = int(np.sum(Nj))
        j_lim = np.concatenate(([0], np.cumsum(Nj)))
        j_ind = np.repeat(np.arange(J), Nj).astype(np.int64)
        alpha_j, beta_any, phi_true, beta_out = self._draw_truth(rng)
        mu_x_j, sigma_x_j = calc_input_param_classification(alpha_j, beta_any, Sigma_x)
        Z = rng.randn(N, D)
        if Sigma_x is None:
            X = mu_x_j[j_ind, None] + Z * sigma_x_j[j_ind, None]
        else:
            X = mu_x_j[j_ind, None] + (Z @ cholesky(Sigma_x)) * sigma_x_j[j_ind, None]
        if beta_any.ndim == 1:
            f = alpha_j[j_ind] + X.dot(beta_any)
        else:
            f = alpha_j[j_ind] + np.einsum('nd,nd->n', X, beta_any[j_ind])
        p = 1 / (1 + np.exp(-f))
        y_true = (0.5 < p).astype(int)
        y = (rng.rand(N) < p).astype(int)
        return data(X, y, {'mu_x': mu_x_j, 'sigma_x': sigma_x_j, 'Sigma_x': Sigma_x},
                    y_true, Nj, j_lim, j_ind, {'phi': phi_true, 'alpha': alpha_j, 'beta': beta_out})

    # -- prior ---------------------------------------------------------------
    def get_prior(self):
        """Returns S, m, Q, r of the Gaussian prior on phi."""
        D = self.D
        if self.family == 'm4b':
            var = np.concatenate(([4.0 ** 2, 2.0 ** 2], np.full(D, 4.0 ** 2), np.full(D, 2.0 ** 2)))
        else:
            var = np.full(self.dphi, 1.5 ** 2)
        m0 = np.zeros(self.dphi)
        return np.diag(var).T, m0, np.diag(1.0 / var).T, m0 / var

    def get_param_definitions(self):
        """names, shapes, hierarchical-dimension index of the inferred parameters."""
        if self.family in ('m1b', 'm2b'):
            return ('alpha', 'beta'), ((self.J,), (self.D,)), (0, None)
        return ('alpha', 'beta'), ((self.J,), (self.J, self.D)), (0, 0)
