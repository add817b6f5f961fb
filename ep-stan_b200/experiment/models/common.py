"""Shared pieces of the simulated hierarchical models (host side, NumPy).

Written for this repository; behaviour follows the reference's
experiment/models/common.py: `rand_corr_vine` (:33-78), the class-balance rule
of `calc_input_param_classification` (:132-317) and the `data` container
(:320-404).  Given the same seed these produce the same arrays as the reference.
"""

import numpy as np
from scipy.special import erfinv, logit

# classification design targets (reference common.py:20-31)
P_0 = 0.2            # a group's expected class share stays within [P_0, 1-P_0] ...
GAMMA_0 = 0.01       # ... except for a GAMMA_0 tail of the inputs
SIGMA_F0 = 0.25      # smallest allowed sd of the linear predictor
_ERFINV = erfinv(2 * GAMMA_0 - 1)
_LOGIT_P0 = logit(P_0)
DELTA_MAX = np.sqrt(2) * SIGMA_F0 * _ERFINV - _LOGIT_P0


def rand_corr_vine(d, alpha=2, beta=2, pmin=-0.8, pmax=0.8, seed=None):
    """Random correlation matrix by the C-vine construction of Lewandowski,
    Kurowicka and Joe (2009): partial correlations ~ Beta(alpha, beta) scaled to
    [pmin, pmax], converted to raw correlations, variables permuted."""
    rs = seed if isinstance(seed, np.random.RandomState) else np.random.RandomState(seed)
    iu = np.triu_indices(d, 1)
    pc = rs.beta(alpha, beta, size=len(iu[0])) * (pmax - pmin) + pmin
    P = np.zeros((d, d))
    P[iu] = pc                    # partial correlation rho_{ij ; 0..i-1}
    P2 = np.zeros((d, d))
    P2[iu] = pc ** 2
    C = np.eye(d)
    for i in range(d - 1):
        for j in range(i + 1, d):
            r = P[i, j]
            for k in range(i - 1, -1, -1):      # peel the conditioning variables off
                r = r * np.sqrt((1 - P2[k, i]) * (1 - P2[k, j])) + P[k, i] * P[k, j]
            C[i, j] = C[j, i] = r
    perm = rs.permutation(d)
    return C[np.ix_(perm, perm)]


def calc_input_param_classification(alpha, beta, Sigma_x=None):
    """Per-group input mean `mu_x` and scale `sigma_x` that keep the classes of a
    logistic model balanced: with x ~ N(mu_x, sigma_x^2 Sigma_x) the linear
    predictor alpha + beta'x keeps P(y=1) within [P_0, 1-P_0] but for a GAMMA_0
    tail; groups whose intercept alone is too extreme get a mean shift and the
    minimum predictor sd SIGMA_F0.  `alpha` () or (J,), `beta` (D,) or (J, D)."""
    alpha = np.asarray(alpha, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    scalar = alpha.ndim == 0 and beta.ndim < 2
    a = np.atleast_1d(alpha)
    B = np.atleast_2d(beta)
    J = max(a.shape[0], B.shape[0])
    a = np.broadcast_to(a, (J,))
    B = np.broadcast_to(B, (J, B.shape[1]))
    quad = np.sum(B * B, axis=1) if Sigma_x is None else np.sum((B @ Sigma_x) * B, axis=1)
    sd_unit = np.sqrt(quad)                       # sd of beta'x at sigma_x = 1
    mild = np.abs(a) < DELTA_MAX
    mu_x = np.zeros(J)
    sigma_x = np.empty(J)
    sigma_x[mild] = (_LOGIT_P0 + np.abs(a[mild])) / (np.sqrt(2) * _ERFINV * sd_unit[mild])
    hard = ~mild
    mu_x[hard] = (np.sign(a[hard]) * DELTA_MAX - a[hard]) / np.sum(B[hard], axis=1)
    sigma_x[hard] = SIGMA_F0 / sd_unit[hard]
    if scalar:
        return mu_x[0], sigma_x[0]
    return mu_x, sigma_x


class data(object):
    """Simulated data set: X, y, X_param, y_true, Nj, N, J, j_lim, j_ind, true_values."""

    def __init__(self, X, y, X_param, y_true, Nj, j_lim, j_ind, true_values):
        self.X, self.y, self.X_param, self.y_true = X, y, X_param, y_true
        self.Nj, self.j_lim, self.j_ind, self.true_values = Nj, j_lim, j_ind, true_values
        self.N = int(np.sum(Nj))
        self.J = Nj.shape[0]

    def calc_uncertainty(self):
        """(global, per group) share of observations whose class differs from the
        noise-free class (classification) or R^2 (regression)."""
        grp = np.repeat(np.arange(self.J), self.Nj)
        if issubclass(self.y.dtype.type, np.integer):
            wrong = (self.y_true != self.y).astype(np.float64)
            return wrong.sum() / self.N, np.bincount(grp, wrong, self.J) / self.Nj
        sse = np.square(self.y - self.y_true)
        ymean_g = np.bincount(grp, self.y, self.J) / self.Nj
        sst_g = np.bincount(grp, np.square(self.y - ymean_g[grp]), self.J)
        return (1 - sse.sum() / np.sum(np.square(self.y - self.y.mean())),
                1 - np.bincount(grp, sse, self.J) / sst_g)
