"""fp64 NumPy restatement of Stan 2.17's adaptive diag_e NUTS.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference samples every tilted distribution with PyStan 2.17.0.0
(epstan/util.py:716 ``model.sampling``; call parameters from
Worker.DEFAULT_STAN_PARAMS, method.py:154-160 and fit.py:301-304).  PyStan /
Stan are un-vendored dependencies and are not installed here, so this follows
the published algorithm of that release:

  * NUTS with multinomial sampling over the trajectory and the generalised
    U-turn criterion (Betancourt 2017, "A conceptual introduction to HMC",
    appendix A; stan/mcmc/hmc/nuts/base_nuts.hpp), max_treedepth 10,
    divergence threshold 1000;
  * diagonal Euclidean metric, leapfrog integrator;
  * step-size dual averaging (Hoffman & Gelman 2014): delta 0.8, gamma 0.05,
    t0 10, kappa 0.75, mu = log(10 eps); step-size initialisation heuristic;
  * windowed variance adaptation: init buffer 75, term buffer 50, base window
    25, rescaled to 15 % / 10 % / rest when warm-up is shorter than 150, none
    below 20; regularised variance  n/(n+5) var + 1e-3 * 5/(n+5);
  * random inits U(-2, 2) on the unconstrained scale.

It is written recursively (like Stan), whereas the CUDA kernel builds its trees
iteratively with an explicit stack -- the two are independent formulations.

Parity status: UNPINNED (no Stan here).  Statistical checks only
(tests/test_oracle_nuts.py: exact Gaussian targets; tests/test_gpu_sampler.py:
GPU vs this sampler within 4x MCSE).
"""

import numpy as np

MAX_DELTA_H = 1000.0


class DualAveraging(object):
    def __init__(self, delta=0.8, gamma=0.05, t0=10.0, kappa=0.75):
        self.delta, self.gamma, self.t0, self.kappa = delta, gamma, t0, kappa
        self.mu = np.log(10.0)
        self.restart()

    def restart(self):
        self.counter = 0
        self.s_bar = 0.0
        self.x_bar = 0.0

    def learn(self, accept_stat):
        self.counter += 1
        a = min(1.0, accept_stat)
        eta = 1.0 / (self.counter + self.t0)
        self.s_bar = (1 - eta) * self.s_bar + eta * (self.delta - a)
        x = self.mu - self.s_bar * np.sqrt(self.counter) / self.gamma
        x_eta = self.counter ** (-self.kappa)
        self.x_bar = (1 - x_eta) * self.x_bar + x_eta * x
        return np.exp(x)

    def final(self):
        return np.exp(self.x_bar)


class VarWindows(object):
    def __init__(self, num_warmup, dim):
        self.num_warmup = num_warmup
        self.init_buffer, self.term_buffer, self.base_window = 75, 50, 25
        self.enabled = num_warmup >= 20
        if self.enabled and self.init_buffer + self.base_window + self.term_buffer > num_warmup:
            self.init_buffer = int(0.15 * num_warmup)
            self.term_buffer = int(0.1 * num_warmup)
            self.base_window = num_warmup - (self.init_buffer + self.term_buffer)
        self.counter = 0
        self.window_size = self.base_window
        self.next_window = self.init_buffer + self.base_window - 1
        self.n = 0
        self.mean = np.zeros(dim)
        self.m2 = np.zeros(dim)

    def _compute_next_window(self):
        last = self.num_warmup - self.term_buffer - 1
        if self.next_window == last:
            return
        self.window_size *= 2
        self.next_window = self.counter + self.window_size
        if self.next_window == last:
            return
        boundary = self.next_window + 2 * self.window_size
        if boundary >= self.num_warmup - self.term_buffer:
            self.next_window = last

    def learn(self, q):
        """returns the new inverse metric at the end of a window, else None"""
        out = None
        if self.enabled:
            in_window = (self.init_buffer <= self.counter < self.num_warmup - self.term_buffer
                         and self.counter != self.num_warmup)
            if in_window:
                self.n += 1
                delta = q - self.mean
                self.mean += delta / self.n
                self.m2 += (q - self.mean) * delta
            if self.counter == self.next_window and self.counter != self.num_warmup:
                self._compute_next_window()
                n = float(self.n)
                var = self.m2 / (n - 1.0)
                out = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0))
                self.n = 0
                self.mean[:] = 0.0
                self.m2[:] = 0.0
        self.counter += 1
        return out


class NUTS(object):
    """One chain.  ``lp_grad(q) -> (lp, grad)`` for a single point q (1-D)."""

    def __init__(self, lp_grad, dim, rng, max_depth=10):
        self.lp_grad, self.dim, self.rng, self.max_depth = lp_grad, dim, rng, max_depth
        self.eps = 1.0
        self.minv = np.ones(dim)
        self.n_grad = 0

    # -- Hamiltonian pieces --
    def _vg(self, q):
        lp, g = self.lp_grad(q)
        self.n_grad += 1
        return -lp, -g

    def _H(self, V, p):
        h = V + 0.5 * np.dot(p, self.minv * p)
        return np.inf if np.isnan(h) else h

    def _leapfrog(self, z, eps):
        q, p, V, g = z
        p = p - 0.5 * eps * g
        q = q + eps * self.minv * p
        V, g = self._vg(q)
        p = p - 0.5 * eps * g
        return (q, p, V, g)

    def _sample_p(self):
        return self.rng.standard_normal(self.dim) / np.sqrt(self.minv)

    def init_stepsize(self, q, V, g):
        def trial():
            p = self._sample_p()
            H0 = self._H(V, p)
            z = self._leapfrog((q, p, V, g), self.eps)
            return H0 - self._H(z[2], z[1])
        direction = 1 if trial() > np.log(0.8) else -1
        while True:
            dH = trial()
            if direction == 1 and not (dH > np.log(0.8)):
                break
            if direction == -1 and not (dH < np.log(0.8)):
                break
            self.eps = 2 * self.eps if direction == 1 else 0.5 * self.eps
            if self.eps > 1e7 or self.eps == 0:
                raise RuntimeError("step size heuristic failed")

    # -- tree building --
    def _criterion(self, psharp_l, psharp_r, rho):
        return np.dot(psharp_r, rho) > 0 and np.dot(psharp_l, rho) > 0

    def _build(self, depth, z, sign, H0, st):
        """Returns (valid, z_end, z_propose, rho, log_sum_weight, psharp_left, psharp_right)."""
        if depth == 0:
            z = self._leapfrog(z, sign * self.eps)
            st['n_leapfrog'] += 1
            h = self._H(z[2], z[1])
            dH = H0 - h
            st['sum_metro'] += 1.0 if dH > 0 else np.exp(dH)
            if -dH > MAX_DELTA_H:
                st['divergent'] = True
                return False, z, z, z[1], dH, None, None
            ps = self.minv * z[1]
            return True, z, z, z[1].copy(), dH, ps, ps
        ok, z, prop_l, rho_l, lsw_l, psl, _ = self._build(depth - 1, z, sign, H0, st)
        if not ok:
            return False, z, prop_l, rho_l, lsw_l, None, None
        ok, z, prop_r, rho_r, lsw_r, _, psr = self._build(depth - 1, z, sign, H0, st)
        if not ok:
            return False, z, prop_l, rho_l, lsw_l, None, None
        lsw = np.logaddexp(lsw_l, lsw_r)
        if lsw_r > lsw or self.rng.uniform() < np.exp(lsw_r - lsw):
            prop = prop_r
        else:
            prop = prop_l
        rho = rho_l + rho_r
        return self._criterion(psl, psr, rho), z, prop, rho, lsw, psl, psr

    def transition(self, q, V, g):
        p = self._sample_p()
        z0 = (q, p, V, g)
        H0 = self._H(V, p)
        z_minus = z_plus = z_sample = z0
        rho = p.copy()
        lsw = 0.0
        depth = 0
        st = {'n_leapfrog': 0, 'sum_metro': 0.0, 'divergent': False}
        while depth < self.max_depth:
            if self.rng.uniform() > 0.5:
                ok, z_plus, prop, rho_sub, lsw_sub, _, _ = self._build(depth, z_plus, 1, H0, st)
            else:
                ok, z_minus, prop, rho_sub, lsw_sub, _, _ = self._build(depth, z_minus, -1, H0, st)
            if not ok:
                break
            depth += 1
            if lsw_sub > lsw or self.rng.uniform() < np.exp(lsw_sub - lsw):
                z_sample = prop
            lsw = np.logaddexp(lsw, lsw_sub)
            rho = rho + rho_sub
            if not self._criterion(self.minv * z_minus[1], self.minv * z_plus[1], rho):
                break
        accept = st['sum_metro'] / max(st['n_leapfrog'], 1)
        return z_sample[0], z_sample[2], z_sample[3], accept, st


def sample_chain(lp_grad, dim, n_iter, n_warmup, rng, q0=None, delta=0.8, max_depth=10):
    """Adaptive NUTS run.  Returns dict(draws (n_iter-n_warmup, dim), stepsize, n_grad, last)."""
    s = NUTS(lp_grad, dim, rng, max_depth)
    tries = 0
    while True:
        q = rng.uniform(-2, 2, size=dim) if q0 is None else np.array(q0, dtype=np.float64)
        V, g = s._vg(q)
        if np.isfinite(V) and np.all(np.isfinite(g)):
            break
        tries += 1
        if q0 is not None or tries > 100:
            raise RuntimeError("initialisation failed")
    da = DualAveraging(delta=delta)
    win = VarWindows(n_warmup, dim)
    s.init_stepsize(q, V, g)
    draws = np.empty((n_iter - n_warmup, dim))
    eps_used = []
    n_div = 0
    for it in range(n_iter):
        q, V, g, accept, st = s.transition(q, V, g)
        n_div += st['divergent'] and it >= n_warmup     # post-warm-up divergences
        if it < n_warmup:
            s.eps = da.learn(accept)
            new_minv = win.learn(q)
            if new_minv is not None:
                s.minv = new_minv
                s.init_stepsize(q, V, g)
                da.mu = np.log(10 * s.eps)
                da.restart()
            if it == n_warmup - 1:
                s.eps = da.final()
        else:
            draws[it - n_warmup] = q
            eps_used.append(s.eps)
    return dict(draws=draws, stepsize=float(np.mean(eps_used)) if eps_used else s.eps,
                n_grad=s.n_grad, last=q, minv=s.minv, n_divergent=n_div)


def sample(lp_grad, dim, chains, n_iter, n_warmup=None, seed=0, inits=None, **kw):
    """`chains` independent chains; draws stacked chain-major like util.copy_fit_samples."""
    if n_warmup is None:
        n_warmup = n_iter // 2
    res = [sample_chain(lp_grad, dim, n_iter, n_warmup, np.random.RandomState([seed, c]),
                        q0=None if inits is None else inits[c], **kw) for c in range(chains)]
    return dict(draws=np.concatenate([r['draws'] for r in res], axis=0),
                per_chain=[r['draws'] for r in res],
                stepsize=float(np.mean([r['stepsize'] for r in res])),
                n_grad=sum(r['n_grad'] for r in res), last=[r['last'] for r in res],
                n_divergent=sum(r['n_divergent'] for r in res))


def split_rhat(per_chain):
    """Split R-hat (BDA3) of a list of (n, dim) chains."""
    halves = []
    for c in per_chain:
        h = c.shape[0] // 2
        halves += [c[:h], c[h:2 * h]]
    n = halves[0].shape[0]
    means = np.array([h.mean(axis=0) for h in halves])
    W = np.mean([h.var(axis=0, ddof=1) for h in halves], axis=0)
    B = n * means.var(axis=0, ddof=1)
    return np.sqrt(((n - 1) / n * W + B / n) / W)


def ess_mcse(per_chain):
    """Crude effective sample size / MCSE of the mean via batch means."""
    out = []
    x = np.concatenate(per_chain, axis=0)
    n = x.shape[0]
    nb = max(int(np.sqrt(n)), 2)
    bs = n // nb
    bm = x[:nb * bs].reshape(nb, bs, -1).mean(axis=1)
    mcse = np.sqrt(bm.var(axis=0, ddof=1) / nb)
    return x.var(axis=0, ddof=1) / np.maximum(mcse ** 2, 1e-300), mcse
