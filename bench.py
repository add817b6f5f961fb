#!/usr/bin/env python
"""Benchmark of the EP inner loop (BASELINE.json metric: EP iterations/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg3|cfg4|cfg1like] [--sites K] [--chains C] [--siter I]

A "step" is one full EP iteration over all K sites: tilted NUTS sampling of
every site x chain (all draws), moment matching, damped update with the
natural-parameter all-reduce, cavities and global moments.  Default workload is
BASELINE.json configs[3] (SURVEY 8d "config 4"): varying-slope logistic model
m3b_sg, K=1024 sites, n_k=5000, D=49 (d=50), sharded by site over the ranks
(strong scaling: K is fixed, each of N GPUs owns K/N sites).

Prints ONE JSON line (rank 0).  `value` = EP iterations/s with the state
resident in HBM; `e2e` = the same through Master.run() with host state buffers
(upload of Q,r,Qi,ri,dQi,dri,cavities before, download of all mirrors after);
`roofline` = the sampler kernel (dominant) against the measured bf16 tensor
peak, algorithmic flops = 4*n_k*D per gradient evaluation times the
evaluations counted on the device; `cpu_baseline` = the fp64 NumPy oracle
(oracle/nuts.py + oracle/ep_linalg.py) timed on this box's host cores on a
bounded sample of the same sites.
"""

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'ep-stan_b200'))

WORKLOADS = {
    # name: (model, K, n_k, D, chains, siter)
    'cfg4': ('m3b', 1024, 5000, 49, 4, 200),      # BASELINE configs[3]
    'cfg3': ('m1b', 64, 2000, 19, 8, 200),        # BASELINE configs[2]
    'cfg1like': ('m1b', 4, 320, 16, 4, 200),      # shape of configs[0] (single group per site)
}


def dphi(model, D):
    return 2 * D + 2 if model == 'm4b' else D + 1


def site_data(model, k, n_k, D, seed=100):
    """Synthetic data of one site (vectorised; same model family as
    experiment/models/m{1,3,4}b.py simulate_data): standardised correlated
    inputs, group intercept and (m3b/m4b) group slopes."""
    rng = np.random.RandomState([seed, k])
    g = np.random.RandomState(seed)                       # shared truth
    beta = g.standard_normal(D) * (0.5 if model != 'm1b' else 1.0) / np.sqrt(D / 4.0)
    sigma_b = np.exp(0.3 * g.standard_normal(D) - 1.0)
    alpha = rng.standard_normal() * 1.0
    bk = beta + (rng.standard_normal(D) * sigma_b if model != 'm1b' else 0.0)
    X = rng.standard_normal((n_k, D))
    X += 0.3 * rng.standard_normal((n_k, 1))              # common factor -> correlated inputs
    f = alpha + X @ bk
    y = (rng.uniform(size=n_k) < 1.0 / (1.0 + np.exp(-f))).astype(np.int64)
    return X, y


def build_problem(model, K, n_k, D, k_begin, k_end):
    """Full-shape X, y with only the local sites' rows filled in."""
    X = np.zeros((K * n_k, D))
    y = np.zeros(K * n_k, dtype=np.int64)
    for k in range(k_begin, k_end):
        X[k * n_k:(k + 1) * n_k], y[k * n_k:(k + 1) * n_k] = site_data(model, k, n_k, D)
    d = dphi(model, D)
    prior = {'Q': np.eye(d) / 1.5 ** 2, 'r': np.zeros(d)}
    return X, y, prior


def default_df0(K):
    """experiment/fit.py:176-186"""
    b = min(1.0 / K, 0.2)
    a = 0.5 - b
    t = -np.log(0.1) / (K - 1)
    return lambda it: a * np.exp(-t * (it - 1)) + b


class ClockSampler(object):
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc = None
        self.rows = []
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, smax, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(max(smax)) if smax else None, 'reasons': sorted(reasons)}


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p.get('hbm_gbs', 6650.0), p.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


# ---------------------------------------------------------------------------
# CPU baseline: the fp64 oracle on a bounded sample of sites
# ---------------------------------------------------------------------------
def _cpu_site(args):
    """One site's tilted step with the oracle: NUTS draws + moment matching."""
    model, k, n_k, D, chains, siter, Q, r = args
    from oracle import density as dens, nuts, ep_linalg as orc
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)                  # one BLAS thread per worker process
    except Exception:
        pass
    X, y = site_data(model, k, n_k, D)
    d = Q.shape[0]
    td = dens.TiltedDensity(model, X, y, np.linalg.solve(Q, r), Q)      # cavity ~ global (Qi=0)
    t0 = time.perf_counter()
    res = nuts.sample(lambda q: tuple(v[0] for v in td.lp_grad(q[None])), td.p, chains=chains,
                      n_iter=siter, seed=k)
    if res['draws'].shape[0] > d + 2:
        ok, dQ, dr_ = orc.tilted_moments(res['draws'][:, :d], Q, r, 'sample')
        orc.cavity(Q + 0.1 * dQ, r + 0.1 * dr_, 0.1 * dQ, 0.1 * dr_)
    return time.perf_counter() - t0, res['n_grad']


def cpu_baseline(model, K, n_k, D, chains, siter, n_sample, procs):
    """Bounded sample: ONE chain of `siter` iterations on each of `n_sample`
    sites (one process per core), scaled to K sites x `chains` chains."""
    import multiprocessing as mp
    d = dphi(model, D)
    Q = np.eye(d) / 1.5 ** 2
    r = np.zeros(d)
    jobs = [(model, k, n_k, D, 1, siter, Q, r) for k in range(n_sample)]
    t0 = time.perf_counter()
    if procs > 1:
        with mp.get_context('fork').Pool(procs) as pool:
            out = pool.map(_cpu_site, jobs)
    else:
        out = [_cpu_site(j) for j in jobs]
    wall = time.perf_counter() - t0
    n_grad = sum(o[1] for o in out)
    its = 1.0 / (wall * (K * chains) / float(n_sample))   # EP iterations/s for K sites x chains at this rate
    return its, n_grad / wall, wall


def cpu_linalg_baseline(d, n, n_sites=32):
    """Moment matching + cavity of the oracle (fp64 NumPy/LAPACK restatement of method.py:267-302,413-468)
    on `n_sites` sites of n draws, one core: microseconds per site."""
    from oracle import ep_linalg as orc
    rng = np.random.RandomState(0)
    Q = np.eye(d) * 3.0
    r = rng.standard_normal(d)
    Qi = np.eye(d) * 0.5
    ri = np.zeros(d)
    samps = [rng.standard_normal((n, d)) for _ in range(n_sites)]
    from threadpoolctl import threadpool_limits
    with threadpool_limits(1):                          # one BLAS thread: these are 50 x 800 problems
        orc.tilted_moments(samps[0], Q, r, 'sample')   # (library warm-up)
        orc.cavity(Q, r, Qi, ri)
        t0 = time.perf_counter()
        for sm in samps:
            orc.tilted_moments(sm, Q, r, 'sample')
        t1 = time.perf_counter()
        for _ in samps:
            orc.cavity(Q, r, Qi, ri)
        t2 = time.perf_counter()
    return 1e6 * (t1 - t0) / n_sites, 1e6 * (t2 - t1) / n_sites


def run_reference(args, model, K, n_k, D, chains, siter):
    """--impl reference: the reference's CPU path for this workload.  PyStan is
    not installable here, so this is the oracle port (kind "port") with one
    process per host core over sites."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = min(K, max(cores, 2))
    # Bounded sample: a step = ONE chain on each of n_sample sites; the warm-up steps (siter/4 iterations)
    # also calibrate the cost per sampling iteration, and when --steps full-length steps would not fit in
    # REF_BUDGET_S the timed steps run fewer iterations per chain and are scaled to `siter`.
    REF_BUDGET_S = 240.0
    t_begin = time.perf_counter()
    siter_w = max(20, siter // 4)
    wall_w = None
    for i in range(max(args.warmup, 1)):
        _, _, wall_w = cpu_baseline(model, K, n_k, D, chains, siter_w, n_sample, cores)
    per_iter = wall_w / siter_w
    remaining = max(REF_BUDGET_S - (time.perf_counter() - t_begin), 30.0)
    siter_step = siter if per_iter * siter * args.steps <= remaining else \
        max(20, min(siter, int(remaining / (args.steps * per_iter))))
    vals = []
    for i in range(args.steps):
        its, gps, wall = cpu_baseline(model, K, n_k, D, chains, siter_step, n_sample, cores)
        vals.append((its * siter_step / float(siter), gps, wall))
    its = float(np.mean([v[0] for v in vals]))
    line = {
        'impl': 'reference', 'metric': 'EP iterations/sec (K sites, all draws)', 'value': its,
        'unit': 'it/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 / its, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'model': model + '_sg', 'K': K, 'n_k': n_k, 'D': D,
                   'chains': chains, 'siter': siter},
        'grad_evals_per_s': float(np.mean([v[1] for v in vals])),
        'cpu_baseline': {'value': its, 'unit': 'it/s', 'cores': cores, 'kind': 'port',
                         'sample': 'one chain of %d (of %d) iterations on each of %d sites per step (of %d sites x %d '
                                   'chains), fp64 NumPy oracle NUTS, one process per core; scaled to the full workload'
                                   % (siter_step, siter, n_sample, K, chains)},
        'e2e': {'value': its, 'unit': 'it/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg4', choices=sorted(WORKLOADS))
    ap.add_argument('--sites', type=int, default=None)
    ap.add_argument('--chains', type=int, default=None)
    ap.add_argument('--siter', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    model, K, n_k, D, chains, siter = WORKLOADS[args.workload]
    K = args.sites or K
    chains = args.chains or chains
    siter = args.siter or siter
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == 'reference':
        run_reference(args, model, K, n_k, D, chains, siter)
        return

    import torch
    import torch.distributed as dist
    if os.environ.get('BENCH_DEBUG'):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['BENCH_DEBUG']), exit=True)
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                                timeout=datetime.timedelta(seconds=int(os.environ.get('BENCH_NCCL_TIMEOUT', 600))))
    import epstan.method as method
    from epstan import _lib
    method.set_default_stream(torch.cuda.current_stream().cuda_stream)

    d = dphi(model, D)
    base, rem = divmod(K, world)
    k_begin = rank * base + min(rank, rem)
    k_end = k_begin + base + (1 if rank < rem else 0)
    X, y, prior = build_problem(model, K, n_k, D, k_begin, k_end)
    m = method.Master('experiment/models/%s_sg' % model, X, y, site_sizes=np.full(K, n_k), prior=prior,
                      chains=chains, iter=siter, df0=default_df0(K))
    del X, y
    ctx = m._shard.ctx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        nl0 = sum(w.last_n_leapfrog for w in m.workers[k_begin:k_end])
        t0 = time.perf_counter()
        e0.record()
        res = m.run(nsteps, verbose=False, seed=1234, return_analytics=True)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t[0]), float(t[1]), ctx.launch_count() - l0

    # ---- warm-up (untimed) ----
    m.keep_on_device = False
    res, _, _, _ = timed_run(args.warmup)
    if res[0] != 0:
        raise SystemExit("bench.py: EP failed during warm-up with info %d" % res[0])

    # ---- value: state resident in HBM ----
    m.keep_on_device = True
    m.n_leapfrog_total = 0
    clocks = ClockSampler(local_rank) if rank == 0 else None
    res, ms_dev, ms_wall, launches = timed_run(args.steps)
    clk = clocks.stop() if clocks else None
    info, (ms_, Ss_), (stimes, msteps, mrhats, other) = res
    # gradient evaluations of the timed steps (device counters, all ranks)
    n_leap = 0
    samp_s = float(np.sum(stimes))
    evals = torch.tensor([0.0], dtype=torch.float64, device='cuda')
    # last_n_leapfrog holds only the final step; re-derive the total from a counter run below
    # (every step's count is accumulated by the Master in _tilted_all)
    evals[0] = float(getattr(m, 'n_leapfrog_total', 0))
    if world > 1:
        dist.all_reduce(evals, op=dist.ReduceOp.SUM)
    n_leap = float(evals[0])

    # ---- e2e: through Master.run() with host state buffers ----
    m.keep_on_device = False
    m.n_leapfrog_total = 0
    # (long steps: one end-to-end step is enough -- the extra cost is one state upload/download per run() call)
    e2e_steps = args.steps if ms_wall / args.steps < 5e3 else 1
    res2, ms_dev2, ms_wall2, _ = timed_run(e2e_steps)
    n_loc = k_end - k_begin
    h2d = 8 * ((d * d + d) + n_loc * (2 * (d * d + d)) + n_loc * (d * d + d))      # Q,r + Qi,ri,dQi,dri + cavities
    d2h = 8 * ((d * d + d) + n_loc * 4 * (d * d + d)) + 8 * e2e_steps * (d * d + d)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    hbm_peak, tf_peak, peak_src = load_peaks()
    its = args.steps / (ms_wall * 1e-3)
    its_e2e = e2e_steps / (ms_wall2 * 1e-3)
    flops_per_eval = 4.0 * n_k * D
    # sampler kernel time = stimes (device events around the kernel), max over ranks per step
    tf_achieved = (n_leap / max(world, 1)) * flops_per_eval / max(samp_s, 1e-9) / 1e12   # per GPU
    line = {
        'metric': 'EP iterations/sec (K sites, all draws)', 'value': its, 'unit': 'it/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_wall / args.steps,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16',
        'dtype_detail': 'sampler contractions bf16 x bf16 -> f32 on tcgen05, energies f64; moments/updates f64',
        'data': 'synthetic',
        'config': {'workload': args.workload, 'model': model + '_sg', 'K': K, 'n_k': n_k, 'D': D, 'd': d,
                   'chains': chains, 'siter': siter, 'parallelism': 'sites sharded %d-way' % world,
                   'l2': ('inputs larger than L2 (X %.0f MB fp32 + bf16 copy per GPU)' if n_loc * n_k * (D + 3) * 4 > 126e6
                          else 'inputs smaller than L2 (X %.0f MB fp32 per GPU); not flushed: a step re-reads each '
                               'site\'s X thousands of times by design, the first touch is <0.1 %% of a step')
                         % (n_loc * n_k * (D + 3) * 4 / 1e6)},
        'grad_evals_per_s': n_leap / max(samp_s, 1e-9),
        'sampling_share': samp_s / (ms_wall * 1e-3),
        'device_ms_per_step': ms_dev / args.steps,
        'e2e': {'value': its_e2e, 'unit': 'it/s', 'h2d_bytes_per_step': h2d // e2e_steps,
                'd2h_bytes_per_step': d2h // e2e_steps, 'steps': e2e_steps},
        'gpu_launches': launches,
        'clocks': clk,
        'roofline': {'bound': 'tensor', 'achieved': tf_achieved, 'peak': tf_peak, 'unit': 'TFLOP/s',
                     'frac': tf_achieved / tf_peak, 'traffic': None, 'kernel': 'k_nuts',
                     'peak_source': peak_src + ' bf16 sustained; contractions on tcgen05 (bf16 in, fp32 accumulate)',
                     # supplementary view: every gradient evaluation of a site's chains streams the site's bf16
                     # design matrix (64 padded columns) + responses once, from L2 (the tile phase is
                     # bound by shared-memory bandwidth, 3 crossings per tile)
                     'x_stream': {'achieved': (n_leap / chains / max(world, 1)) * n_k * (64 * 2 + 4) / max(samp_s, 1e-9) / 1e9,
                                  'unit': 'GB/s per GPU (L2 -> shared memory)', 'hbm_peak': hbm_peak}},
        'info': int(info), 'mean_stepsize': float(np.mean(msteps)), 'max_rhat': float(np.max(mrhats)),
    }
    if not args.no_cpu_baseline and world == 1:          # (the CPU baseline is reported at N=1 only)
        cores = os.cpu_count() or 1
        n_sample = min(K, max(cores, 2))
        cits, cgps, cwall = cpu_baseline(model, K, n_k, D, chains, siter, n_sample, cores)
        line['cpu_baseline'] = {
            'value': cits, 'unit': 'it/s', 'cores': cores, 'kind': 'port',
            'grad_evals_per_s': cgps,
            'sample': 'one chain on each of %d sites (of %d sites x %d chains), one EP iteration, fp64 NumPy '
                      'oracle NUTS, one process per core (%.1f s); scaled to the full workload'
                      % (n_sample, K, chains, cwall)}
        mom_us, cav_us = cpu_linalg_baseline(d, chains * (siter - siter // 2))
        line['cpu_baseline']['moments_us_per_site'] = mom_us       # one core; the GPU kernels: profiles/*linalg_timing*
        line['cpu_baseline']['cavity_us_per_site'] = cav_us
    print(json.dumps(line))


if __name__ == '__main__':
    main()
