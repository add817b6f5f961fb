// Cavity, damped update + aggregation, global factorisation/moments, PD forcing,
// damping sweep, and the stand-alone invert/olse utilities.
//
// Reference rows: a10 Worker.cavity (method.py:267-302), a11 damped
// update (method.py:1071-1074), a12 PD check / forcing / moments
// (method.py:1077-1133,1160-1219), a13 sweep (find_damp.py:144-174), a7
// invert_normal_params (util.py:51-125), a6 olse (util.py:128-194).
#include "epg_internal.h"
#include "epg_linalg.cuh"

namespace {

__host__ __device__ inline size_t linalg_smem(int d) {
    return sizeof(double) * ((size_t)pk_size(d) + 5 * (size_t)d + 48);
}
// (measured: one warp per matrix with warp-level barriers is slower than these CTA sizes at d = 20..50)
inline int linalg_threads(int d) { return d <= 16 ? 32 : (d <= 32 ? 64 : (d <= 64 ? 128 : (d <= 128 ? 256 : 512))); }

// ------------------------------------------------------------------ cavity
__global__ void k_cavity(const double* __restrict__ Q, const double* __restrict__ r,
                         const double* __restrict__ Qi_all, const double* __restrict__ ri_all,
                         double* __restrict__ cavQ, double* __restrict__ cavm, int* __restrict__ ok,
                         int k0, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* vec = A + pk_size(d);
    const int k = k0 + blockIdx.x;
    const double* Qi = Qi_all + (size_t)k * d * d;
    const double* ri = ri_all + (size_t)k * d;
    double* oQ = cavQ + (size_t)k * d * d;
    for (int e = g.tid; e < d * d; e += g.n) {
        const double v = Q[e] - Qi[e];
        oQ[e] = v;
        const int j = e / d, i = e - j * d;       // column-major (i,j)
        if (j >= i) A[pk(j, i, d)] = v;           // upper triangle -> L' layout (as LAPACK 'U')
    }
    for (int i = g.tid; i < d; i += g.n) vec[i] = r[i] - ri[i];
    g.sync();
    const bool good = chol_packed(g, A, d);
    if (good) {
        fwd_solve_packed(g, A, vec, d);
        bwd_solve_packed(g, A, vec, d);
    }
    for (int i = g.tid; i < d; i += g.n) cavm[(size_t)k * d + i] = vec[i];
    if (g.tid == 0) ok[k] = good ? 1 : 0;
}

// ------------------------------------------------------- damped update (a11)
// grid (tiles over E = d*d + d, site chunks).  part[chunk][E] partial sums.
__global__ void k_update_partial(const double* __restrict__ Qi, const double* __restrict__ dQi,
                                 double* __restrict__ Qi2, const double* __restrict__ ri,
                                 const double* __restrict__ dri, double* __restrict__ ri2,
                                 double* __restrict__ part, double df, int K, int d, int ks) {
    const int E = d * d + d;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int kb = blockIdx.y * ks;
    const int ke = min(K, kb + ks);
    double s = 0.0;
    if (e < d * d) {
        const size_t st = (size_t)d * d;
        for (int k = kb; k < ke; ++k) {
            const double v = Qi[k * st + e] + df * dQi[k * st + e];
            Qi2[k * st + e] = v;
            s += v;
        }
    } else {
        const int i = e - d * d;
        for (int k = kb; k < ke; ++k) {
            const double v = ri[(size_t)k * d + i] + df * dri[(size_t)k * d + i];
            ri2[(size_t)k * d + i] = v;
            s += v;
        }
    }
    part[(size_t)blockIdx.y * E + e] = s;
}

// sums of plain site arrays (used by the sweep)
__global__ void k_site_sum_partial(const double* __restrict__ src, double* __restrict__ part,
                                   int K, int E, int ks) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int kb = blockIdx.y * ks, ke = min(K, kb + ks);
    double s = 0.0;
    for (int k = kb; k < ke; ++k) s += src[(size_t)k * E + e];
    part[(size_t)blockIdx.y * E + e] = s;
}

__global__ void k_sum_chunks(const double* __restrict__ part, double* __restrict__ out, int E, int nchunks) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[(size_t)c * E + e];   // fixed order
    out[e] = s;
}

// ----------------------------------------- global Q = Q0 + sum, Cholesky (a12)
__global__ void k_update_finish(const double* __restrict__ partial, const double* __restrict__ Q0,
                                const double* __restrict__ r0, double* __restrict__ Q,
                                double* __restrict__ r, double* __restrict__ chol, int* __restrict__ flags, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    for (int e = g.tid; e < d * d; e += g.n) {
        const double v = Q0[e] + partial[e];
        Q[e] = v;
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = v;
    }
    for (int i = g.tid; i < d; i += g.n) r[i] = r0[i] + partial[d * d + i];
    g.sync();
    const bool good = chol_packed(g, A, d);
    if (good)
        for (int e = g.tid; e < pk_size(d); e += g.n) chol[e] = A[e];
    if (g.tid == 0) flags[0] = good ? 1 : 0;
}

// ------------------------------- (S, m) from the kept factor (method.py:1215)
__global__ void k_global_moments(const double* __restrict__ chol, const double* __restrict__ r,
                                 double* __restrict__ S, double* __restrict__ m, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* vec = A + pk_size(d);
    double* col = vec + d;
    for (int e = g.tid; e < pk_size(d); e += g.n) A[e] = chol[e];
    for (int i = g.tid; i < d; i += g.n) vec[i] = r[i];
    g.sync();
    fwd_solve_packed(g, A, vec, d);
    bwd_solve_packed(g, A, vec, d);
    for (int i = g.tid; i < d; i += g.n) m[i] = vec[i];
    trtri_packed(g, A, col, d);
    lauum_full(g, A, d, 1.0, S);
}

// ----------------------------------------------- PD forcing (method.py:1119-1132)
__global__ void k_force_pd(const double* __restrict__ Qi2, double* __restrict__ Qi,
                           double* __restrict__ lam_out, int* __restrict__ forced, double thr,
                           double min_eig, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* dg = A + pk_size(d);
    double* od = dg + d;
    double* v = od + d;
    double* w = v + d;
    double* red = w + d;
    const int k = blockIdx.x;
    const double* src = Qi2 + (size_t)k * d * d;
    for (int e = g.tid; e < d * d; e += g.n) {
        const int j = e / d, i = e - j * d;
        if (i >= j) A[pk(i, j, d)] = src[e];      // LAPACK dsyevr default reads the lower triangle
    }
    g.sync();
    const double lam = min_eig_packed(g, A, d, dg, od, v, w, red);
    const bool f = lam < thr;
    if (f) {
        double* dst = Qi + (size_t)k * d * d;
        for (int i = g.tid; i < d; i += g.n) dst[i + (size_t)i * d] += min_eig - lam;
    }
    if (g.tid == 0) { lam_out[k] = lam; forced[k] = f ? 1 : 0; }
}

// ------------------------------------------------------------ damping sweep
// per damping value: global approximation, its Cholesky, mean, MSE and KL
// sums layout: [SQi (d*d) | SdQi (d*d) | Sri (d) | Sdri (d)]
// tgt layout:  [m_tgt (d) | S_tgt (d*d)]
// out layout:  [mse (n_df) | kl (n_df) | okflag (n_df)]
__global__ void k_sweep_global(const double* __restrict__ sums, const double* __restrict__ Q0,
                               const double* __restrict__ r0, const double* __restrict__ dfs,
                               const double* __restrict__ tgt, double* __restrict__ out,
                               double* __restrict__ Qdf, double* __restrict__ rdf, int n_df, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* vec = A + pk_size(d);
    double* dm = vec + d;
    double* red = dm + 4 * d;
    const int q = blockIdx.x;
    const double df = dfs[q];
    const double* m_t = tgt;
    const double* S_t = tgt + d;
    const double* SQi = sums;
    const double* SdQi = sums + (size_t)d * d;
    const double* Sri = sums + 2 * (size_t)d * d;
    const double* Sdri = Sri + d;
    double* Qg = Qdf + (size_t)q * d * d;
    // sum log diag chol(S_tgt)
    for (int e = g.tid; e < d * d; e += g.n) {
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = S_t[e];
    }
    g.sync();
    bool good = chol_packed(g, A, d);
    double part = 0.0;
    if (good) for (int i = g.tid; i < d; i += g.n) part += log(A[pk(i, i, d)]);
    const double ld0 = block_sum(g, part, red);
    g.sync();
    // Q(df), r(df)
    for (int e = g.tid; e < d * d; e += g.n) {
        const double v = Q0[e] + SQi[e] + df * SdQi[e];
        Qg[e] = v;
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = v;
    }
    for (int i = g.tid; i < d; i += g.n) {
        const double v = r0[i] + Sri[i] + df * Sdri[i];
        vec[i] = v;
        rdf[(size_t)q * d + i] = v;
    }
    g.sync();
    good = chol_packed(g, A, d) && good;
    double mse = NAN, kl = NAN;
    if (good) {
        part = 0.0;
        for (int i = g.tid; i < d; i += g.n) part += log(A[pk(i, i, d)]);
        const double ldq = block_sum(g, part, red);      // = -sum log diag chol(S_approx)
        fwd_solve_packed(g, A, vec, d);
        bwd_solve_packed(g, A, vec, d);
        part = 0.0;
        for (int i = g.tid; i < d; i += g.n) {
            const double e = vec[i] - m_t[i];
            dm[i] = e;
            part += e * e;
        }
        mse = block_sum(g, part, red) / d;
        __threadfence_block();
        g.sync();
        // tr(S1^-1 S0) + dm' S1^-1 dm with S1^-1 = Q(df)
        part = 0.0;
        for (int e = g.tid; e < d * d; e += g.n) {
            const int j = e / d, i = e - j * d;
            part += Qg[e] * (S_t[e] + dm[i] * dm[j]);
        }
        const double trq = block_sum(g, part, red);
        kl = 0.5 * (trq - d) - ld0 - ldq;
    }
    if (g.tid == 0) { out[q] = mse; out[n_df + q] = kl; out[2 * n_df + q] = good ? 1.0 : 0.0; }
}

// cavity check for every (site, damping value): grid (K, n_df)
__global__ void k_sweep_cavity(const double* __restrict__ Qdf, const double* __restrict__ Qi,
                               const double* __restrict__ dQi, const double* __restrict__ dfs,
                               int* __restrict__ okdf, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    const int k = blockIdx.x, q = blockIdx.y;
    const double df = dfs[q];
    const double* Qg = Qdf + (size_t)q * d * d;
    const double* a = Qi + (size_t)k * d * d;
    const double* b = dQi + (size_t)k * d * d;
    for (int e = g.tid; e < d * d; e += g.n) {
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = Qg[e] - (a[e] + df * b[e]);
    }
    g.sync();
    if (!chol_packed(g, A, d) && g.tid == 0) atomicAnd(&okdf[q], 0);
}

__global__ void k_sweep_final(double* __restrict__ out, const int* __restrict__ okdf, int n_df) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_df) return;
    if (out[2 * n_df + q] == 0.0 || okdf[q] == 0) { out[q] = NAN; out[n_df + q] = NAN; }
}

// ------------------------------------------- util.invert_normal_params (a7)
__global__ void k_invert(const double* __restrict__ A_in, const double* __restrict__ b_in, int cho_form,
                         double* __restrict__ outA, double* __restrict__ outb, int* __restrict__ ok, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* vec = A + pk_size(d);
    double* col = vec + d;
    const int b = blockIdx.x;
    const double* src = A_in + (size_t)b * d * d;
    for (int e = g.tid; e < d * d; e += g.n) {
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = src[e];      // upper triangle (factor U or matrix), L = U'
    }
    if (b_in) for (int i = g.tid; i < d; i += g.n) vec[i] = b_in[(size_t)b * d + i];
    g.sync();
    bool good = true;
    if (!cho_form) good = chol_packed(g, A, d);
    else {
        // dpotri reports an exactly zero diagonal of the factor as singular
        double bad = 0.0;
        for (int i = g.tid; i < d; i += g.n) if (A[pk(i, i, d)] == 0.0) bad = 1.0;
        double* red = col + d;
        good = block_sum(g, bad, red) == 0.0;
    }
    if (good) {
        if (b_in) {
            fwd_solve_packed(g, A, vec, d);
            bwd_solve_packed(g, A, vec, d);
            for (int i = g.tid; i < d; i += g.n) outb[(size_t)b * d + i] = vec[i];
        }
        trtri_packed(g, A, col, d);
        lauum_full(g, A, d, 1.0, outA + (size_t)b * d * d);
    }
    if (g.tid == 0) ok[b] = good ? 1 : 0;
}

// ------------------------------------------------------------ util.olse (a6)
__global__ void k_olse(const double* __restrict__ S_in, int n, const double* __restrict__ P_in,
                       double* __restrict__ out_all, int* __restrict__ ok, int d) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* col = A + pk_size(d);
    double* red = col + d;
    const int b = blockIdx.x;
    const double* src = S_in + (size_t)b * d * d;
    const double* P = P_in ? P_in + (size_t)b * d * d : nullptr;
    double* out = out_all + (size_t)b * d * d;
    for (int e = g.tid; e < d * d; e += g.n) {
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = src[e];
    }
    g.sync();
    const bool good = chol_packed(g, A, d);
    if (good) {
        trtri_packed(g, A, col, d);
        lauum_full(g, A, d, 1.0, out);
        __threadfence_block();
        g.sync();
        double ptr = 0.0, pf2 = 0.0, pf2p = 0.0, psp = 0.0;
        for (int e = g.tid; e < d * d; e += g.n) {
            const double s = out[e];
            pf2 += s * s;
            if (e % (d + 1) == 0) ptr += s;
            if (P) { const double p = P[e]; pf2p += p * p; psp += s * p; }
        }
        const double tr = block_sum(g, ptr, red);
        const double f2 = block_sum(g, pf2, red);
        const double dn = (double)d / (double)n;
        if (P) {
            const double f2p = block_sum(g, pf2p, red);
            const double trSP = block_sum(g, psp, red);
            const double alpha = 1.0 - ((double)d + tr * tr * f2p / (f2 * f2p - trSP * trSP)) / (double)n;
            const double beta = (trSP / f2p) * (1.0 - dn - alpha);
            for (int e = g.tid; e < d * d; e += g.n) out[e] = alpha * out[e] + beta * P[e];
        } else {
            const double alpha = 1.0 - ((double)d + tr * tr / (f2 - tr * tr / d)) / (double)n;
            const double beta = tr * (1.0 - dn - alpha);
            for (int e = g.tid; e < d * d; e += g.n)
                out[e] = alpha * out[e] + ((e % (d + 1) == 0) ? beta / d : 0.0);
        }
    }
    if (g.tid == 0) ok[b] = good ? 1 : 0;
}

inline cudaError_t set_smem(const void* f, size_t bytes) {
    return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

inline void chunking(int K, int E, int& nchunks, int& ks) {
    long want = (148L * 8 * 256) / (E > 0 ? E : 1);
    if (want < 1) want = 1;
    if (want > K) want = K;
    ks = (int)((K + want - 1) / want);
    nchunks = (K + ks - 1) / ks;
}

}  // namespace

#define LAUNCH_CHECK(c)            \
    do {                           \
        (c)->launches++;           \
        cudaError_t _le = cudaGetLastError(); \
        if (_le != cudaSuccess) return _le;   \
    } while (0)

cudaError_t epg_launch_cavity(epg_ctx* c, int k0, int k1, int proposal) {
    const int d = c->d;
    const size_t sm = linalg_smem(d);
    cudaError_t e = set_smem((const void*)k_cavity, sm);
    if (e != cudaSuccess) return e;
    k_cavity<<<k1 - k0, linalg_threads(d), sm, c->stream>>>(
        c->arr[EPG_Q], c->arr[EPG_R], c->arr[proposal ? EPG_QI2 : EPG_QI],
        c->arr[proposal ? EPG_RI2 : EPG_RI], c->arr[EPG_CAVQ], c->arr[EPG_CAVM], c->site_ok, k0, d);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

cudaError_t epg_launch_update_partial(epg_ctx* c, double df) {
    const int d = c->d, K = c->K, E = d * d + d;
    int nchunks, ks;
    chunking(K, E, nchunks, ks);
    const size_t need = sizeof(double) * (size_t)nchunks * E;
    cudaError_t e = epg_reserve((void**)&c->scratch, &c->scratch_bytes, need);
    if (e != cudaSuccess) return e;
    dim3 grid((E + 255) / 256, nchunks);
    k_update_partial<<<grid, 256, 0, c->stream>>>(c->arr[EPG_QI], c->arr[EPG_DQI], c->arr[EPG_QI2],
                                                  c->arr[EPG_RI], c->arr[EPG_DRI], c->arr[EPG_RI2],
                                                  c->scratch, df, K, d, ks);
    LAUNCH_CHECK(c);
    k_sum_chunks<<<(E + 255) / 256, 256, 0, c->stream>>>(c->scratch, c->arr[EPG_PARTIAL], E, nchunks);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

cudaError_t epg_launch_update_finish(epg_ctx* c) {
    const int d = c->d;
    const size_t sm = linalg_smem(d);
    cudaError_t e = set_smem((const void*)k_update_finish, sm);
    if (e != cudaSuccess) return e;
    k_update_finish<<<1, linalg_threads(d), sm, c->stream>>>(c->arr[EPG_PARTIAL], c->arr[EPG_Q0], c->arr[EPG_R0],
                                                             c->arr[EPG_Q], c->arr[EPG_R], c->chol, c->flags, d);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

cudaError_t epg_launch_global_moments(epg_ctx* c) {
    const int d = c->d;
    const size_t sm = linalg_smem(d);
    cudaError_t e = set_smem((const void*)k_global_moments, sm);
    if (e != cudaSuccess) return e;
    k_global_moments<<<1, linalg_threads(d), sm, c->stream>>>(c->chol, c->arr[EPG_R], c->arr[EPG_S],
                                                              c->arr[EPG_M], d);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

cudaError_t epg_launch_force_pd(epg_ctx* c, double thr, double min_eig, double* lam_dev) {
    const int d = c->d;
    const size_t sm = linalg_smem(d);
    cudaError_t e = set_smem((const void*)k_force_pd, sm);
    if (e != cudaSuccess) return e;
    k_force_pd<<<c->K, linalg_threads(d), sm, c->stream>>>(c->arr[EPG_QI2], c->arr[EPG_QI], lam_dev,
                                                           c->site_ok, thr, min_eig, d);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

// scratch layout for the sweep (doubles):
//   sums[2dd+2d] | Qdf[n_df*dd] | rdf[n_df*d] | okdf (ints, n_df) | chunk partials
cudaError_t epg_launch_damp_sweep(epg_ctx* c, int n_df, const double* dfs_dev, const double* tgt_dev,
                                  double* out_dev) {
    const int d = c->d, K = c->K;
    const size_t dd = (size_t)d * d;
    int nchunks, ks;
    chunking(K, (int)dd, nchunks, ks);
    const size_t n_sums = 2 * dd + 2 * d;
    const size_t n_q = (size_t)n_df * dd, n_r = (size_t)n_df * d;
    const size_t n_ok = (n_df + 1) / 2 + 1;
    const size_t n_part = (size_t)nchunks * dd;
    const size_t need = sizeof(double) * (n_sums + n_q + n_r + n_ok + n_part);
    cudaError_t e = epg_reserve((void**)&c->scratch, &c->scratch_bytes, need);
    if (e != cudaSuccess) return e;
    double* sums = c->scratch;
    double* Qdf = sums + n_sums;
    double* rdf = Qdf + n_q;
    int* okdf = reinterpret_cast<int*>(rdf + n_r);
    double* part = rdf + n_r + n_ok;
    struct { const double* src; double* dst; int E; } jobs[4] = {
        {c->arr[EPG_QI], sums, (int)dd}, {c->arr[EPG_DQI], sums + dd, (int)dd},
        {c->arr[EPG_RI], sums + 2 * dd, d}, {c->arr[EPG_DRI], sums + 2 * dd + d, d}};
    for (auto& jb : jobs) {
        dim3 grid((jb.E + 255) / 256, nchunks);
        k_site_sum_partial<<<grid, 256, 0, c->stream>>>(jb.src, part, K, jb.E, ks);
        LAUNCH_CHECK(c);
        k_sum_chunks<<<(jb.E + 255) / 256, 256, 0, c->stream>>>(part, jb.dst, jb.E, nchunks);
        LAUNCH_CHECK(c);
    }
    e = cudaMemsetAsync(okdf, 0xff, sizeof(int) * n_df, c->stream);
    if (e != cudaSuccess) return e;
    const size_t sm = linalg_smem(d);
    if ((e = set_smem((const void*)k_sweep_global, sm)) != cudaSuccess) return e;
    if ((e = set_smem((const void*)k_sweep_cavity, sm)) != cudaSuccess) return e;
    k_sweep_global<<<n_df, linalg_threads(d), sm, c->stream>>>(sums, c->arr[EPG_Q0], c->arr[EPG_R0], dfs_dev,
                                                               tgt_dev, out_dev, Qdf, rdf, n_df, d);
    LAUNCH_CHECK(c);
    k_sweep_cavity<<<dim3(K, n_df), linalg_threads(d), sm, c->stream>>>(Qdf, c->arr[EPG_QI], c->arr[EPG_DQI],
                                                                        dfs_dev, okdf, d);
    LAUNCH_CHECK(c);
    k_sweep_final<<<(n_df + 63) / 64, 64, 0, c->stream>>>(out_dev, okdf, n_df);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

cudaError_t epg_launch_invert(epg_ctx* c, int batch, int d, const double* A, const double* b, int cho_form,
                              double* outA, double* outb, int* ok) {
    const size_t sm = linalg_smem(d);
    cudaError_t e = set_smem((const void*)k_invert, sm);
    if (e != cudaSuccess) return e;
    k_invert<<<batch, linalg_threads(d), sm, c->stream>>>(A, b, cho_form, outA, outb, ok, d);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}

cudaError_t epg_launch_olse(epg_ctx* c, int batch, int d, const double* S, int n, const double* P, double* out,
                            int* ok) {
    const size_t sm = linalg_smem(d);
    cudaError_t e = set_smem((const void*)k_olse, sm);
    if (e != cudaSuccess) return e;
    k_olse<<<batch, linalg_threads(d), sm, c->stream>>>(S, n, P, out, ok, d);
    LAUNCH_CHECK(c);
    return cudaSuccess;
}
