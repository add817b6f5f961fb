// Control-variate moment estimation, batched over sites.
//
// Replaces util.cv_moments / _cv_estim (reference epstan/util.py:197-411) and the
// Cython helpers it leans on (auto_outer pyx:45-81, ravel_triu pyx:111-143,
// unravel_triu pyx:148-183).  The reference materialises the n x d2 feature
// matrices f and h (d2 = d(d+1)/2); here every feature value is formed on the
// fly from the draws held in shared memory, so HBM sees the draws once per
// output tile and the d2 x d2 Gram matrices are the only large intermediates.
//
// Per item:  k_cv_prep   Cholesky of Q_tilde -> (S_tilde, m_tilde, log-normaliser),
//                        probability ratios pr_i, the m_treshold guard
//            k_cv_sums   feature means          (stage 0: mean, stage 1: covariance)
//            k_cv_gram   hc'hc and hc'fc tiles  (multiple_cv)  |  k_cv_diag (single)
//            k_cv_solve  a = (hc'hc * var_k)^-1 (hc'fc * cov_k) by LU with partial
//                        pivoting (the reference calls scipy.linalg.solve == dgesv),
//                        regulate / clip a, f_hat = mean(f) - mean(hc) a
#include "epg_internal.h"
#include "epg_linalg.cuh"
#include <vector>

namespace {

struct CvArgs {
    int n, d, d2, batch0;
    const double* draws;     // [batch][d][n]
    const double* lp;        // [batch][n]
    const double* Qt;        // [batch][d*d]
    const double* rt;        // [batch][d]
    const int2* feat;        // [d2] (a,b), a<=b, np.triu_indices order
    double* St;              // [batch][d*d]  S_tilde
    double* mt;              // [batch][d]    m_tilde
    double* pr;              // [batch][n]
    double* mhat;            // [batch][d]
    double* shat;            // [batch][d*d]
    int* status;             // [batch] 1 cv, 0 fallback, -1 failed
    // per-stage work areas (nf = d or d2)
    double* fm;              // [batch][d2] feature means of f
    double* hm;              // [batch][d2] mean(hc)
    double* G1;              // [batch][nf*nf]
    double* G2;              // [batch][nf*nf]
    double* res;             // [batch][d2] stage result
    double* aout;            // coefficients a of the stage (ret_a): [batch][nf*nf] (multiple_cv) or [batch][nf]; may be null
    int multiple_cv;
    double regulate_a, max_a, m_treshold;
};

// feature values of one draw held in xs[dim] (stage 0: linear, stage 1: triu outer product)
__device__ __forceinline__ void feature(const CvArgs& a, int stage, int e, const double* xs, int ld, int t,
                                        const double* cen_f, const double* cen_h, double prv, double& f, double& h) {
    if (stage == 0) {
        const double x = xs[e * ld + t];
        f = x;
        h = x * prv;
    } else {
        const int2 ab = a.feat[e];
        const double xa = xs[ab.x * ld + t], xb = xs[ab.y * ld + t];
        f = (xa - cen_f[ab.x]) * (xb - cen_f[ab.y]);
        h = (xa - cen_h[ab.x]) * (xb - cen_h[ab.y]) * prv;
    }
}
__device__ __forceinline__ double feature_Eh(const CvArgs& a, int stage, int b, int e) {
    if (stage == 0) return a.mt[(size_t)b * a.d + e];
    const int2 ab = a.feat[e];
    return a.St[(size_t)b * a.d * a.d + ab.x + (size_t)ab.y * a.d];
}

// ---------------------------------------------------------------- preparation
__global__ void k_cv_prep(const CvArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    const int d = a.d, n = a.n, b = a.batch0 + blockIdx.x;
    double* A = reinterpret_cast<double*>(smem_raw);
    double* vec = A + pk_size(d);
    double* col = vec + d;
    double* red = col + d;
    double* cnt = red + 40;          // d counters
    const double* Q = a.Qt + (size_t)b * d * d;
    for (int e = g.tid; e < d * d; e += g.n) {
        const int j = e / d, i = e - j * d;
        if (j >= i) A[pk(j, i, d)] = Q[e];
    }
    for (int i = g.tid; i < d; i += g.n) { vec[i] = a.rt[(size_t)b * d + i]; cnt[i] = 0.0; }
    g.sync();
    if (!chol_packed(g, A, d)) {
        if (g.tid == 0) a.status[b] = -1;
        return;
    }
    double part = 0.0;
    for (int i = g.tid; i < d; i += g.n) part += log(A[pk(i, i, d)]);
    const double cst = block_sum(g, part, red) - 0.5 * d * 1.8378770664093453;   // log(2 pi)
    fwd_solve_packed(g, A, vec, d);
    bwd_solve_packed(g, A, vec, d);
    for (int i = g.tid; i < d; i += g.n) a.mt[(size_t)b * d + i] = vec[i];
    trtri_packed(g, A, col, d);
    lauum_full(g, A, d, 1.0, a.St + (size_t)b * d * d);
    g.sync();
    // probability ratios and the treshold counts
    const double* x = a.draws + (size_t)b * d * n;
    for (int t = g.tid; t < n; t += g.n) {
        double quad = 0.0;
        for (int i = 0; i < d; ++i) {
            const double di = x[(size_t)i * n + t] - vec[i];
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += Q[i + (size_t)j * d] * (x[(size_t)j * n + t] - vec[j]);
            quad += di * s;
        }
        a.pr[(size_t)b * n + t] = exp(cst - 0.5 * quad - a.lp[(size_t)b * n + t]);
    }
    int use_cv = 1;
    if (a.m_treshold > 0.0) {
        const double thr = a.m_treshold < 0.5 ? 1.0 - a.m_treshold : a.m_treshold;
        double bad = 0.0;
        for (int i = g.tid; i < d; i += g.n) {
            int c = 0;
            for (int t = 0; t < n; ++t) c += x[(size_t)i * n + t] < vec[i];
            const double ratio = (double)c / n;
            if (ratio > thr || ratio < 1.0 - thr) bad = 1.0;
        }
        if (block_sum(g, bad, red) > 0.0) use_cv = 0;
    }
    if (g.tid == 0) a.status[b] = use_cv;
}

// plain sample estimates for the items that failed the treshold test (util.py:359-367)
__global__ void k_cv_fallback(const CvArgs a) {
    const int b = a.batch0 + blockIdx.x, d = a.d, n = a.n;
    if (a.status[b] != 0) return;
    const double* x = a.draws + (size_t)b * d * n;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* mean = reinterpret_cast<double*>(smem_raw);
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < n; ++t) s += x[(size_t)i * n + t];
        mean[i] = s / n;
        a.mhat[(size_t)b * d + i] = s / n;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
        const int j = e / d, i = e - j * d;
        double s = 0.0;
        for (int t = 0; t < n; ++t) s += (x[(size_t)i * n + t] - mean[i]) * (x[(size_t)j * n + t] - mean[j]);
        a.shat[(size_t)b * d * d + e] = s / (n - 1);
    }
}

// -------------------------------------------------------------- feature means
// grid (ceil(nf/128), batch); block 128: one feature per thread, draws in chunks via smem
__global__ void k_cv_sums(const CvArgs a, int stage) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d, n = a.n, b = a.batch0 + blockIdx.y;
    if (a.status[b] != 1) return;
    const int nf = stage == 0 ? d : a.d2;
    const int T = 32;
    double* xs = reinterpret_cast<double*>(smem_raw);     // [d][T]
    double* prs = xs + d * T;                             // [T]
    double* cf = prs + T;                                 // centre of f (m_hat)
    double* ch = cf + d;                                  // centre of h (m_tilde)
    const double* x = a.draws + (size_t)b * d * n;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        cf[i] = stage ? a.mhat[(size_t)b * d + i] : 0.0;
        ch[i] = a.mt[(size_t)b * d + i];
    }
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double sf = 0.0, sh = 0.0;
    for (int c0 = 0; c0 < n; c0 += T) {
        const int tn = min(T, n - c0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < d * T; idx += blockDim.x) {
            const int t = idx % T, i = idx / T;
            xs[i * T + t] = t < tn ? x[(size_t)i * n + c0 + t] : 0.0;
        }
        for (int t = threadIdx.x; t < T; t += blockDim.x) prs[t] = t < tn ? a.pr[(size_t)b * n + c0 + t] : 0.0;
        __syncthreads();
        if (e < nf)
            for (int t = 0; t < tn; ++t) {
                double f, h;
                feature(a, stage, e, xs, T, t, cf, ch, prs[t], f, h);
                sf += f; sh += h;
            }
    }
    if (e < nf) {
        const int ddof_f = stage ? 1 : 0;
        a.fm[(size_t)b * a.d2 + e] = sf / (n - ddof_f);
        a.hm[(size_t)b * a.d2 + e] = sh / n - feature_Eh(a, stage, b, e);
    }
}

// -------------------------------------------------------------------- Gram tiles
// grid (tiles, tiles, batch): G1[e1][e2] = sum hc_e1 hc_e2 ; G2[e1][e2] = sum hc_e1 fc_e2
#define CV_TILE 64
#define CV_T 16
__global__ void __launch_bounds__(256) k_cv_gram(const CvArgs a, int stage) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d, n = a.n, b = a.batch0 + blockIdx.z;
    if (a.status[b] != 1) return;
    const int nf = stage == 0 ? d : a.d2;
    double* xs = reinterpret_cast<double*>(smem_raw);     // [d][CV_T]
    double* prs = xs + d * CV_T;                          // [CV_T]
    double* cf = prs + CV_T;                              // [d]
    double* ch = cf + d;                                  // [d]
    double* hr = ch + d;                                  // [CV_T][CV_TILE]  hc of the row features
    double* hcn = hr + CV_T * CV_TILE;                    // [CV_T][CV_TILE]  hc of the column features
    double* fcn = hcn + CV_T * CV_TILE;                   // [CV_T][CV_TILE]  fc of the column features
    const int tid = threadIdx.x;
    const int e1_0 = blockIdx.y * CV_TILE, e2_0 = blockIdx.x * CV_TILE;
    const double* x = a.draws + (size_t)b * d * n;
    for (int i = tid; i < d; i += 256) {
        cf[i] = stage ? a.mhat[(size_t)b * d + i] : 0.0;
        ch[i] = a.mt[(size_t)b * d + i];
    }
    // this thread's fixed feature for the fill phase
    const int fe = tid & 63, fwhich = tid >> 6;           // 0: rows h, 1: cols h+f, 2,3: help with t split
    const int ty = tid >> 4, tx = tid & 15;               // 16 x 16 threads, 4 x 4 outputs each
    double acc1[16], acc2[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) { acc1[q] = 0.0; acc2[q] = 0.0; }
    for (int c0 = 0; c0 < n; c0 += CV_T) {
        const int tn = min(CV_T, n - c0);
        __syncthreads();
        for (int idx = tid; idx < d * CV_T; idx += 256) {
            const int t = idx % CV_T, i = idx / CV_T;
            xs[i * CV_T + t] = t < tn ? x[(size_t)i * n + c0 + t] : 0.0;
        }
        if (tid < CV_T) prs[tid] = tid < tn ? a.pr[(size_t)b * n + c0 + tid] : 0.0;
        __syncthreads();
        // fill: 4 thread groups of 64 features; group g handles draws t = g, g+4, ... for rows and cols
        {
            const int er = e1_0 + fe, ec = e2_0 + fe;
            const double Ehr = er < nf ? feature_Eh(a, stage, b, er) : 0.0;
            const double Ehc = ec < nf ? feature_Eh(a, stage, b, ec) : 0.0;
            const double fmc = ec < nf ? a.fm[(size_t)b * a.d2 + ec] : 0.0;
            for (int t = fwhich; t < CV_T; t += 4) {
                double f = 0.0, h = 0.0, vr = 0.0, vh = 0.0, vf = 0.0;
                if (t < tn) {
                    if (er < nf) { feature(a, stage, er, xs, CV_T, t, cf, ch, prs[t], f, h); vr = h - Ehr; }
                    if (ec < nf) { feature(a, stage, ec, xs, CV_T, t, cf, ch, prs[t], f, h); vh = h - Ehc; vf = f - fmc; }
                }
                hr[t * CV_TILE + fe] = vr;
                hcn[t * CV_TILE + fe] = vh;
                fcn[t * CV_TILE + fe] = vf;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < CV_T; ++t) {
            double rv[4], c1[4], c2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                rv[q] = hr[t * CV_TILE + 4 * ty + q];
                c1[q] = hcn[t * CV_TILE + 4 * tx + q];
                c2[q] = fcn[t * CV_TILE + 4 * tx + q];
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc1[p * 4 + q] = fma(rv[p], c1[q], acc1[p * 4 + q]);
                    acc2[p * 4 + q] = fma(rv[p], c2[q], acc2[p * 4 + q]);
                }
        }
    }
    double* G1 = a.G1 + (size_t)blockIdx.z * nf * nf;
    double* G2 = a.G2 + (size_t)blockIdx.z * nf * nf;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e1 = e1_0 + 4 * ty + p, e2 = e2_0 + 4 * tx + q;
            if (e1 < nf && e2 < nf) {
                G1[(size_t)e1 * nf + e2] = acc1[p * 4 + q];
                G2[(size_t)e1 * nf + e2] = acc2[p * 4 + q];
            }
        }
}

// ------------------------------------------------- solve + estimate (multiple_cv)
// one CTA per item; G1 (row-major nf x nf) * var_k, G2 * cov_k; a = G1^-1 G2 by LU
// with partial pivoting in place in global memory; res_j = fm_j - sum_i hm_i a_ij
__global__ void __launch_bounds__(1024) k_cv_solve(const CvArgs a, int stage) {
    const int b = a.batch0 + blockIdx.x;
    if (a.status[b] != 1) return;
    const int nf = stage == 0 ? a.d : a.d2;
    const double nn = (double)a.n;
    const double var_k = stage == 0 ? nn - 1.0 : (nn - 1.0) * (nn - 1.0);
    const double cov_k = stage == 0 ? nn : nn * nn;
    double* A = a.G1 + (size_t)blockIdx.x * nf * nf;
    double* B = a.G2 + (size_t)blockIdx.x * nf * nf;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ int s_fail;
    for (size_t e = tid; e < (size_t)nf * nf; e += nt) { A[e] *= var_k; B[e] *= cov_k; }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int k = 0; k < nf; ++k) {
        // pivot search in column k
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < nf; i += nt) {
            const double v = fabs(A[(size_t)i * nf + k]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_val[tid >> 5] = best; s_idx[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double bv = s_val[0]; int bb = s_idx[0];
            for (int w = 1; w < (nt + 31) / 32; ++w)
                if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bb)) { bv = s_val[w]; bb = s_idx[w]; }
            s_piv = bb;
            if (!(bv > 0.0) || !isfinite(bv)) s_fail = 1;
        }
        __syncthreads();
        if (s_fail) break;
        const int p = s_piv;
        if (p != k) {
            for (int j = tid; j < 2 * nf; j += nt) {
                double* M = j < nf ? A : B;
                const int jj = j < nf ? j : j - nf;
                const double t = M[(size_t)k * nf + jj];
                M[(size_t)k * nf + jj] = M[(size_t)p * nf + jj];
                M[(size_t)p * nf + jj] = t;
            }
            __syncthreads();
        }
        const double pivinv = 1.0 / A[(size_t)k * nf + k];
        // eliminate: rows i>k, columns (k+1..nf) of A and all of B
        const int wA = nf - k - 1, wtot = wA + nf;
        const size_t work = (size_t)(nf - k - 1) * wtot;
        for (size_t e = tid; e < work; e += nt) {
            const int i = k + 1 + (int)(e / wtot);
            const int c = (int)(e % wtot);
            const double l = A[(size_t)i * nf + k] * pivinv;
            if (c < wA) A[(size_t)i * nf + k + 1 + c] -= l * A[(size_t)k * nf + k + 1 + c];
            else B[(size_t)i * nf + (c - wA)] -= l * B[(size_t)k * nf + (c - wA)];
        }
        __syncthreads();
    }
    if (s_fail) {
        if (tid == 0) a.status[b] = -1;
        return;
    }
    // back substitution, one right-hand-side column per thread
    for (int j = tid; j < nf; j += nt) {
        for (int k = nf - 1; k >= 0; --k) {
            double s = B[(size_t)k * nf + j];
            for (int i = k + 1; i < nf; ++i) s -= A[(size_t)k * nf + i] * B[(size_t)i * nf + j];
            B[(size_t)k * nf + j] = s / A[(size_t)k * nf + k];
        }
    }
    __syncthreads();
    // regulate / clip a (util.py:227-230), keep a copy for ret_a
    double* aout = a.aout ? a.aout + (size_t)b * nf * nf : nullptr;
    for (size_t e = tid; e < (size_t)nf * nf; e += nt) {
        double aij = B[e];
        if (a.regulate_a > 0.0) aij *= a.regulate_a;
        if (a.max_a > 0.0) aij = fmin(fmax(aij, -a.max_a), a.max_a);
        B[e] = aij;
        if (aout) aout[e] = aij;
    }
    __syncthreads();
    const double* hm = a.hm + (size_t)b * a.d2;
    const double* fm = a.fm + (size_t)b * a.d2;
    for (int j = tid; j < nf; j += nt) {
        double s = 0.0;
        for (int i = 0; i < nf; ++i) s += hm[i] * B[(size_t)i * nf + j];
        a.res[(size_t)b * a.d2 + j] = fm[j] - s;
    }
}

// single control variate per dimension (multiple_cv = False): util.py:219-225,241
__global__ void k_cv_diag(const CvArgs a, int stage) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d, n = a.n, b = a.batch0 + blockIdx.y;
    if (a.status[b] != 1) return;
    const int nf = stage == 0 ? d : a.d2;
    const int T = 32;
    double* xs = reinterpret_cast<double*>(smem_raw);
    double* prs = xs + d * T;
    double* cf = prs + T;
    double* ch = cf + d;
    const double* x = a.draws + (size_t)b * d * n;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        cf[i] = stage ? a.mhat[(size_t)b * d + i] : 0.0;
        ch[i] = a.mt[(size_t)b * d + i];
    }
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const double Eh = e < nf ? feature_Eh(a, stage, b, e) : 0.0;
    const double fme = e < nf ? a.fm[(size_t)b * a.d2 + e] : 0.0;
    double vh = 0.0, cfh = 0.0;
    for (int c0 = 0; c0 < n; c0 += T) {
        const int tn = min(T, n - c0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < d * T; idx += blockDim.x) {
            const int t = idx % T, i = idx / T;
            xs[i * T + t] = t < tn ? x[(size_t)i * n + c0 + t] : 0.0;
        }
        for (int t = threadIdx.x; t < T; t += blockDim.x) prs[t] = t < tn ? a.pr[(size_t)b * n + c0 + t] : 0.0;
        __syncthreads();
        if (e < nf)
            for (int t = 0; t < tn; ++t) {
                double f, h;
                feature(a, stage, e, xs, T, t, cf, ch, prs[t], f, h);
                vh += (h - Eh) * (h - Eh);
                cfh += (f - fme) * (h - Eh);
            }
    }
    if (e < nf) {
        const double nn = (double)n;
        const double var_k = stage == 0 ? nn - 1.0 : (nn - 1.0) * (nn - 1.0);
        const double cov_k = stage == 0 ? nn : nn * nn;
        double av = (cfh * cov_k) / (vh * var_k);
        if (a.regulate_a > 0.0) av *= a.regulate_a;
        if (a.max_a > 0.0) av = fmin(fmax(av, -a.max_a), a.max_a);
        a.res[(size_t)b * a.d2 + e] = fme - a.hm[(size_t)b * a.d2 + e] * av;
        if (a.aout) a.aout[(size_t)b * nf + e] = av;
    }
}

// stage results -> m_hat / S_hat (unravel_triu)
__global__ void k_cv_store(const CvArgs a, int stage) {
    const int b = a.batch0 + blockIdx.y;
    if (a.status[b] != 1) return;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = stage == 0 ? a.d : a.d2;
    if (e >= nf) return;
    const double v = a.res[(size_t)b * a.d2 + e];
    if (stage == 0) a.mhat[(size_t)b * a.d + e] = v;
    else {
        const int2 ab = a.feat[e];
        a.shat[(size_t)b * a.d * a.d + ab.x + (size_t)ab.y * a.d] = v;
        a.shat[(size_t)b * a.d * a.d + ab.y + (size_t)ab.x * a.d] = v;
    }
}

}  // namespace

extern "C" int epg_cv_moments(epg_ctx* c, int batch, int n, int d, const double* draws, const double* lp,
                              const double* Q_tilde, const double* r_tilde, int multiple_cv, double regulate_a,
                              double max_a, double m_treshold, double* S_hat, double* m_hat, int32_t* used_cv) {
    return epg_cv_moments_ex(c, batch, n, d, draws, lp, Q_tilde, r_tilde, multiple_cv, regulate_a, max_a, m_treshold,
                             S_hat, m_hat, used_cv, nullptr, nullptr);
}

extern "C" int epg_cv_moments_ex(epg_ctx* c, int batch, int n, int d, const double* draws, const double* lp,
                                 const double* Q_tilde, const double* r_tilde, int multiple_cv, double regulate_a,
                                 double max_a, double m_treshold, double* S_hat, double* m_hat, int32_t* used_cv,
                                 double* a_S_out, double* a_m_out) {
    if (batch < 1 || n < 2 || d < 1 || d > 200 || !draws || !lp || !Q_tilde || !r_tilde || !S_hat || !m_hat)
        return epg_fail_msg(c, "epg_cv_moments: bad args");
    const int d2 = d * (d + 1) / 2;
    if (multiple_cv && d2 > 4096) return epg_fail_msg(c, "epg_cv_moments: d(d+1)/2 > 4096 with multiple_cv is not supported");
    const size_t dd = (size_t)d * d;
    // items per pass bounded by the Gram work area (<= ~1 GiB)
    const size_t per_item_gram = multiple_cv ? 2 * (size_t)d2 * d2 : 0;
    int chunk = batch;
    if (per_item_gram) chunk = (int)std::max<size_t>(1, std::min<size_t>(batch, ((size_t)1 << 27) / per_item_gram));
    // device buffers
    size_t off = 0;
    auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 1) & ~(size_t)1; return o; };
    const size_t o_draws = take((size_t)batch * d * n), o_lp = take((size_t)batch * n), o_Q = take(batch * dd),
                 o_r = take((size_t)batch * d), o_St = take(batch * dd), o_mt = take((size_t)batch * d),
                 o_pr = take((size_t)batch * n), o_mh = take((size_t)batch * d), o_sh = take(batch * dd),
                 o_fm = take((size_t)batch * d2), o_hm = take((size_t)batch * d2), o_res = take((size_t)batch * d2),
                 o_G1 = take((size_t)chunk * per_item_gram / 2 + 2), o_G2 = take((size_t)chunk * per_item_gram / 2 + 2),
                 o_feat = take((size_t)d2 + 2), o_status = take((size_t)batch / 2 + 2);
    const size_t a_m_cnt = (size_t)batch * (multiple_cv ? dd : (size_t)d), a_S_cnt = (size_t)batch * (multiple_cv ? (size_t)d2 * d2 : (size_t)d2);
    const size_t o_am = take(a_m_out ? a_m_cnt : 0), o_aS = take(a_S_out ? a_S_cnt : 0);
    EPG_CHECK(c, epg_reserve((void**)&c->util_buf, &c->util_bytes, sizeof(double) * off));
    double* base = c->util_buf;
    CvArgs a;
    a.n = n; a.d = d; a.d2 = d2; a.batch0 = 0;
    a.draws = base + o_draws; a.lp = base + o_lp; a.Qt = base + o_Q; a.rt = base + o_r;
    a.St = base + o_St; a.mt = base + o_mt; a.pr = base + o_pr; a.mhat = base + o_mh; a.shat = base + o_sh;
    a.fm = base + o_fm; a.hm = base + o_hm; a.res = base + o_res; a.G1 = base + o_G1; a.G2 = base + o_G2;
    a.feat = reinterpret_cast<const int2*>(base + o_feat);
    a.status = reinterpret_cast<int*>(base + o_status);
    a.multiple_cv = multiple_cv; a.regulate_a = regulate_a; a.max_a = max_a; a.m_treshold = m_treshold;
    a.aout = nullptr;
    // (items that take the plain-estimate fallback report zero coefficients: util.py:364-365)
    if (a_m_out) EPG_CHECK(c, cudaMemsetAsync(base + o_am, 0, sizeof(double) * a_m_cnt, c->stream));
    if (a_S_out) EPG_CHECK(c, cudaMemsetAsync(base + o_aS, 0, sizeof(double) * a_S_cnt, c->stream));
    std::vector<int2> feat(d2);
    {
        int e = 0;
        for (int x = 0; x < d; ++x) for (int y = x; y < d; ++y) feat[e++] = make_int2(x, y);
    }
    cudaStream_t st = c->stream;
    EPG_CHECK(c, cudaMemcpyAsync(base + o_draws, draws, sizeof(double) * (size_t)batch * d * n, cudaMemcpyHostToDevice, st));
    EPG_CHECK(c, cudaMemcpyAsync(base + o_lp, lp, sizeof(double) * (size_t)batch * n, cudaMemcpyHostToDevice, st));
    EPG_CHECK(c, cudaMemcpyAsync(base + o_Q, Q_tilde, sizeof(double) * batch * dd, cudaMemcpyHostToDevice, st));
    EPG_CHECK(c, cudaMemcpyAsync(base + o_r, r_tilde, sizeof(double) * (size_t)batch * d, cudaMemcpyHostToDevice, st));
    EPG_CHECK(c, cudaMemcpyAsync(base + o_feat, feat.data(), sizeof(int2) * d2, cudaMemcpyHostToDevice, st));
    const int lt = d <= 16 ? 32 : (d <= 32 ? 64 : (d <= 64 ? 128 : 256));
    const size_t sm_prep = sizeof(double) * ((size_t)d * (d + 1) / 2 + 3 * (size_t)d + 48);
    EPG_CHECK(c, cudaFuncSetAttribute(k_cv_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_prep));
    k_cv_prep<<<batch, lt, sm_prep, st>>>(a);
    k_cv_fallback<<<batch, 128, sizeof(double) * d, st>>>(a);
    c->launches += 2;
    const size_t sm_sums = sizeof(double) * ((size_t)d * 32 + 32 + 2 * (size_t)d);
    const size_t sm_gram = sizeof(double) * ((size_t)d * CV_T + CV_T + 2 * (size_t)d + 3 * CV_T * CV_TILE);
    for (int stage = 0; stage < 2; ++stage) {
        const int nf = stage == 0 ? d : d2;
        a.aout = stage == 0 ? (a_m_out ? base + o_am : nullptr) : (a_S_out ? base + o_aS : nullptr);
        k_cv_sums<<<dim3((nf + 127) / 128, batch), 128, sm_sums, st>>>(a, stage);
        c->launches++;
        if (multiple_cv) {
            const int tiles = (nf + CV_TILE - 1) / CV_TILE;
            for (int b0 = 0; b0 < batch; b0 += chunk) {
                const int nb = std::min(chunk, batch - b0);
                a.batch0 = b0;
                k_cv_gram<<<dim3(tiles, tiles, nb), 256, sm_gram, st>>>(a, stage);
                k_cv_solve<<<nb, nf >= 256 ? 1024 : 256, 0, st>>>(a, stage);
                c->launches += 2;
            }
            a.batch0 = 0;
        } else {
            k_cv_diag<<<dim3((nf + 127) / 128, batch), 128, sm_sums, st>>>(a, stage);
            c->launches++;
        }
        k_cv_store<<<dim3((nf + 127) / 128, batch), 128, 0, st>>>(a, stage);
        c->launches++;
        EPG_CHECK(c, cudaGetLastError());
    }
    EPG_CHECK(c, cudaMemcpyAsync(S_hat, base + o_sh, sizeof(double) * batch * dd, cudaMemcpyDeviceToHost, st));
    EPG_CHECK(c, cudaMemcpyAsync(m_hat, base + o_mh, sizeof(double) * (size_t)batch * d, cudaMemcpyDeviceToHost, st));
    if (a_m_out) EPG_CHECK(c, cudaMemcpyAsync(a_m_out, base + o_am, sizeof(double) * a_m_cnt, cudaMemcpyDeviceToHost, st));
    if (a_S_out) EPG_CHECK(c, cudaMemcpyAsync(a_S_out, base + o_aS, sizeof(double) * a_S_cnt, cudaMemcpyDeviceToHost, st));
    std::vector<int> stat(batch);
    EPG_CHECK(c, cudaMemcpyAsync(stat.data(), a.status, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
    EPG_CHECK(c, cudaStreamSynchronize(st));
    if (used_cv) for (int i = 0; i < batch; ++i) used_cv[i] = stat[i];
    return 0;
}
