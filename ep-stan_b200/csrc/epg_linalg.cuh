// Small dense fp64 linear algebra on one symmetric d x d matrix held in shared
// memory as a packed lower triangle (column by column), worked on by one CTA.
// These replace the reference's per-site LAPACK calls:
//   dpotrf (linalg.cho_factor: method.py:294,872,1080; util.py:112,328)
//   dpotrs (linalg.cho_solve:  method.py:295; util.py:117)
//   dpotri (util.py:45,118) + copy_triu_to_tril (cython_util.pyx:86-106)
//   dsyevr smallest eigenvalue (linalg.eigvalsh(eigvals=(0,0)): method.py:1124,1199)
#pragma once
#include "epg_common.cuh"
#include <math.h>

// advance a walk over packed-lower storage restricted to columns >= c0:
// (i,k) -> the element `step` places later (column k holds rows k..d-1)
__device__ __forceinline__ void pk_advance(int& i, int& k, int step, int d) {
    i += step;
    while (i >= d && k < d) { i = i - d + k + 1; ++k; }
}

// Optional lookup table of the (row, col) of every packed-lower element, built once per
// kernel in shared memory: tri[e] = (row << 8) | col.  Replaces the index walk in the
// O(d^2)-per-column loops.
__device__ inline void build_tri_table(const Grp& g, unsigned short* tri, int d) {
    for (int j = g.tid; j < d; j += g.n) {
        const int cj = pk_col(j, d);
        for (int i = j; i < d; ++i) tri[cj + i] = (unsigned short)((i << 8) | j);
    }
    g.sync();
}

// In-place Cholesky  A = L L'.  Returns false (uniformly) when a pivot is not
// strictly positive or not finite (LAPACK dpotrf info>0).  Ends with a barrier.
__device__ inline bool chol_packed(const Grp& g, double* A, int d, const unsigned short* tri = nullptr) {
    for (int j = 0; j < d; ++j) {
        g.sync();
        const double ajj = A[pk(j, j, d)];
        if (!(ajj > 0.0) || !isfinite(ajj)) return false;
        const double ljj = sqrt(ajj), inv = 1.0 / ljj;
        const int cj = pk_col(j, d);
        g.sync();
        for (int i = j + g.tid; i < d; i += g.n) A[cj + i] = (i == j) ? ljj : A[cj + i] * inv;
        g.sync();
        // trailing update A(i,k) -= L(i,j) L(k,j), j<k<=i : the packed tail
        const int base = pk(j + 1, j + 1, d);           // valid also when j+1==d (== size)
        const int m = pk_size(d) - base;
        if (tri) {
            for (int e = g.tid; e < m; e += g.n) {
                const unsigned ik = tri[base + e];
                A[base + e] -= A[cj + (ik >> 8)] * A[cj + (ik & 255u)];
            }
        } else {
            int i = j + 1, k = j + 1;
            pk_advance(i, k, g.tid, d);
            for (int e = g.tid; e < m; e += g.n) {
                A[base + e] -= A[cj + i] * A[cj + k];
                pk_advance(i, k, g.n, d);
            }
        }
    }
    g.sync();
    return true;
}

// b <- L^-1 b  (forward substitution, column oriented). b in shared memory.
__device__ inline void fwd_solve_packed(const Grp& g, const double* L, double* b, int d) {
    for (int j = 0; j < d; ++j) {
        g.sync();
        const int cj = pk_col(j, d);
        const double yj = b[j] / L[cj + j];
        for (int i = j + 1 + g.tid; i < d; i += g.n) b[i] -= L[cj + i] * yj;
    }
    g.sync();
    for (int j = g.tid; j < d; j += g.n) b[j] /= L[pk(j, j, d)];
    g.sync();
}

// b <- L^-T b  (backward substitution, row oriented on L).
__device__ inline void bwd_solve_packed(const Grp& g, const double* L, double* b, int d) {
    for (int j = d - 1; j >= 0; --j) {
        g.sync();
        const double xj = b[j] / L[pk(j, j, d)];
        for (int k = g.tid; k < j; k += g.n) b[k] -= L[pk(j, k, d)] * xj;   // (row j of L)
    }
    g.sync();
    for (int j = g.tid; j < d; j += g.n) b[j] /= L[pk(j, j, d)];
    g.sync();
}

// In-place inverse of the lower-triangular factor: L <- L^-1 (LAPACK dtrti2,
// lower).  `col` is shared scratch of d doubles.
__device__ inline void trtri_packed(const Grp& g, double* L, double* col, int d) {
    for (int j = d - 1; j >= 0; --j) {
        const int cj = pk_col(j, d);
        g.sync();
        for (int i = j + g.tid; i < d; i += g.n) col[i] = L[cj + i];
        g.sync();
        const double xjj = 1.0 / col[j];
        for (int i = j + g.tid; i < d; i += g.n) {
            double acc;
            if (i == j) acc = xjj;
            else {
                // new(i) = -xjj * sum_{k=j+1..i} X(i,k) * Lold(k,j)
                acc = 0.0;
                int idx = pk(i, j + 1, d);                 // (i, k) -> (i, k+1): + d - k - 1
                for (int k = j + 1; k <= i; ++k) { acc += L[idx] * col[k]; idx += d - k - 1; }
                acc *= -xjj;
            }
            L[cj + i] = acc;
        }
    }
    g.sync();
}

// out = scale * X' X for lower-triangular X (packed): the full symmetric d x d
// result, written column-major to `out` (global or shared).  Both triangles are
// written: this is dpotri followed by copy_triu_to_tril.
__device__ inline void lauum_full(const Grp& g, const double* X, int d, double scale, double* out,
                                  const unsigned short* tri = nullptr) {
    const int m = pk_size(d);
    int i = 0, j = 0;
    pk_advance(i, j, g.tid, d);
    for (int e = g.tid; e < m; e += g.n) {
        if (tri) { const unsigned ij = tri[e]; i = (int)(ij >> 8); j = (int)(ij & 255u); }
        // (i,j), i >= j : sum_{k>=i} X(k,i) X(k,j)
        const double* ci = X + pk_col(i, d);
        const double* cj = X + pk_col(j, d);
        double acc = 0.0;
        for (int k = i; k < d; ++k) acc += ci[k] * cj[k];
        acc *= scale;
        out[i + j * d] = acc;
        out[j + i * d] = acc;
        if (!tri) pk_advance(i, j, g.n, d);
    }
}

// ---------------------------------------------------------------------------
// Smallest eigenvalue of a symmetric matrix: Householder tridiagonalisation of
// the packed lower triangle (LAPACK dsptrd, lower) followed by Sturm-count
// multisection (dstebz-style).  A is destroyed.  dg/od/v/w: shared scratch of d
// doubles each; red: >= 33 doubles.  Result returned to all threads.
// ---------------------------------------------------------------------------
__device__ inline double min_eig_packed(const Grp& g, double* A, int d, double* dg, double* od,
                                        double* v, double* w, double* red) {
    for (int k = 0; k + 2 < d; ++k) {
        const int ck = pk_col(k, d);
        g.sync();
        // x = A(k+1:d, k)
        double part = 0.0;
        for (int i = k + 2 + g.tid; i < d; i += g.n) part += A[ck + i] * A[ck + i];
        const double xnorm2 = block_sum(g, part, red);     // ||x(1:)||^2 (without head)
        const double alpha = A[ck + k + 1];
        double tau = 0.0, beta = alpha;
        if (xnorm2 > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            tau = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            for (int i = k + 1 + g.tid; i < d; i += g.n) v[i] = (i == k + 1) ? 1.0 : A[ck + i] * sc;
        } else {
            for (int i = k + 1 + g.tid; i < d; i += g.n) v[i] = (i == k + 1) ? 1.0 : 0.0;
        }
        g.sync();
        if (g.tid == 0) { od[k] = beta; dg[k] = A[ck + k]; }
        if (tau != 0.0) {
            // p = tau * A22 v   (A22 = trailing (d-k-1) block, symmetric, packed lower)
            for (int i = k + 1 + g.tid; i < d; i += g.n) {
                double acc = 0.0;
                for (int c = k + 1; c < d; ++c) {
                    const double a = (i >= c) ? A[pk(i, c, d)] : A[pk(c, i, d)];
                    acc += a * v[c];
                }
                w[i] = tau * acc;
            }
            g.sync();
            part = 0.0;
            for (int i = k + 1 + g.tid; i < d; i += g.n) part += w[i] * v[i];
            const double pv = block_sum(g, part, red);
            const double hf = -0.5 * tau * pv;
            for (int i = k + 1 + g.tid; i < d; i += g.n) w[i] += hf * v[i];
            g.sync();
            // A22 -= v w' + w v'
            const int base = pk(k + 1, k + 1, d);
            const int m = pk_size(d) - base;
            int i = k + 1, c = k + 1;
            pk_advance(i, c, g.tid, d);
            for (int e = g.tid; e < m; e += g.n) {
                A[base + e] -= v[i] * w[c] + w[i] * v[c];
                pk_advance(i, c, g.n, d);
            }
        }
    }
    g.sync();
    if (g.tid == 0) {
        if (d >= 2) {
            dg[d - 2] = A[pk(d - 2, d - 2, d)];
            od[d - 2] = A[pk(d - 1, d - 2, d)];
        }
        dg[d - 1] = A[pk(d - 1, d - 1, d)];
    }
    g.sync();
    // Gershgorin bounds (block-wide min/max through shared memory)
    double lo, hi;
    {
        double l2 = 1e300, h2 = -1e300;
        for (int i = g.tid; i < d; i += g.n) {
            const double rad = (i > 0 ? fabs(od[i - 1]) : 0.0) + (i + 1 < d ? fabs(od[i]) : 0.0);
            l2 = fmin(l2, dg[i] - rad);
            h2 = fmax(h2, dg[i] + rad);
        }
        l2 = -warp_max(-l2);
        h2 = warp_max(h2);
        const int wv = g.tid >> 5, ln = g.tid & 31, nw = (g.n + 31) >> 5;
        __shared__ double s_lo[32], s_hi[32];
        if (ln == 0) { s_lo[wv] = l2; s_hi[wv] = h2; }
        g.sync();
        lo = s_lo[0]; hi = s_hi[0];
        for (int q = 1; q < nw; ++q) { lo = fmin(lo, s_lo[q]); hi = fmax(hi, s_hi[q]); }
        g.sync();
    }
    const double scale = fmax(fabs(lo), fabs(hi));
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, scale * scale);
    lo -= 2.0 * 2.2e-16 * scale * d + 2.0 * pivmin;
    hi += 2.0 * 2.2e-16 * scale * d + 2.0 * pivmin;
    // multisection: thread t evaluates the Sturm count at lo + (t+1)/(n+1)*(hi-lo);
    // lambda_min lies in the first sub-interval whose right end has count >= 1
    __shared__ int s_first;
    for (int round = 0; round < 64; ++round) {
        if (!(hi - lo > 4.0 * 2.2e-16 * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin)) break;
        if (g.tid == 0) s_first = g.n;
        g.sync();
        const double x = lo + (hi - lo) * (double)(g.tid + 1) / (double)(g.n + 1);
        int cnt = 0;
        double q = dg[0] - x;
        if (fabs(q) < pivmin) q = -pivmin;
        if (q < 0.0) ++cnt;
        for (int i = 1; i < d; ++i) {
            q = dg[i] - x - od[i - 1] * od[i - 1] / q;
            if (fabs(q) < pivmin) q = -pivmin;
            if (q < 0.0) ++cnt;
        }
        if (cnt >= 1) atomicMin(&s_first, g.tid);
        g.sync();
        const int f = s_first;
        const double nlo = (f == 0) ? lo : lo + (hi - lo) * (double)f / (double)(g.n + 1);
        const double nhi = (f == g.n) ? hi : lo + (hi - lo) * (double)(f + 1) / (double)(g.n + 1);
        g.sync();
        lo = nlo; hi = nhi;
    }
    return 0.5 * (lo + hi);
}
