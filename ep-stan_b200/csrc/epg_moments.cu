// Tilted-moment estimation -> site natural-parameter deltas, one CTA per site.
//
// Replaces, for all K sites in one launch, the second half of Worker.tilted
// (reference epstan/method.py:408-468):
//   'sample': mean, centre, dgeqrf -> R, invert_normal_params(cho_form) on R'R,
//             x (n-d-2)                                   method.py:413-437
//   'olse'  : scatter/n, util.olse(S, n, P=Q), dri = Qhat mt   method.py:440-451, util.py:128-194
//   then  dQi -= Q, dri -= r                              method.py:457-458
//   failure -> zero fill, flag 0                          method.py:460-465
//
// Data flow: the (n,d) F-order draws of the site are streamed from HBM exactly
// once in chunks of T draws into shared memory (transposed to draw-major and
// shifted by a provisional mean); the augmented vector xt=[1, x-a] is
// accumulated into a packed lower-triangular Gram matrix with 4x4 register
// tiles; column 0 of that Gram holds n and the sums, the rest IS the packed
// scatter matrix, which is then factorised / inverted in place in shared memory.
#include "epg_internal.h"
#include "epg_linalg.cuh"

namespace {

struct MomSmem {
    int da, dpad, T, nb_rows, NB, nslices;
    size_t off_gram, off_chunk, off_vec, off_part, off_tri, total;
};

__host__ __device__ inline MomSmem mom_layout(int d, int nthr) {
    MomSmem L;
    L.da = d + 1;
    L.dpad = (L.da + 3) & ~3;
    if ((L.dpad & 15) == 0) L.dpad += 2;          // fewer bank conflicts on the transposing store
    L.nb_rows = (L.da + 3) / 4;
    L.NB = L.nb_rows * (L.nb_rows + 1) / 2;
    L.nslices = L.NB <= nthr ? nthr / L.NB : 1;
    if (L.nslices > 16) L.nslices = 16;
    L.T = d > 96 ? 16 : 32;
    size_t o = 0;
    L.off_gram = o;  o += sizeof(double) * (size_t)pk_size(L.da);
    L.off_vec = o;   o += sizeof(double) * (size_t)(5 * d + 40);
    o = (o + 15) & ~(size_t)15;
    L.off_chunk = o; o += sizeof(double) * (size_t)L.T * L.dpad;
    L.off_part = o;  // per-slice partial tiles (only when nslices > 1)
    if (L.nslices > 1) o += sizeof(double) * (size_t)L.nslices * L.NB * 16;
    L.off_tri = o;   // (row,col) lookup of the packed scatter matrix
    if (d <= 128) o += (sizeof(unsigned short) * (size_t)pk_size(d) + 15) & ~(size_t)15;   // (large d: index walk)
    L.total = o;
    return L;
}

// block index b -> (bi, bj), bi >= bj, enumerated row by row
__device__ __forceinline__ void blk_decode(int b, int& bi, int& bj) {
    bi = (int)((sqrtf(8.0f * (float)b + 1.0f) - 1.0f) * 0.5f);
    while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
    while (bi * (bi + 1) / 2 > b) --bi;
    bj = b - bi * (bi + 1) / 2;
}

template <int MODE>
__global__ void k_moments(const double* __restrict__ draws, int n, int d, int k0,
                          const double* __restrict__ Q, const double* __restrict__ r,
                          double* __restrict__ dQi, double* __restrict__ dri,
                          double* __restrict__ tmean, int* __restrict__ ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    const MomSmem L = mom_layout(d, g.n);
    double* gram = reinterpret_cast<double*>(smem_raw + L.off_gram);   // packed lower, dim da
    double* vec = reinterpret_cast<double*>(smem_raw + L.off_vec);
    double* xs = reinterpret_cast<double*>(smem_raw + L.off_chunk);    // [T][dpad]
    double* part = reinterpret_cast<double*>(smem_raw + L.off_part);
    unsigned short* tri = d <= 128 ? reinterpret_cast<unsigned short*>(smem_raw + L.off_tri) : nullptr;
    double* shift = vec;            // d
    double* mt = vec + d;           // d
    double* sol = vec + 2 * d;      // d
    double* col = vec + 3 * d;      // d
    double* red = vec + 4 * d;      // 40
    const int da = L.da, dpad = L.dpad, T = L.T;
    const int k = k0 + blockIdx.x;
    const double* x = draws + (size_t)k * d * n;

    for (int e = g.tid; e < pk_size(da); e += g.n) gram[e] = 0.0;
    if (tri) build_tri_table(g, tri, d);
    // provisional mean of the first chunk
    {
        const int t0n = n < T ? n : T;
        for (int i = g.tid; i < d; i += g.n) {
            double s = 0.0;
            for (int t = 0; t < t0n; ++t) s += x[(size_t)i * n + t];
            shift[i] = s / t0n;
        }
    }
    g.sync();

    // ---- streaming accumulation -------------------------------------------
    const bool single = L.NB <= g.n;          // one tile per thread, kept in registers
    const int my_slice = single ? g.tid / L.NB : 0;
    const bool active = single ? (my_slice < L.nslices) : true;
    const int b_first = single ? g.tid % L.NB : g.tid;
    const int b_step = single ? L.NB : g.n;   // (single: loop body runs once)
    double acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0;

    // software pipeline: the draws of chunk c+1 are fetched into registers while chunk c is reduced
    constexpr int PF = 8;                       // elements per thread per chunk that are prefetched
    double pre[PF];
    const bool pf_ok = (T * d + g.n - 1) / g.n <= PF;
    auto fetch = [&](int c0) {
        const int tn = (n - c0) < T ? (n - c0) : T;
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int idx = g.tid + u * g.n;
            const int t = idx % T, i = idx / T;
            pre[u] = (idx < T * d && t < tn) ? x[(size_t)i * n + c0 + t] : 0.0;
        }
    };
    if (pf_ok) fetch(0);
    for (int c0 = 0; c0 < n; c0 += T) {
        const int tn = (n - c0) < T ? (n - c0) : T;
        // load + shift + transpose: xs[t][0]=1, xs[t][1+i] = x_i - a_i, zero padding
        if (pf_ok) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int idx = g.tid + u * g.n;
                if (idx < T * d) {
                    const int t = idx % T, i = idx / T;
                    xs[t * dpad + 1 + i] = (t < tn) ? pre[u] - shift[i] : 0.0;
                }
            }
            if (c0 + T < n) fetch(c0 + T);
        } else {
            for (int idx = g.tid; idx < T * d; idx += g.n) {
                const int t = idx % T, i = idx / T;
                xs[t * dpad + 1 + i] = (t < tn) ? x[(size_t)i * n + c0 + t] - shift[i] : 0.0;
            }
        }
        for (int idx = g.tid; idx < T * (dpad - d); idx += g.n) {
            const int t = idx % T, i = idx / T;       // i = 0 -> the constant, else padding
            xs[t * dpad + (i == 0 ? 0 : d + i)] = (i == 0 && t < tn) ? 1.0 : 0.0;
        }
        g.sync();
        if (active) {
            for (int b = b_first; b < L.NB; b += b_step) {
                int bi, bj;
                blk_decode(b, bi, bj);
                if (!single) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int i = 4 * bi + (q >> 2), j = 4 * bj + (q & 3);
                        acc[q] = (i < da && j <= i) ? gram[pk(i, j, da)] : 0.0;
                    }
                }
                const double* pi = xs + 4 * bi;
                const double* pj = xs + 4 * bj;
                for (int t = my_slice; t < tn; t += (single ? L.nslices : 1)) {
                    const double2 a01 = *reinterpret_cast<const double2*>(pi + t * dpad);
                    const double2 a23 = *reinterpret_cast<const double2*>(pi + t * dpad + 2);
                    const double2 b01 = *reinterpret_cast<const double2*>(pj + t * dpad);
                    const double2 b23 = *reinterpret_cast<const double2*>(pj + t * dpad + 2);
                    const double av[4] = {a01.x, a01.y, a23.x, a23.y};
                    const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[p * 4 + q] = fma(av[p], bv[q], acc[p * 4 + q]);
                }
                if (!single) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int i = 4 * bi + (q >> 2), j = 4 * bj + (q & 3);
                        if (i < da && j <= i) gram[pk(i, j, da)] = acc[q];
                    }
                }
                if (single) break;
            }
        }
        g.sync();
    }
    if (single) {
        // fixed-order reduction over slices -> deterministic
        if (L.nslices > 1) {
            if (active) {
#pragma unroll
                for (int q = 0; q < 16; ++q) part[((size_t)my_slice * L.NB + b_first) * 16 + q] = acc[q];
            }
            g.sync();
            if (g.tid < L.NB) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    double s = 0.0;
                    for (int sl = 0; sl < L.nslices; ++sl) s += part[((size_t)sl * L.NB + g.tid) * 16 + q];
                    acc[q] = s;
                }
            }
        }
        if (g.tid < L.NB) {
            int bi, bj;
            blk_decode(g.tid, bi, bj);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int i = 4 * bi + (q >> 2), j = 4 * bj + (q & 3);
                if (i < da && j <= i) gram[pk(i, j, da)] = acc[q];
            }
        }
        g.sync();
    }

    // ---- mean and centred scatter -----------------------------------------
    // gram column 0 = [n, s_0 .. s_{d-1}],  gram + da = packed lower of sum (x-a)(x-a)'
    double* A = gram + da;
    const double inv_n = 1.0 / (double)n;
    for (int i = g.tid; i < d; i += g.n) {
        const double m = shift[i] + gram[1 + i] * inv_n;
        mt[i] = m;
        sol[i] = m;
        tmean[(size_t)k * d + i] = m;
    }
    g.sync();
    {
        const double post = (MODE == EPG_PREC_OLSE) ? inv_n : 1.0;
        if (tri) {
            for (int e = g.tid; e < pk_size(d); e += g.n) {
                const unsigned ij = tri[e];
                A[e] = (A[e] - gram[1 + (ij >> 8)] * gram[1 + (ij & 255u)] * inv_n) * post;
            }
        } else {
            int i = 0, j = 0;
            pk_advance(i, j, g.tid, d);
            for (int e = g.tid; e < pk_size(d); e += g.n) {
                A[e] = (A[e] - gram[1 + i] * gram[1 + j] * inv_n) * post;
                pk_advance(i, j, g.n, d);
            }
        }
    }
    g.sync();

    double* outQ = dQi + (size_t)k * d * d;
    double* outr = dri + (size_t)k * d;
    // Small matrices: the column-sequential factorisation / inversion is done by ONE warp with
    // warp-level barriers (a CTA-wide barrier per column step costs more than the step);
    // the other warps wait.  Large matrices use the whole CTA.
    const double kf = (MODE == EPG_PREC_SAMPLE) ? (double)(n - d - 2) : 1.0;
    // (measured at K=1024, d=50: the one-warp variant is ~30 % slower than the CTA-wide one -> disabled)
    const bool one_warp = false;
    bool good;
    if (one_warp) {
        __shared__ int s_good;
        if (g.tid < EPG_WARP) {
            const Grp w(g.tid, EPG_WARP);
            bool ok_w = chol_packed(w, A, d, tri);
            if (ok_w) {
                if (MODE == EPG_PREC_SAMPLE) {
                    fwd_solve_packed(w, A, sol, d);
                    bwd_solve_packed(w, A, sol, d);
                }
                trtri_packed(w, A, col, d);
                lauum_full(w, A, d, kf, outQ, tri);
            }
            if (g.tid == 0) s_good = ok_w ? 1 : 0;
        }
        __threadfence_block();
        g.sync();
        good = s_good != 0;
    } else {
        good = chol_packed(g, A, d, tri);
        if (good) {
            if (MODE == EPG_PREC_SAMPLE) {
                fwd_solve_packed(g, A, sol, d);
                bwd_solve_packed(g, A, sol, d);
            }
            trtri_packed(g, A, col, d);
            lauum_full(g, A, d, kf, outQ, tri);
        }
        __threadfence_block();
        g.sync();
    }
    if (good) {
        if (MODE == EPG_PREC_SAMPLE) {
            double bad = 0.0;
            for (int e = g.tid; e < d * d; e += g.n) {
                const double v = outQ[e] - Q[e];
                outQ[e] = v;
                if (!isfinite(v)) bad = 1.0;
            }
            for (int i = g.tid; i < d; i += g.n) {
                const double v = kf * sol[i] - r[i];
                outr[i] = v;
                if (!isfinite(v)) bad = 1.0;
            }
            good = block_sum(g, bad, red) == 0.0;
        } else {
            // util.olse with prior P = Q (util.py:174-193)
            double ptr = 0.0, pf2 = 0.0, pf2p = 0.0, psp = 0.0;
            for (int e = g.tid; e < d * d; e += g.n) {
                const double s = outQ[e], p = Q[e];
                pf2 += s * s;
                pf2p += p * p;
                psp += s * p;
                if (e % (d + 1) == 0) ptr += s;
            }
            const double tr = block_sum(g, ptr, red);
            const double f2 = block_sum(g, pf2, red);
            const double f2p = block_sum(g, pf2p, red);
            const double trSP = block_sum(g, psp, red);
            const double dn = (double)d / (double)n;
            const double alpha = 1.0 - ((double)d + tr * tr * f2p / (f2 * f2p - trSP * trSP)) / (double)n;
            const double beta = (trSP / f2p) * (1.0 - dn - alpha);
            for (int e = g.tid; e < d * d; e += g.n) outQ[e] = alpha * outQ[e] + beta * Q[e];
            __threadfence_block();
            g.sync();
            double bad = 0.0;
            for (int i = g.tid; i < d; i += g.n) {
                double s = 0.0;
                for (int j = 0; j < d; ++j) s += outQ[i + (size_t)j * d] * mt[j];
                s -= r[i];
                outr[i] = s;
                if (!isfinite(s)) bad = 1.0;
            }
            g.sync();
            for (int e = g.tid; e < d * d; e += g.n) {
                const double v = outQ[e] - Q[e];
                outQ[e] = v;
                if (!isfinite(v)) bad = 1.0;
            }
            good = block_sum(g, bad, red) == 0.0;
        }
    }
    if (!good) {
        g.sync();
        for (int e = g.tid; e < d * d; e += g.n) outQ[e] = 0.0;
        for (int i = g.tid; i < d; i += g.n) outr[i] = 0.0;
    }
    if (g.tid == 0) ok[k] = good ? 1 : 0;
}

}  // namespace

int epg_moments_threads(int d) { return d <= 32 ? 128 : (d <= 96 ? 256 : 512); }

cudaError_t epg_launch_moments(epg_ctx* c, int k0, int k1, int n, int mode) {
    const int d = c->d;
    const int nthr = epg_moments_threads(d);
    const MomSmem L = mom_layout(d, nthr);
    auto kern = (mode == EPG_PREC_OLSE) ? k_moments<EPG_PREC_OLSE> : k_moments<EPG_PREC_SAMPLE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
    if (e != cudaSuccess) return e;
    kern<<<k1 - k0, nthr, L.total, c->stream>>>(c->draws, n, d, k0, c->arr[EPG_Q], c->arr[EPG_R],
                                                c->arr[EPG_DQI], c->arr[EPG_DRI], c->arr[EPG_TMEAN],
                                                c->site_ok);
    c->launches++;
    return cudaGetLastError();
}
