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
#include <algorithm>
#include <stdlib.h>
#include <stdio.h>

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

// ---------------------------------------------------------------------------
// Second half, shared by both accumulation kernels.  In: the packed lower Gram of the augmented
// vector [1, x - shift] in `gram` (dimension da = d + 1: column 0 holds n and the sums, the
// remainder is the packed scatter matrix about `shift`).  Out: the site's (dQi, dri), tilted mean, flag.
// vec: 5 d + 40 doubles of shared scratch whose first d entries hold `shift`.
// ---------------------------------------------------------------------------
template <int MODE>
__device__ void moments_tail(const Grp& g, double* gram, double* vec, const unsigned short* tri, int n, int d, int k,
                             const double* __restrict__ Q, const double* __restrict__ r,
                             double* __restrict__ dQi, double* __restrict__ dri,
                             double* __restrict__ tmean, int* __restrict__ ok) {
    const int da = d + 1;
    double* shift = vec;            // d
    double* mt = vec + d;           // d
    double* sol = vec + 2 * d;      // d
    double* col = vec + 3 * d;      // d
    double* red = vec + 4 * d;      // 40
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
    // A^-1 mt is taken from the explicit inverse factor X = L^-1 (two triangular matrix-vector products, two
    // barriers) instead of dpotrs' forward / backward substitutions (2 d barrier-separated column steps)
    bool good = chol_packed(g, A, d, tri);
    if (good) {
        trtri_packed(g, A, col, d);                      // A <- X = L^-1 (packed lower, by columns)
        if (MODE == EPG_PREC_SAMPLE) {
            for (int i = g.tid; i < d; i += g.n) {       // y = X mt
                double a0 = 0.0, a1 = 0.0;
                int k = 0;
                for (; k + 2 <= i + 1; k += 2) {
                    a0 = fma(A[pk(i, k, d)], mt[k], a0);
                    a1 = fma(A[pk(i, k + 1, d)], mt[k + 1], a1);
                }
                if (k <= i) a0 = fma(A[pk(i, k, d)], mt[k], a0);
                col[i] = a0 + a1;
            }
            g.sync();
            for (int j = g.tid; j < d; j += g.n) {       // sol = X' y
                const double* cj = A + pk_col(j, d);
                double a0 = 0.0, a1 = 0.0;
                int i = j;
                for (; i + 2 <= d; i += 2) {
                    a0 = fma(cj[i], col[i], a0);
                    a1 = fma(cj[i + 1], col[i + 1], a1);
                }
                if (i < d) a0 = fma(cj[i], col[i], a0);
                sol[j] = a0 + a1;
            }
        }
        lauum_full(g, A, d, kf, outQ, tri);
    }
    __threadfence_block();
    g.sync();
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
    double* shift = vec;            // d (the rest of vec: scratch of moments_tail)
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

    moments_tail<MODE>(g, gram, vec, tri, n, d, k, Q, r, dQi, dri, tmean, ok);
}


// ===========================================================================
// Tensor-core accumulation (fp64 DMMA) with TMA-fed shared-memory stages.
//
//   * a producer warp streams the site's (n,d) F-order draws with 1-D bulk copies
//     (cp.async.bulk, one per parameter row and stage of MM_TCH draws) into a ring of
//     stages guarded by full / empty mbarriers -- HBM is read exactly once, no thread
//     ever waits on a global load;
//   * compute warps own 4 x 4 tiles of 8 x 8 blocks of the lower triangle of the Gram
//     of the augmented vector [1, x - shift] and accumulate them with
//     mma.sync.m8n8k4.f64 (DMMA): a stage row IS both the A (row-major 8 x 4) and the
//     B (col-major 4 x 8) fragment of its block, so one k-step of a tile costs 8 fragment
//     loads for up to 16 DMMAs and the accumulators never leave registers;
//   * `S` warps per tile take alternate k-steps (draw slices) and are summed in a fixed
//     order at the end (deterministic).
// The factorisation / inversion tail is shared with the SIMT kernel (moments_tail).
// ===========================================================================
namespace mm {
constexpr int TCH = 32;            // draws per stage
constexpr int PITCH = 36;          // doubles per stage row (== 4 mod 16: conflict-free fragment loads)
constexpr int TB = 4;              // blocks per tile edge
constexpr int MAX_TILES = 32;     // tiles of the lower triangle (d <= 200: 28), spread over `parts` CTAs per site

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t par) {
    const uint32_t a = s32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(a), "r"(par) : "memory");
        if (spin > (1u << 26)) __trap();          // never hang the GPU
    }
}
// global -> shared bulk copy (TMA engine), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void compute_sync(int nthreads) { asm volatile("bar.sync 1, %0;\n" ::"r"(nthreads) : "memory"); }

struct Plan {                      // kernel parameter (host-built)
    int W, S, G, NS, rows;         // compute warps = G * S, stages, stage rows (8 * blocks)
    int parts, Gtot, direct;       // CTAs per site (each owns G of the Gtot tiles and streams all the draws);
                                   // direct: accumulators go straight to the global Gram slot (no smem copy; S == 1)
    int off_gram, off_vec, off_bar, off_stage, total;
    unsigned char ti[MAX_TILES], tj[MAX_TILES];      // tile (block-row / block-column index, in tiles) of every group
};
}  // namespace mm

__global__ void k_gram_mma(const double* __restrict__ draws, int n, int d, int k0,
                           double* __restrict__ gbuf, size_t gstride, const mm::Plan P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];     // (the stage offset is 128-byte aligned: Plan)
    using namespace mm;
    const Grp g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int da = d + 1;
    double* gram = reinterpret_cast<double*>(smem_raw + P.off_gram);
    double* vec = reinterpret_cast<double*>(smem_raw + P.off_vec);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + P.off_bar);
    uint64_t* empty = full + P.NS;
    double* stage0 = reinterpret_cast<double*>(smem_raw + P.off_stage);
    const int stage_doubles = P.rows * PITCH;
    double* shiftaug = vec + 5 * d + 40;             // [rows]: 0, shift_0 .. shift_{d-1}, 0 ...
    const int site = blockIdx.x / P.parts, part = blockIdx.x - site * P.parts;
    const double* x = draws + (size_t)(k0 + site) * d * n;
    double* G = gbuf + (size_t)site * gstride;
    const int nchunks = (n + TCH - 1) / TCH;
    const int ncomp = P.W * 32;

    if (tid == 0) {
        for (int i = 0; i < P.NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, P.W); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // constant rows of every stage: row 0 = the 1 of the augmented vector, rows > d = padding
    for (int e = tid; e < P.NS * (P.rows - d) * PITCH; e += blockDim.x) {
        const int st = e / ((P.rows - d) * PITCH), q = e - st * (P.rows - d) * PITCH;
        const int rr = q / PITCH, cc = q - rr * PITCH;
        stage0[(size_t)st * stage_doubles + (rr == 0 ? 0 : d + rr) * PITCH + cc] = rr == 0 ? 1.0 : 0.0;
    }
    for (int e = tid; e < P.rows; e += blockDim.x) shiftaug[e] = 0.0;
    __syncthreads();

    if (warp == P.W) {
        // ===== producer =====
        for (int c = 0; c < nchunks; ++c) {
            const int st = c % P.NS;
            if (c >= P.NS) mbar_wait(empty + st, ((c / P.NS) - 1) & 1);
            const int tn = min(TCH, n - c * TCH);
            if (lane == 0) mbar_expect_tx(full + st, (uint32_t)(d * tn * 8));
            __syncwarp();
            double* dst = stage0 + (size_t)st * stage_doubles + PITCH;       // row 1
            for (int i = lane; i < d; i += 32)
                bulk_g2s(dst + i * PITCH, x + (size_t)i * n + (size_t)c * TCH, (uint32_t)(tn * 8), full + st);
        }
    } else if (warp < P.W) {
        // ===== compute: tile (ti, tj) of 4 x 4 blocks, draw slice `sl` =====
        const int grp = warp / P.S, sl = warp - grp * P.S;
        const int gidx = part * P.G + grp;          // this warp's tile; warps beyond the last tile only keep the ring moving
        const bool owner = gidx < P.Gtot;
        const int bi0 = owner ? P.ti[gidx] * TB : 0, bj0 = owner ? P.tj[gidx] * TB : 0;
        const int nblk = (da + 7) >> 3;             // blocks that hold data (the stage has rows for whole tiles)
        const bool diag = bi0 == bj0;
        double acc[TB][TB][2];
#pragma unroll
        for (int i = 0; i < TB; ++i)
#pragma unroll
            for (int j = 0; j < TB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const int frow = lane >> 2, fcol = lane & 3;
        // fragment of block b at k-step ks: stage[(8 b + frow) * PITCH + 4 ks + fcol] - shift[8 b + frow]
        const int offa = (8 * bi0 + frow) * PITCH + fcol, offb = (8 * bj0 + frow) * PITCH + fcol;
        double sha[TB], shb[TB];
        auto kstep = [&](const double* sg, int ks, bool valid) {
            double fa[TB], fb[TB];
#pragma unroll
            for (int i = 0; i < TB; ++i) {
                fa[i] = sg[offa + i * 8 * PITCH + 4 * ks] - sha[i];
                fb[i] = sg[offb + i * 8 * PITCH + 4 * ks] - shb[i];
                if (!valid) fa[i] = fb[i] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < TB; ++i)
#pragma unroll
                for (int j = 0; j < TB; ++j)
                    if ((!diag || j <= i) && bi0 + i < nblk) dmma(acc[i][j], fa[i], fb[j]);
        };
        for (int c = 0; c < nchunks; ++c) {
            const int st = c % P.NS;
            mbar_wait(full + st, (c / P.NS) & 1);
            const double* sg = stage0 + (size_t)st * stage_doubles;
            const int tn = min(TCH, n - c * TCH);
            if (c == 0) {
                // provisional mean (first stage) -> shift; every compute warp needs it before its first fragment
                for (int i = tid; i < d; i += ncomp) {
                    double sum = 0.0;
                    for (int t = 0; t < tn; ++t) sum += sg[(1 + i) * PITCH + t];
                    shiftaug[1 + i] = sum / tn;
                }
                compute_sync(ncomp);
#pragma unroll
                for (int i = 0; i < TB; ++i) {
                    sha[i] = shiftaug[8 * (bi0 + i) + frow];
                    shb[i] = shiftaug[8 * (bj0 + i) + frow];
                }
            }
            if (!owner) {
            } else if (tn == TCH) {
                for (int ks = sl; ks < TCH / 4; ks += P.S) kstep(sg, ks, true);
            } else {
                for (int ks = sl; ks < TCH / 4; ks += P.S) kstep(sg, ks, 4 * ks + fcol < tn);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
        }
        // fixed-order sum over the draw slices into the packed Gram (direct: one slice, straight to global memory)
        double* gdst = P.direct ? G : gram;
        for (int round = 0; round < P.S; ++round) {
            if (sl == round && owner) {
#pragma unroll
                for (int i = 0; i < TB; ++i)
#pragma unroll
                    for (int j = 0; j < TB; ++j) {
                        if (diag && j > i) continue;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int gi = 8 * (bi0 + i) + frow, gj = 8 * (bj0 + j) + 2 * fcol + e;
                            if (gi < da && gj <= gi) {
                                const int idx = pk(gi, gj, da);
                                gdst[idx] = round == 0 ? acc[i][j][e] : gdst[idx] + acc[i][j][e];
                            }
                        }
                    }
            }
            if (P.S > 1) compute_sync(ncomp);
        }
    }
    __syncthreads();
    // packed Gram (dimension d + 1) followed by the shift -> this site's slot of the scratch buffer
    const int ng = pk_size(da);
    if (!P.direct) for (int e = tid; e < ng; e += blockDim.x) G[e] = gram[e];
    if (part == 0) for (int e = tid; e < d; e += blockDim.x) G[ng + e] = shiftaug[1 + e];
}

// ---- tail: the shared-memory factorisation / inversion on the Gram left by k_gram_mma ----
template <int MODE>
__global__ void k_moments_tail_smem(const double* __restrict__ gbuf, size_t gstride, int n, int d, int k0,
                                    const double* __restrict__ Q, const double* __restrict__ r,
                                    double* __restrict__ dQi, double* __restrict__ dri,
                                    double* __restrict__ tmean, int* __restrict__ ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp g;
    const int da = d + 1, ng = pk_size(da);
    double* gram = reinterpret_cast<double*>(smem_raw);
    double* vec = gram + ng;                                         // 5 d + 40
    unsigned short* tri = d <= 128 ? reinterpret_cast<unsigned short*>(vec + 5 * d + 40) : nullptr;
    const double* G = gbuf + (size_t)blockIdx.x * gstride;
    for (int e = g.tid; e < ng; e += g.n) gram[e] = G[e];
    for (int e = g.tid; e < d; e += g.n) vec[e] = G[ng + e];
    if (tri) build_tri_table(g, tri, d); else g.sync();
    moments_tail<MODE>(g, gram, vec, tri, n, d, k0 + blockIdx.x, Q, r, dQi, dri, tmean, ok);
}

// host: tile plan of the DMMA kernel; returns false when the shape is left to the SIMT kernel
static bool mma_plan(int d, int n, mm::Plan& P) {
    if (d < 4 || d > 200 || (n & 1) || n < 8) return false;
    const int nblk = (d + 1 + 7) / 8;
    const int ntile = (nblk + mm::TB - 1) / mm::TB;
    int G = 0;
    for (int ti = 0; ti < ntile; ++ti)
        for (int tj = 0; tj <= ti; ++tj) {
            if (G >= mm::MAX_TILES) return false;
            P.ti[G] = (unsigned char)ti; P.tj[G] = (unsigned char)tj; ++G;
        }
    P.Gtot = G;
    // d <= 104 (<= 10 tiles): one CTA per site, Gram assembled in shared memory.  Larger d: the tiles are spread over
    // `parts` CTAs per site (<= 14 compute warps each; every CTA streams all the draws, the second read hits L2) and
    // the accumulators go straight to the site's global Gram slot -- the packed Gram (162 kB at d = 200) would not
    // fit next to the stages
    P.parts = G <= 10 ? 1 : (G + 13) / 14;
    P.direct = P.parts > 1;
    P.G = (G + P.parts - 1) / P.parts;
    P.S = P.direct ? 1 : (G == 1 ? 4 : (G <= 3 ? 2 : 1));
    P.W = P.G * P.S;
    P.rows = 8 * mm::TB * ntile;                 // whole tiles: the padding rows hold zeros
    size_t o = 0;
    P.off_gram = (int)o; if (!P.direct) o += sizeof(double) * (size_t)pk_size(d + 1);
    P.off_vec = (int)o;  o += sizeof(double) * (size_t)(5 * d + 40 + P.rows);
    o = (o + 15) & ~(size_t)15;
    P.off_bar = (int)o;  o += 16 * 2 * 8;
    o = (o + 127) & ~(size_t)127;
    P.off_stage = (int)o;
    const size_t stage_bytes = sizeof(double) * (size_t)P.rows * mm::PITCH;
    P.NS = (int)std::min<size_t>(4, ((o + 4 * stage_bytes <= 100 * 1024 ? 100 : 220) * 1024 - o) / stage_bytes);
    if (P.NS < 2) return false;
    P.NS = std::min(P.NS, 8);
    o += stage_bytes * P.NS;
    P.total = (int)o;
    return true;
}

}  // namespace

int epg_moments_threads(int d) { return d <= 32 ? 128 : (d <= 96 ? 256 : 512); }

cudaError_t epg_launch_moments(epg_ctx* c, int k0, int k1, int n, int mode) {
    const int d = c->d;
    static const bool no_mma = getenv("EPGPU_MOMENTS_SIMT") != nullptr;      // (A/B timing of the two kernels)
    mm::Plan P;
    if (!no_mma && mma_plan(d, n, P) && (reinterpret_cast<uintptr_t>(c->draws) & 15) == 0) {
        const int K = k1 - k0;
        const size_t gstride = (size_t)pk_size(d + 1) + d;
        cudaError_t e = epg_reserve((void**)&c->mom_buf, &c->mom_bytes, sizeof(double) * gstride * K);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_gram_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, P.total);
        if (e != cudaSuccess) return e;
        k_gram_mma<<<K * P.parts, 32 * (P.W + 1), P.total, c->stream>>>(c->draws, n, d, k0, c->mom_buf, gstride, P);
        c->launches++;
        const bool olse = mode == EPG_PREC_OLSE;
#define EPG_TAIL_ARGS c->mom_buf, gstride, n, d, k0, c->arr[EPG_Q], c->arr[EPG_R], c->arr[EPG_DQI], c->arr[EPG_DRI], \
                      c->arr[EPG_TMEAN], c->site_ok
        {
            const size_t sm = sizeof(double) * ((size_t)pk_size(d + 1) + 5 * d + 40) +
                              ((sizeof(unsigned short) * (size_t)pk_size(d) + 15) & ~(size_t)15);
            auto kern = olse ? k_moments_tail_smem<EPG_PREC_OLSE> : k_moments_tail_smem<EPG_PREC_SAMPLE>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return e;
            kern<<<K, epg_moments_threads(d), sm, c->stream>>>(EPG_TAIL_ARGS);
        }
#undef EPG_TAIL_ARGS
        c->launches++;
        return cudaGetLastError();
    }
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
