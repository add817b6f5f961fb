// "Wide" tensor-core (tcgen05 + TMEM + TMA) likelihood pass: single-group sites with up to 256 padded
// input columns (D+1 <= 256) and up to 32 chains -- BASELINE config 5 (n_k = 200 000, D = 199, 32 chains),
// where a site's design matrix (83 MB in bf16) is streamed from HBM once per gradient evaluation of all
// chains: the pass is HBM-bound, the job of the kernel is to keep enough TMA bytes in flight per SM.
//
//   design matrix: bf16 [N][pitch], pitch = (D+1 rounded up to 16) columns, centred per site, column D = 1;
//   a 128-row tile is nkc = ceil((D+1)/64) sub-tiles of 64 columns, each a TMA box [128 rows x 128 B]
//   (128-byte swizzle; the columns beyond `pitch` of the last box are out of bounds = zero-filled without
//   HBM traffic) in its own stage of a ring of up to 10 x 16 kB stages (9 fit on config 5).
//     GEMM1  F[128 x 64] = sum_j X_sub_j[128 x 64] . [B_hi | B_lo]_j'[64 x 64]   (K-chunked over the sub-tiles,
//              accumulated in one of 4 TMEM buffers; M=128, N=64 = 32 chains hi | 32 chains lo, K=16 per MMA)
//     epilogue (8 warps = two groups on alternate tiles, one row per thread): tcgen05.ld F_hi, F_lo ->
//              e = y - sigmoid(f), lp += y f - softplus(f); E (bf16, 32 chains) -> one of 4 shared E buffers
//     GEMM2  G_j[64 x 32] += X_sub_j'[64 x 128] . E[128 x 32] for every sub-tile j (the SAME smem sub-tile read as an
//              MN-major A operand; G_j has its own 64 TMEM columns); 4 instructions of M=128, N=64, K=16 per
//              sub-tile with the stacked-slices arrangement of epg_lik_tc.cuh (the two diagonal blocks of the
//              128 x 64 accumulator are the wanted products)
//   TMEM: 4 F buffers x 64 columns + 4 G chunks x 64 columns = 512 columns (the whole tensor memory of the SM).
//   Buffer reuse without extra barriers: E buffer / F buffer index = global tile counter & 3.  F(t) is free when
//   the epilogue of tile t-4 has arrived on eready; E(t) is free because fready(t) implies that every sub-tile of
//   tile t has landed, whose ring slots were released by GEMM2 of the sub-tiles nst before -- with nst <= 4 nkc
//   these include all of tile t-4 (GEMM2s complete in issue order).
//
// Descriptor encodings as in epg_lik_tc.cuh (its helpers are reused).
#pragma once
#include "epg_lik_tc.cuh"

namespace tcw {

using tc::TILE_M;
using tc::KW;
using tc::TILE_BYTES;
constexpr int NCH = 32;              // chains (N of GEMM2, half the N of GEMM1)
constexpr int NB1 = 2 * NCH;         // GEMM1 N: columns [0,32) hi parts, [32,64) lo parts of the coefficients
constexpr int NKC = 4;               // max sub-tiles (64-column chunks) per tile
constexpr int KWT = NKC * KW;        // 256 padded coefficient columns
constexpr int NST = 10;              // max ring stages (one sub-tile each)
constexpr int NF = 4;                // F accumulator buffers in TMEM == E buffers in smem
constexpr int B_BYTES = NB1 * KWT * 2;               // 32768
constexpr int E_BYTES = NCH * TILE_M * 2;            // 8192
constexpr int TMEM_COLS = 512;
constexpr int G_COL0 = NF * NB1;                     // 256: G chunk j at columns [256 + 64 j, 256 + 64 j + 64)
constexpr int NBAR = 2 * NST + 2 * NF + 1;

struct Smem {            // offsets relative to a 1024-byte aligned base
    static constexpr int BM = 0;                     // coefficient operand [64 rows (hi | lo)] x [256 columns], K-major interleaved
    static constexpr int BAR = BM + B_BYTES;
    static constexpr int TMEM_PTR = BAR + NBAR * 8;
    static constexpr int E = (TMEM_PTR + 16 + 1023) & ~1023;          // NF E buffers
    static constexpr int X = E + NF * E_BYTES;                        // nst X sub-tiles (1024-aligned)
    __host__ __device__ static constexpr int total(int nst) { return X + nst * TILE_BYTES; }
};
static_assert(Smem::X % 1024 == 0, "swizzled tiles need 1024-byte alignment");

// byte offset of coefficient element (row n of [hi | lo], column k) in the K-major INTERLEAVE operand of GEMM1:
// 8 x 16 B core matrices, the eight 8-row groups 128 B apart (SBO), K chunks of 8 columns 1024 B apart (LBO)
__device__ __forceinline__ int b_off(int n, int k) {
    return (k >> 3) * 1024 + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
}
// byte offset of E element (chain n, tile row rho): instruction i = rho >> 5 covers rows [32i, 32i+32); its B
// operand has N = 64 = [32 chains of rows 32i..32i+15 | 32 chains of rows 32i+16..32i+31], K = 16 rows in two
// chunks of 8 (LBO = 1024 B), 8-row N groups 128 B apart (SBO); 2048 B per instruction
__device__ __forceinline__ int e_off(int n, int rho) {
    return (rho >> 5) * 2048 + ((rho >> 3) & 1) * 1024 + ((((rho >> 4) & 1) << 2) + (n >> 3)) * 128 + (n & 7) * 16 +
           (rho & 7) * 2;
}

struct Bars {
    uint64_t* full; uint64_t* empty; uint64_t* fready; uint64_t* eready; uint64_t* gready;
    __device__ explicit Bars(unsigned char* base) {
        uint64_t* b = reinterpret_cast<uint64_t*>(base + Smem::BAR);
        full = b; empty = b + NST; fready = b + 2 * NST; eready = fready + NF; gready = eready + NF;
    }
};

// one-time setup by the whole CTA: barriers + the whole tensor memory
__device__ inline uint32_t setup(unsigned char* base) {
    Bars B(base);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(B.full + i, 1); tc::mbar_init(B.empty + i, 1); }
        for (int i = 0; i < NF; ++i) { tc::mbar_init(B.fready + i, 1); tc::mbar_init(B.eready + i, 128); }
        tc::mbar_init(B.gready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int e = tid; e < NF * E_BYTES / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(base + Smem::E)[e] = 0u;
    for (int e = tid; e < B_BYTES / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(base + Smem::BM)[e] = 0u;
    tc::fence_proxy_async();
    if ((tid >> 5) == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(tc::smem_u32(base + Smem::TMEM_PTR)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    return *reinterpret_cast<volatile uint32_t*>(base + Smem::TMEM_PTR);
}
__device__ inline void teardown(uint32_t tmem_base) {
    __syncthreads();
    if ((threadIdx.x >> 5) == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

struct State {
    uint32_t tiles = 0, ticks = 0;       // global tile counter (F / E buffer = tiles & 3), passes done
    uint32_t nst = NST;                  // ring stages in use (<= min(NST, 4 nkc))
    uint32_t slot = 0, use = 0;          // ring position of the next sub-tile
};

// The pass.  Before the call the CTA has written the coefficient operand (b_off) and executed
// fence_proxy_async + a barrier.  Returns with the likelihood gradient wrt the nkc*64 coefficient columns in
// `gout` (global memory, float [2][NCH][KWT]: the two row-parity halves, summed by the reader) and the lp partial
// sums per epilogue warp in `lpw` ([8][NCH] doubles).  All tc::NTHREADS threads call it; the caller follows with
// a barrier over them.  ksteps = ceil((D+1)/16) MMA K-steps over all sub-tiles.
template <class Prologue>
__device__ inline void pass(unsigned char* base, uint32_t tmem_base, const CUtensorMap* tmap, State& st,
                            int64_t row_begin, int n_rows, int nkc, int ksteps, const float* __restrict__ yglob,
                            double* lpw, float* __restrict__ gout, Prologue&& prologue) {
    Bars B(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (n_rows + TILE_M - 1) / TILE_M;
    const uint32_t t0 = st.tiles;
    const uint32_t nst = st.nst;
    const uint32_t xs = tc::smem_u32(base + Smem::X);
    constexpr uint32_t IDESC1 = tc::make_idesc(128, NB1, 0);
    constexpr uint32_t IDESC2 = tc::make_idesc(128, 2 * NCH, 1);

    if (warp == 10) {
        // ===== TMA producer =====
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t)
            for (int j = 0; j < nkc; ++j, tc::ring_next(nst, slot, use)) {
                if (use > 0) tc::mbar_wait(B.empty + slot, (use - 1) & 1);     // GEMM2 of the previous user is done
                if (tc::elect_one()) {
                    tc::mbar_expect_tx(B.full + slot, TILE_BYTES);
                    tc::tma_load_2d(base + Smem::X + slot * TILE_BYTES, tmap, B.full + slot, j * KW,
                                    (int)(row_begin + (int64_t)t * TILE_M));
                }
                __syncwarp();
            }
    } else if (warp == 8) {
        // ===== GEMM1 issuer: F(t) = sum over the sub-tiles =====
        const uint64_t bm_d = tc::make_desc(tc::smem_u32(base + Smem::BM), 1024, 128, 0);
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t gt = t0 + t, fb = gt & (NF - 1);
            if (t >= NF) tc::mbar_wait(B.eready + fb, ((gt >> 2) - 1) & 1);    // epilogue(gt - 4) has read F buffer fb
            const uint32_t dF = tmem_base + fb * NB1;
            for (int j = 0; j < nkc; ++j, tc::ring_next(nst, slot, use)) {
                tc::mbar_wait(B.full + slot, use & 1);
                tc::tc_fence_after();
                const uint64_t xa_d = tc::make_desc(xs + slot * TILE_BYTES, 16, 1024, 2);
                const int kst = min(KW / 16, ksteps - j * (KW / 16));
                if (tc::elect_one()) {
                    for (int ks = 0; ks < kst; ++ks)          // start-address field is in 16-byte units
                        tc::umma(dF, xa_d + (uint64_t)(ks * 2), bm_d + (uint64_t)((j * (KW / 16) + ks) * 128), IDESC1,
                                 (j | ks) != 0);
                    if (j == nkc - 1) tc::umma_commit(B.fready + fb);
                }
                __syncwarp();
            }
        }
    } else if (warp == 9) {
        // ===== GEMM2 issuer: G_j += X_sub_j' E(t) =====
        const uint32_t eb0 = tc::smem_u32(base + Smem::E);
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t gt = t0 + t, eb = gt & (NF - 1);
            tc::mbar_wait(B.eready + eb, (gt >> 2) & 1);
            tc::tc_fence_after();
            const uint64_t ea_d = tc::make_desc(eb0 + eb * E_BYTES, 1024, 128, 0);
            for (int j = 0; j < nkc; ++j, tc::ring_next(nst, slot, use)) {
                // A = X_sub' (MN-major, SW128): second 64-element M block = 16 rows further (LBO 2048 B),
                // 8-row K groups 1024 B apart (SBO)
                const uint64_t xa_d = tc::make_desc(xs + slot * TILE_BYTES, 2048, 1024, 2);
                if (tc::elect_one()) {
#pragma unroll
                    for (int i = 0; i < TILE_M / 32; ++i)
                        // 32 rows further: A + 4096 B, B + 2048 B
                        tc::umma(tmem_base + G_COL0 + j * NB1, xa_d + (uint64_t)(i * 256), ea_d + (uint64_t)(i * 128), IDESC2,
                                 i > 0 ? 1u : (t > 0 ? 1u : 0u));
                    tc::umma_commit(B.empty + slot);
                }
                __syncwarp();
            }
        }
        if (tc::elect_one()) tc::umma_commit(B.gready);
        __syncwarp();
    } else {
        // ===== epilogue: one row of the tile per thread, 32 chain columns in two halves =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int grp = warp >> 2;                    // tiles t with (t & 1) == grp
        const int r = q * 32 + lane;                  // row within the tile
        prologue();                                   // independent work that hides the pipeline fill
        float lpacc[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) lpacc[c] = 0.0f;
        float y_next = (grp * TILE_M + r < n_rows) ? yglob[row_begin + grp * TILE_M + r] : 0.0f;
        for (int t = grp; t < n_tiles; t += 2) {
            const uint32_t gt = t0 + t, fb = gt & (NF - 1);
            const int row = t * TILE_M + r;
            const float keep = row < n_rows ? 1.0f : 0.0f;
            const float yv = y_next;
            {
                const int rn = row + 2 * TILE_M;
                y_next = yglob[row_begin + (rn < n_rows ? rn : n_rows - 1)];
            }
            tc::mbar_wait(B.fready + fb, (gt >> 2) & 1);
            tc::tc_fence_after();
            unsigned char* eb = base + Smem::E + fb * E_BYTES;
            const uint32_t fa = tmem_base + ((uint32_t)(q * 32) << 16) + fb * NB1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t fr[16], fl[16];
                tc::TmemPair<16>::ld(fa + h * 16, fa + NCH + h * 16, fr, fl);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float e;
                    const float l = logit_terms(__uint_as_float(fr[c]) + __uint_as_float(fl[c]), yv, e);
                    lpacc[h * 16 + c] = fmaf(keep, l, lpacc[h * 16 + c]);
                    *reinterpret_cast<__nv_bfloat16*>(eb + e_off(h * 16 + c, r)) = __float2bfloat16(keep * e);
                }
            }
            tc::tc_fence_before();
            tc::fence_proxy_async();
            tc::mbar_arrive(B.eready + fb);
        }
        // lp partial sums (fp64 across the warp)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const double v = warp_sum((double)lpacc[c]);
            if (lane == 0) lpw[warp * NCH + c] = v;
        }
        // G read-out.  Accumulator of chunk j [128 lanes x 64 columns]: lanes 0..63 x columns 0..31 hold the sum over
        // the rows with (row & 16) == 0, lanes 64..127 x columns 32..63 the sum over the others; lane & 63 is the
        // coefficient column within the chunk.  Warp (q, grp) reads 16 chain columns of its lane quarter.
        tc::mbar_wait(B.gready, st.ticks & 1);
        tc::tc_fence_after();
        const int hsel = q >> 1;
        for (int j = 0; j < nkc; ++j) {
            uint32_t gr[16];
            tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + G_COL0 + j * NB1 + hsel * NCH + grp * 16, gr);
            float* go = gout + (size_t)(hsel * NCH + grp * 16) * KWT + j * KW + (q & 1) * 32 + lane;
#pragma unroll
            for (int c = 0; c < 16; ++c) go[(size_t)c * KWT] = __uint_as_float(gr[c]);
        }
        tc::tc_fence_before();
    }
    st.tiles += (uint32_t)n_tiles;
    st.ticks += 1;
    {   // advance the ring position by n_tiles * nkc sub-tiles
        const uint32_t adv = st.slot + (uint32_t)(n_tiles * nkc);
        st.use += adv / nst;
        st.slot = adv % nst;
    }
}

}  // namespace tcw
