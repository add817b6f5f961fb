// Tensor-core (tcgen05 + TMEM + TMA) likelihood pass of the tilted densities,
// single-group sites, D+1 <= 64 inputs, <= 16 chains.
//
//   per 128-row tile of the site's design matrix (bf16, 64 columns = 128 B rows,
//   column D holds 1.0, TMA-loaded with the 128-byte swizzle):
//     GEMM1  F[128 x 16]  = X_tile[128 x 64] . (B_hi + B_lo)'[64 x 16]    (K-major A, M=128)
//     epilogue (4 warps, one row per thread): tcgen05.ld F -> e = y - sigmoid(f),
//              lp += y f - softplus(f);  E (bf16) -> shared memory
//     GEMM2  G[64 x 16] += X_tile'[64 x 128] . E[128 x 16]   (the SAME smem tile read
//              as an MN-major A operand, M=64, accumulated in TMEM over all tiles)
//   X is read once per tick from L2/HBM; both contractions run on the tensor
//   cores; the coefficient matrix is split B = B_hi + B_lo (two bf16 MMAs) so that
//   the log-density sees fp32-accurate coefficients.
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor) and the canonical layouts documented in
// cute/atom/mma_traits_sm100.hpp (vendored CUTLASS headers).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

constexpr int TILE_M = 128;          // rows per tile
constexpr int KW = 64;               // padded input columns (bf16) == one 128-byte swizzle row
constexpr int NCH = 16;              // chains padded (N of both GEMMs)
constexpr int NST = 8;               // max X tile stages (TMA runs this far ahead); the launch picks nst in NF..NST
                                     // (State::nst): 8 when one site owns the CTA, fewer in the two-site kernel
constexpr int NF = 4;                // F accumulator buffers in TMEM (GEMM1 runs this far ahead of the epilogue);
                                     // NF <= nst, so that a barrier is never more than one phase ahead of a waiter
constexpr int NE = NST;              // E operand buffers in smem, one per X stage: a stage's E buffer is free
                                     // exactly when its X tile is (GEMM2 of the previous user done), which the TMA
                                     // producer has already waited for -> no separate "E free" barrier
constexpr int TILE_BYTES = TILE_M * KW * 2;          // 16384
constexpr int NB1 = 2 * NCH;         // GEMM1 N: columns [0,16) = hi parts, [16,32) = lo parts of the coefficients
constexpr int B_BYTES = NB1 * KW * 2;                // 4096  (coefficient operand, hi | lo)
constexpr int E_BYTES = NCH * TILE_M * 2;            // 4096  (one E buffer)
constexpr int TMEM_COLS = 256;                       // F buffers [0, NF*32), G [NF*32, NF*32+16)
constexpr int NBAR = 2 * NST + NF + NE + 1;
constexpr int NTHREADS = 352;        // warps 0-7: epilogue (two halves x four TMEM lane quarters), 8: GEMM1 issue,
                                     // 9: GEMM2 issue (+ TMEM alloc), 10: TMA producer
__host__ __device__ constexpr int chain_col(int c) { return (c & 1) * (NCH / 2) + (c >> 1); }

struct Smem {            // offsets relative to a 1024-byte aligned base
    static constexpr int BM = 0;                     // coefficient operand [32 x 64] (hi rows | lo rows)
    static constexpr int GOUT = BM + B_BYTES;        // float [NCH][KW]
    static constexpr int BAR = GOUT + NCH * KW * 4;
    static constexpr int TMEM_PTR = BAR + NBAR * 8;
    static constexpr int E = (TMEM_PTR + 16 + 1023) & ~1023;          // nst E buffers
    __host__ __device__ static constexpr int X(int nst) { return E + nst * E_BYTES; }     // nst X tiles (1024-aligned)
    __host__ __device__ static constexpr int total(int nst) { return X(nst) + nst * TILE_BYTES; }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// wait for the phase with parity `par`; traps instead of hanging the GPU if it never completes
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t par) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    asm volatile(                       // fast path: the phase has usually completed already
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(a), "r"(par) : "memory");
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(a), "r"(par) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
// one lane of a converged warp (cute::elect_one_sync)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
// non-blocking probe of the phase with parity `par`
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t par) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(par) : "memory");
    return done != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// ---- shared-memory matrix descriptors (SmemDescriptor, version 1) ----
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                    // version_ = 1 (Blackwell)
    d |= (uint64_t)(layout & 7) << 61;         // 0 none(interleave), 2 = SWIZZLE_128B
    return d;
}
// ---- instruction descriptor, kind::f16, bf16 x bf16 -> f32 ----
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// byte offset of element (n, k) of a K-major INTERLEAVE (no swizzle) operand with
// NCH rows: 8x(16 B) core matrices, the two row groups adjacent (SBO = 128 B),
// K chunks of 8 elements LBO = 256 B apart
__device__ __forceinline__ int interleave_off(int n, int k) {
    return (k >> 3) * 256 + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
}
// the same for the 32-row coefficient operand of GEMM1 (four row groups: LBO = 512 B)
__device__ __forceinline__ int interleave_off32(int n, int k) {
    return (k >> 3) * 512 + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
}

struct Bars {
    uint64_t* full; uint64_t* empty; uint64_t* fready; uint64_t* eready; uint64_t* gready;
    __device__ explicit Bars(unsigned char* base) {
        uint64_t* b = reinterpret_cast<uint64_t*>(base + Smem::BAR);
        full = b; empty = b + NST; fready = b + 2 * NST; eready = fready + NF; gready = eready + NE;
    }
};

// one-time setup by the whole CTA (called once per kernel): barriers + TMEM
__device__ inline uint32_t setup(unsigned char* base, int nst) {
    Bars B(base);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(B.full + i, 1); mbar_init(B.empty + i, 1); }
        for (int i = 0; i < NF; ++i) mbar_init(B.fready + i, 1);
        for (int i = 0; i < NE; ++i) mbar_init(B.eready + i, 256);
        mbar_init(B.gready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int e = tid; e < nst * E_BYTES / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(base + Smem::E)[e] = 0u;
    if ((tid >> 5) == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(smem_u32(base + Smem::TMEM_PTR)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *reinterpret_cast<volatile uint32_t*>(base + Smem::TMEM_PTR);
}
__device__ inline void teardown(uint32_t tmem_base) {
    __syncthreads();
    if ((threadIdx.x >> 5) == 9)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// Running counters of the pipelines (they persist over the ticks of a kernel).
struct State {
    uint32_t tiles = 0, ticks = 0;       // tiles: global index of the next tile (F buffer = tiles % NF)
    uint32_t nst = NST;                  // X/E stages in use (NF <= nst <= NST)
    uint32_t slot = 0, use = 0;          // tiles % nst, tiles / nst, kept incrementally (nst need not be a power of two)
    __device__ void set_stages(int n) { nst = (uint32_t)(n < NF ? NF : (n > NST ? NST : n)); }
};
// (slot, use) of the tile after / NF tiles before the given one
__device__ __forceinline__ void ring_next(uint32_t nst, uint32_t& slot, uint32_t& use) {
    if (++slot == nst) { slot = 0; ++use; }
}
#ifdef EPG_TC_PROFILE
__device__ long long g_prof[8];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(i, a, b) if (threadIdx.x == 0 && blockIdx.x == 0) g_prof[i] += (b) - (a)
__device__ long long g_prof2[8];
#define PROF2_ADD(i, a, b) if (blockIdx.x == 0) tc::g_prof2[i] += (b) - (a)
#else
#define PROF_T(var)
#define PROF_ADD(i, a, b)
#define PROF2_ADD(i, a, b)
#endif

// The pass proper.  Before the call the CTA has written the bf16 coefficient
// operands (B_hi, B_lo) with interleave_off(chain_col(c), k) and executed
// fence_proxy_async + __syncthreads.  Returns with G (likelihood gradient wrt the
// 64 coefficient columns, [NCH columns][KW] floats) in smem at Smem::GOUT and the
// lp partial sums per warp in `lpw` ([8 warps][NCH/2] doubles; warp w covers the
// columns of half w>>2).  All NTHREADS threads must call it; the caller follows
// with a __syncthreads.  NCPH = columns each epilogue half really processes.
template <int NCPH, class Prologue>
__device__ inline void pass(unsigned char* base, uint32_t tmem_base, const CUtensorMap* tmap, State& st,
                            int64_t row_begin, int n_rows, int ksteps, const float* __restrict__ yglob, double* lpw,
                            Prologue&& prologue) {
    Bars B(base);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (n_rows + TILE_M - 1) / TILE_M;
    const uint32_t t0 = st.tiles;                  // global index of this tick's first tile
    const uint32_t nst = st.nst;
    const int x_off = Smem::X((int)nst);
    const uint32_t xs = smem_u32(base + x_off);
    constexpr uint32_t IDESC1 = make_idesc(128, NB1, 0);
    constexpr uint32_t IDESC2 = make_idesc(64, NCH, 1);

    if (warp == 10) {
        // ===== TMA producer (warp runs converged; one elected lane issues) =====
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t, ring_next(nst, slot, use)) {
            if (use > 0) mbar_wait(B.empty + slot, (use - 1) & 1);     // GEMM2 of the previous user is done
            if (elect_one()) {
                mbar_expect_tx(B.full + slot, TILE_BYTES);
                tma_load_2d(base + x_off + slot * TILE_BYTES, tmap, B.full + slot, 0,
                            (int)(row_begin + (int64_t)t * TILE_M));
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ===== GEMM1 issuer (warp runs converged; one elected lane issues) =====
        const uint64_t bm_d = make_desc(smem_u32(base + Smem::BM), 512, 128, 0);
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t, ring_next(nst, slot, use)) {
            const uint32_t gt = t0 + t, fb = gt % NF;
            // F buffer fb was last read by epilogue(gt - NF), whose E buffer sits NF stages back in the ring
            if (t >= NF) {
                const bool wrap = slot < (uint32_t)NF;
                mbar_wait(B.eready + (wrap ? slot + nst - NF : slot - NF), (wrap ? use - 1 : use) & 1);
            }
            mbar_wait(B.full + slot, use & 1);
            tc_fence_after();
            const uint64_t xa_d = make_desc(xs + slot * TILE_BYTES, 16, 1024, 2);
            const uint32_t dF = tmem_base + fb * NB1;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KW / 16; ++ks)          // start-address field is in 16-byte units
                    if (ks < ksteps) umma(dF, xa_d + (uint64_t)(ks * 2), bm_d + (uint64_t)(ks * 64), IDESC1, ks > 0);
                umma_commit(B.fready + fb);
            }
            __syncwarp();
        }
    } else if (warp == 9) {
        // ===== GEMM2 issuer (warp runs converged; one elected lane issues) =====
        const uint32_t eb = smem_u32(base + Smem::E);
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t, ring_next(nst, slot, use)) {
            const uint32_t eb_i = slot;
            PROF_T(m0);
            mbar_wait(B.eready + eb_i, use & 1);
            tc_fence_after();
            PROF_T(m1);
            const uint64_t xa_d = make_desc(xs + slot * TILE_BYTES, 1024, 1024, 2);
            const uint64_t ea_d = make_desc(eb + eb_i * E_BYTES, 256, 128, 0);
            const uint32_t acc0 = t > 0 ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < TILE_M / 16; ++ks)
                    // A = X_tile' (MN-major, SW128): 16 rows further = +2048 B;  B = E: 2 K-chunks = +512 B
                    umma(tmem_base + NF * NB1, xa_d + (uint64_t)(ks * 128), ea_d + (uint64_t)(ks * 32), IDESC2,
                         ks > 0 ? 1u : acc0);
                umma_commit(B.empty + slot);
            }
            __syncwarp();
            PROF_T(m3);
            if (lane == 0) { PROF2_ADD(0, m0, m1); PROF2_ADD(1, m1, m3); PROF2_ADD(7, 0, 1); }
        }
        if (elect_one()) umma_commit(B.gready);
        __syncwarp();
    } else {
        // ===== epilogue: one row of the tile per thread, half of the columns per warp group =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = warp >> 2;                   // columns [half*8, half*8+8)
        const int r = q * 32 + lane;                  // row within the tile
        const int col0 = half * (NCH / 2);
        prologue();                                   // independent work that hides the pipeline fill
        float lpacc[NCPH];
#pragma unroll
        for (int c = 0; c < NCPH; ++c) lpacc[c] = 0.0f;
        float y_next = (r < n_rows) ? yglob[row_begin + r] : 0.0f;
        uint32_t ebi = st.slot, euse = st.use;
        for (int t = 0; t < n_tiles; ++t, ring_next(nst, ebi, euse)) {
            const uint32_t gt = t0 + t, fb = gt % NF;
            const int row = t * TILE_M + r;
            const float keep = row < n_rows ? 1.0f : 0.0f;
            const float yv = y_next;
            {   // prefetch the next tile's response while this tile is processed
                const int rn = row + TILE_M;
                y_next = (rn < n_rows) ? yglob[row_begin + rn] : 0.0f;
            }
            PROF_T(p0);
            mbar_wait(B.fready + fb, (gt / NF) & 1);
            tc_fence_after();
            PROF_T(p1);
            uint32_t fr[8], fl[8];
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + fb * NB1 + col0, fr);
            tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + fb * NB1 + NCH + col0, fl);
            tc_fence_before();
            PROF_T(p2);
            __nv_bfloat16 ev[NCPH];
#pragma unroll
            for (int c = 0; c < NCPH; ++c) {
                float e;
                const float l = logit_terms(__uint_as_float(fr[c]) + __uint_as_float(fl[c]), yv, e);
                lpacc[c] = fmaf(keep, l, lpacc[c]);
                ev[c] = __float2bfloat16(keep * e);
            }
            PROF_T(p3);
            PROF_T(p4);                       // (E buffer ebi is free: see NE)
            unsigned char* eb = base + Smem::E + ebi * E_BYTES;
#pragma unroll
            for (int c = 0; c < NCPH; ++c)
                *reinterpret_cast<__nv_bfloat16*>(eb + interleave_off(col0 + c, r)) = ev[c];
            PROF_T(p5);
            fence_proxy_async();
            PROF_T(p6);
            mbar_arrive(B.eready + ebi);
            PROF_T(p7);
            PROF_ADD(0, p0, p1); PROF_ADD(1, p1, p2); PROF_ADD(2, p2, p3); PROF_ADD(3, p3, p4);
            PROF_ADD(4, p4, p5); PROF_ADD(5, p5, p6); PROF_ADD(6, p6, p7);
#ifdef EPG_TC_PROFILE
            if (threadIdx.x == 0 && blockIdx.x == 0) g_prof[7] += 1;
#endif
        }
        // lp partial sums (fp64 across the warp)
#pragma unroll
        for (int c = 0; c < NCPH; ++c) {
            const double v = warp_sum((double)lpacc[c]);
            if (lane == 0) lpw[warp * (NCH / 2) + c] = v;
        }
        // G: rows 16q..16q+15 of the 64 coefficient columns sit in lanes 0..15 of quarter q
        mbar_wait(B.gready, st.ticks & 1);
        tc_fence_after();
        uint32_t gr[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + NF * NB1 + col0, gr);
        tc_fence_before();
        float* gout = reinterpret_cast<float*>(base + Smem::GOUT);
        if (lane < 16) {
#pragma unroll
            for (int c = 0; c < NCH / 2; ++c) gout[(col0 + c) * KW + q * 16 + lane] = __uint_as_float(gr[c]);
        }
    }
    st.tiles += (uint32_t)n_tiles;
    st.ticks += 1;
    {   // advance the ring position by n_tiles
        const uint32_t adv = st.slot + (uint32_t)n_tiles;
        st.use += adv / nst;
        st.slot = adv % nst;
    }
}

}  // namespace tc
