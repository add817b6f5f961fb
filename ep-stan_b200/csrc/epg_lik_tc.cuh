// Tensor-core (tcgen05 + TMEM + TMA) likelihood pass of the tilted densities,
// single-group sites, D+1 <= 64 inputs, <= 16 chains.
//
//   per 128-row tile of the site's design matrix (bf16, 64 columns = 128 B rows,
//   column D holds 1.0, TMA-loaded with the 128-byte swizzle into a ring of 4..8 stages):
//     GEMM1  F[128 x 32]  = X_tile[128 x 64] . [B_hi | B_lo]'[64 x 32]    (K-major A, M=128, N=32)
//     epilogue (8 warps = two groups that take alternate tiles, one row per thread):
//              tcgen05.ld F_hi, F_lo -> e = y - sigmoid(f), lp += y f - softplus(f);
//              E (bf16) -> shared memory, fence.proxy.async, mbarrier arrive
//     GEMM2  G[64 x 16] += X_tile'[64 x 128] . E[128 x 16]   (the SAME smem tile read as an
//              MN-major A operand, accumulated in TMEM over all tiles); issued as 4 instructions
//              of M=128, N=32, K=16 whose two diagonal accumulator blocks are the products of
//              two 16-row slices each (see e_off)
//   X is read once per tick from L2; both contractions run on the tensor cores; the coefficient
//   matrix is split B = B_hi + B_lo (N dimension of GEMM1) so that the log-density sees
//   fp32-accurate coefficients.  Per tile the X data crosses shared memory three times (TMA write,
//   GEMM1 read, GEMM2 read): that bandwidth, not the tensor pipe, bounds the tile phase.
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
__host__ __device__ constexpr int chain_col(int c) { return c; }

struct Smem {            // offsets relative to a 1024-byte aligned base
    static constexpr int BM = 0;                     // coefficient operand [32 x 64] (hi rows | lo rows)
    static constexpr int GOUT = BM + B_BYTES;        // float [2][NCH][KW]: the two row-parity halves of G (summed by the reader)
    static constexpr int BAR = GOUT + 2 * NCH * KW * 4;
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
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// the hi and lo accumulator columns of one row in a single wait
template <int N> struct TmemPair;
template <> struct TmemPair<4> {
    __device__ static __forceinline__ void ld(uint32_t a_hi, uint32_t a_lo, uint32_t (&h)[4], uint32_t (&l)[4]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%8];\n\t"
            "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4,%5,%6,%7}, [%9];\n\t"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]), "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3])
            : "r"(a_hi), "r"(a_lo) : "memory");
    }
};
template <> struct TmemPair<8> {
    __device__ static __forceinline__ void ld(uint32_t a_hi, uint32_t a_lo, uint32_t (&h)[8], uint32_t (&l)[8]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]), "=r"(h[4]), "=r"(h[5]), "=r"(h[6]), "=r"(h[7]),
              "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]), "=r"(l[4]), "=r"(l[5]), "=r"(l[6]), "=r"(l[7])
            : "r"(a_hi), "r"(a_lo) : "memory");
    }
};
template <> struct TmemPair<16> {
    __device__ static __forceinline__ void ld(uint32_t a_hi, uint32_t a_lo, uint32_t (&h)[16], uint32_t (&l)[16]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]), "=r"(h[4]), "=r"(h[5]), "=r"(h[6]), "=r"(h[7]),
              "=r"(h[8]), "=r"(h[9]), "=r"(h[10]), "=r"(h[11]), "=r"(h[12]), "=r"(h[13]), "=r"(h[14]), "=r"(h[15]),
              "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]), "=r"(l[4]), "=r"(l[5]), "=r"(l[6]), "=r"(l[7]),
              "=r"(l[8]), "=r"(l[9]), "=r"(l[10]), "=r"(l[11]), "=r"(l[12]), "=r"(l[13]), "=r"(l[14]), "=r"(l[15])
            : "r"(a_hi), "r"(a_lo) : "memory");
    }
};
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
// byte offset of E element (chain n, tile row rho) in an E buffer.  GEMM2 runs as 4 instructions of
// M = 128, N = 32, K = 16 per tile: instruction i covers rows [32i, 32i+32); its A operand stacks the X'
// slices of rows 32i..32i+15 (M block 0) and 32i+16..32i+31 (M block 1, LBO = 16 rows), its B operand the
// matching E slices as N block 0 / N block 1 -- the two diagonal blocks of the 128 x 32 accumulator are the
// wanted products, so one instruction does the work of two K = 16 steps.  B is a K-major INTERLEAVE operand:
// 8x(16 B) core matrices, the four 8-row groups 128 B apart (SBO), the two K chunks 512 B apart (LBO).
__device__ __forceinline__ int e_off(int n, int rho) {
    return (rho >> 5) * 1024 + ((rho >> 3) & 1) * 512 + ((((rho >> 4) & 1) << 1) + (n >> 3)) * 128 + (n & 7) * 16 +
           (rho & 7) * 2;
}
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
        for (int i = 0; i < NE; ++i) mbar_init(B.eready + i, 128);      // one epilogue group (4 warps) per tile
        mbar_init(B.gready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int e = tid; e < nst * E_BYTES / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(base + Smem::E)[e] = 0u;
    for (int e = tid; e < B_BYTES / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(base + Smem::BM)[e] = 0u;
    fence_proxy_async();                 // the zeros are read by the tensor core (async proxy)
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
#ifdef EPG_TC_EXPERIMENT
__device__ int g_knobs;          // bit 0: no TMA after the first ring fill, 1: GEMM1 one k-step, 2: GEMM2 one instruction, 3: no logistic math
#define KNOB(b) (tc::g_knobs & (1 << (b)))
#else
#define KNOB(b) 0
#endif
#ifdef EPG_TC_TIMELINE
__device__ long long g_tl[8][64];        // event timeline of one pass (block 0, tick 300): [event][tile]
__device__ long long g_tlx[8];           // pass start, prologue end, tiles done, gready seen, end
#define TL(ev, t) if (blockIdx.x == 0 && st.ticks == 300 && (t) < 64) tc::g_tl[ev][t] = clock64()
#define TLX(i) if (blockIdx.x == 0 && st.ticks == 300) tc::g_tlx[i] = clock64()
#else
#define TL(ev, t)
#define TLX(i)
#endif
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
// lp partial sums per warp in `lpw` ([8 warps][NCH] doubles, columns < NCP).  All
// NTHREADS threads must call it; the caller follows with a barrier over them.
// NCP = chain columns really processed (chains padded to 4, 8 or 16).
template <int NCP, class Prologue>
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
    constexpr uint32_t IDESC2 = make_idesc(128, 2 * NCH, 1);

    if (warp == 10) {
        // ===== TMA producer (warp runs converged; one elected lane issues) =====
        uint32_t slot = st.slot, use = st.use;
        for (int t = 0; t < n_tiles; ++t, ring_next(nst, slot, use)) {
            if (use > 0) mbar_wait(B.empty + slot, (use - 1) & 1);     // GEMM2 of the previous user is done
            if (lane == 0) { TL(0, t); }
            if (elect_one()) {
                if (KNOB(0) && st.tiles + t >= nst) mbar_arrive(B.full + slot);
                else {
                mbar_expect_tx(B.full + slot, TILE_BYTES);
                tma_load_2d(base + x_off + slot * TILE_BYTES, tmap, B.full + slot, 0,
                            (int)(row_begin + (int64_t)t * TILE_M));
                }
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
            if (lane == 0) { TL(1, t); }
            const uint64_t xa_d = make_desc(xs + slot * TILE_BYTES, 16, 1024, 2);
            const uint32_t dF = tmem_base + fb * NB1;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KW / 16; ++ks)          // start-address field is in 16-byte units
                    if (ks < (KNOB(1) ? 1 : ksteps)) umma(dF, xa_d + (uint64_t)(ks * 2), bm_d + (uint64_t)(ks * 64), IDESC1, ks > 0);
                umma_commit(B.fready + fb);
            }
            __syncwarp();
            if (lane == 0) { TL(2, t); }
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
            if (lane == 0) { TL(5, t); }
            // A = X_tile' (MN-major, SW128): second 64-element M block = 16 rows further (LBO 2048 B),
            // 8-row K groups 1024 B apart (SBO); B = E (see e_off)
            const uint64_t xa_d = make_desc(xs + slot * TILE_BYTES, 2048, 1024, 2);
            const uint64_t ea_d = make_desc(eb + eb_i * E_BYTES, 512, 128, 0);
            const uint32_t acc0 = t > 0 ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < (KNOB(2) ? 1 : TILE_M / 32); ++i)
                    // 32 rows further: A + 4096 B, B + 1024 B (start-address field is in 16-byte units)
                    umma(tmem_base + NF * NB1, xa_d + (uint64_t)(i * 256), ea_d + (uint64_t)(i * 64), IDESC2,
                         i > 0 ? 1u : acc0);
                umma_commit(B.empty + slot);
            }
            __syncwarp();
            PROF_T(m3);
            if (lane == 0) { TL(6, t); }
            if (lane == 0) { PROF2_ADD(0, m0, m1); PROF2_ADD(1, m1, m3); PROF2_ADD(7, 0, 1); }
        }
        if (elect_one()) umma_commit(B.gready);
        __syncwarp();
    } else {
        // ===== epilogue: one row of the tile per thread, all NCP chain columns; the two groups of four
        //       warps take alternate tiles, so that two tiles' load -> logistic -> store -> fence chains overlap =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int grp = warp >> 2;                    // tiles t with (t & 1) == grp
        const int r = q * 32 + lane;                  // row within the tile
        if (tid == 0) { TLX(0); }
        prologue();                                   // independent work that hides the pipeline fill
        if (tid == 0) { TLX(1); }
        float lpacc[NCP];
#pragma unroll
        for (int c = 0; c < NCP; ++c) lpacc[c] = 0.0f;
        uint32_t ebi = st.slot, euse = st.use;
        if (grp) ring_next(nst, ebi, euse);
        float y_next = (grp * TILE_M + r < n_rows) ? yglob[row_begin + grp * TILE_M + r] : 0.0f;
        for (int t = grp; t < n_tiles; t += 2, ring_next(nst, ebi, euse), ring_next(nst, ebi, euse)) {
            const uint32_t gt = t0 + t, fb = gt % NF;
            const int row = t * TILE_M + r;
            const float keep = row < n_rows ? 1.0f : 0.0f;
            const float yv = y_next;
            {   // prefetch this group's next tile's response while this tile is processed
                const int rn = row + 2 * TILE_M;
                y_next = yglob[row_begin + (rn < n_rows ? rn : n_rows - 1)];      // (clamped: unused when out of range)
            }
            PROF_T(p0);
            mbar_wait(B.fready + fb, (gt / NF) & 1);
            tc_fence_after();
            PROF_T(p1);
            if (lane == 0 && q == 0) { TL(3, t); }
            uint32_t fr[NCP], fl[NCP];
            TmemPair<NCP>::ld(tmem_base + ((uint32_t)(q * 32) << 16) + fb * NB1,
                              tmem_base + ((uint32_t)(q * 32) << 16) + fb * NB1 + NCH, fr, fl);
            tc_fence_before();
            PROF_T(p2);
            __nv_bfloat16 ev[NCP];
#pragma unroll
            for (int c = 0; c < NCP; ++c) {
                float e;
                float l;
                if (KNOB(3)) { e = __uint_as_float(fr[c]); l = __uint_as_float(fl[c]); }
                else l = logit_terms(__uint_as_float(fr[c]) + __uint_as_float(fl[c]), yv, e);
                lpacc[c] = fmaf(keep, l, lpacc[c]);
                ev[c] = __float2bfloat16(keep * e);
            }
            PROF_T(p3);
            PROF_T(p4);                       // (E buffer ebi is free: see NE)
            unsigned char* eb = base + Smem::E + ebi * E_BYTES;
#pragma unroll
            for (int c = 0; c < NCP; ++c)
                *reinterpret_cast<__nv_bfloat16*>(eb + e_off(c, r)) = ev[c];
            PROF_T(p5);
            fence_proxy_async();
            PROF_T(p6);
            mbar_arrive(B.eready + ebi);
            PROF_T(p7);
            if (lane == 0 && q == 0) { TL(4, t); }
            PROF_ADD(0, p0, p1); PROF_ADD(1, p1, p2); PROF_ADD(2, p2, p3); PROF_ADD(3, p3, p4);
            PROF_ADD(4, p4, p5); PROF_ADD(5, p5, p6); PROF_ADD(6, p6, p7);
#ifdef EPG_TC_PROFILE
            if (threadIdx.x == 0 && blockIdx.x == 0) g_prof[7] += 1;
#endif
        }
        if (tid == 0) { TLX(2); }
        // lp partial sums (fp64 across the warp)
#pragma unroll
        for (int c = 0; c < NCP; ++c) {
            const double v = warp_sum((double)lpacc[c]);
            if (lane == 0) lpw[warp * NCH + c] = v;
        }
        // G read-out.  Accumulator [128 lanes x 32 columns]: lanes 0..63 x columns 0..15 hold the sum over the
        // rows with (row & 16) == 0, lanes 64..127 x columns 16..31 the sum over the others; lane & 63 is the
        // coefficient column.  Warp (q, grp) reads 8 chain columns of its lane quarter.
        mbar_wait(B.gready, st.ticks & 1);
        tc_fence_after();
        if (tid == 0) { TLX(3); }
        const int hsel = q >> 1;
        uint32_t gr[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + NF * NB1 + hsel * NCH + grp * 8, gr);
        tc_fence_before();
        float* gout = reinterpret_cast<float*>(base + Smem::GOUT) + hsel * NCH * KW;
#pragma unroll
        for (int c = 0; c < 8; ++c) gout[(grp * 8 + c) * KW + (q & 1) * 32 + lane] = __uint_as_float(gr[c]);
        if (tid == 0) { TLX(4); }
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
