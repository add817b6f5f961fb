// Shared device helpers for libepgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EPG_WARP 32

// A cooperating thread group = the whole CTA.  Single-warp CTAs (small d) get a
// __syncwarp instead of a barrier.
struct Grp {
    int tid, n;
    __device__ __forceinline__ Grp() : tid(threadIdx.x), n(blockDim.x) {}
    __device__ __forceinline__ Grp(int t, int nn) : tid(t), n(nn) {}     // e.g. one warp of the CTA
    __device__ __forceinline__ void sync() const {
        if (n <= EPG_WARP) __syncwarp(); else __syncthreads();
    }
};

// ---- packed lower-triangular storage, column by column ---------------------
// element (i,j), i>=j, of a d x d symmetric/lower matrix
__device__ __forceinline__ int pk(int i, int j, int d) { return j * d - (j * (j - 1)) / 2 + (i - j); }
__device__ __forceinline__ int pk_col(int j, int d) { return j * d - (j * (j - 1)) / 2 - j; }  // + i
__host__ __device__ __forceinline__ int pk_size(int d) { return d * (d + 1) / 2; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` is shared scratch of >= 33 doubles.  Result to all threads.
__device__ __forceinline__ double block_sum(const Grp& g, double v, double* red) {
    v = warp_sum(v);
    if (g.n <= EPG_WARP) return v;
    const int w = g.tid >> 5, l = g.tid & 31, nw = (g.n + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (l < nw) ? red[l] : 0.0;
        t = warp_sum(t);
        if (l == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
