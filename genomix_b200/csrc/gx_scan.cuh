// gx_scan.cuh -- block-level scan/reduce helpers and the small single-CTA scan over tile sums.
#pragma once
#include "gx_internal.cuh"

namespace gx {

// Exclusive scan of one u64 per thread across a CTA of NT threads (NT multiple of 32, <= 1024).
// Returns the exclusive prefix; *total gets the CTA-wide sum (valid in every thread).
template <int NT>
__device__ __forceinline__ u64 block_scan_excl(u64 v, u64* total) {
    __shared__ u64 warp_sums[NT / 32];
    __shared__ u64 block_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u64 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u64 ws = (lane < NT / 32) ? warp_sums[lane] : 0ull;
        u64 wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < NT / 32) warp_sums[lane] = wi - ws;  // exclusive warp offsets
        if (lane == 31) block_total = wi;
    }
    __syncthreads();
    const u64 res = warp_sums[wid] + incl - v;
    *total = block_total;
    __syncthreads();  // allow back-to-back calls
    return res;
}

template <int NT>
__device__ __forceinline__ u64 block_reduce_sum(u64 v) {
    __shared__ u64 red[NT / 32];
    __shared__ u64 out;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        u64 t = (lane < NT / 32) ? red[lane] : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) out = t;
    }
    __syncthreads();
    const u64 r = out;
    __syncthreads();
    return r;
}

// In-place exclusive scan of `n` u64 tile sums by ONE CTA of 1024 threads; total written to *total.
// n is (array length / tile size), i.e. small; this kernel is never on the roofline.
static __global__ void __launch_bounds__(1024) scan_tile_sums_kernel(u64* __restrict__ sums, u64 n, u64* __restrict__ total) {
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += 1024) {
        const u64 i = base + threadIdx.x;
        const u64 v = (i < n) ? sums[i] : 0ull;
        u64 tot;
        const u64 ex = block_scan_excl<1024>(v, &tot);
        const u64 carry = carry_s;
        if (i < n) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}


// Device-wide exclusive scan of a u32 array (three launches: tile sums, scan_tile_sums_kernel, tile scan).
static constexpr int TS_THREADS = 256;
static constexpr int TS_TILE = TS_THREADS * 4;

static __global__ void __launch_bounds__(TS_THREADS) tile_sum_u32_kernel(const u32* __restrict__ in, u64 n,
                                                                  u64* __restrict__ tile_sums) {
    const u64 base = (u64)blockIdx.x * TS_TILE + (u64)threadIdx.x * 4;
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (base + i < n) s += in[base + i];
    const u64 tot = block_reduce_sum<TS_THREADS>(s);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// out has n + 1 entries; out[n] = grand total
static __global__ void __launch_bounds__(TS_THREADS) tile_scan_u32_kernel(const u32* __restrict__ in, u64 n,
                                                                   const u64* __restrict__ tile_base,
                                                                   u32* __restrict__ out) {
    const u64 base = (u64)blockIdx.x * TS_TILE + (u64)threadIdx.x * 4;
    u32 v[4];
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (base + i < n) ? in[base + i] : 0u; s += v[i]; }
    u64 tot;
    u64 run = tile_base[blockIdx.x] + block_scan_excl<TS_THREADS>(s, &tot);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) out[base + i] = (u32)run;
        run += v[i];
        if (base + i + 1 == n) out[n] = (u32)run;
    }
}

}  // namespace gx
