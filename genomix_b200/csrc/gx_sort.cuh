// gx_sort.cuh -- optional: the dense node list in KmerPointable order (gx_config.sort_output).
//
// The reference's part files hold their records sorted by key, because its group-by is pre-clustered on the output of
// ExternalSortOperatorDescriptor with KmerPointable.compare (genomix-hyracks/.../data/primitive/KmerPointable.java:94-107:
// unsigned byte-wise comparison of the Kmer bytes = integer order of the k-letter value). Nothing downstream depends on that
// order (the Pregelix loader hashes vertices by key), so the build does not pay for it by default; with sort_output the node
// list that emit_scan produced in table-slot order is permuted into key order before the records are written:
//
//   LSD radix sort of the node indices, one pass per key byte (ceil(k/4) passes), stable: a warp owns a tile of consecutive
//   items and ranks them chunk by chunk (match_any), tile bases come from a scan of the per-tile digit histograms laid out
//   digit-major;  then a gather of the node list in the new order and a scan of the record sizes for the new offsets.
#pragma once
#include "gx_emit.cuh"

namespace gx {

static constexpr int RS_WARPS = 8;                 // warps per CTA
static constexpr u32 RS_TILE = 8192;               // items per warp tile

struct SortArgs {
    const u64* dense; int k;       // node list: (KW key words, value word) per node
    u64 n;
    const u32* perm_in; u32* perm_out;   // node indices, in / out of one pass (perm_in == nullptr: identity)
    u32 byte;                      // key byte of this pass, 0 = least significant
    u32* hist;                     // [256][n_tiles] digit-major per-tile counts -> exclusive bases after the scan
    u32 n_tiles;
};

template <int KW>
__device__ __forceinline__ u32 key_byte(const u64* __restrict__ dense, u64 node, u32 byte) {
    return (u32)(dense[node * (KW + 1) + (byte >> 3)] >> (8u * (byte & 7u))) & 0xffu;
}

template <int KW>
__global__ void __launch_bounds__(RS_WARPS * 32) sort_hist_kernel(SortArgs a) {
    __shared__ u32 h[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tile = blockIdx.x * RS_WARPS + warp;
    for (int d = lane; d < 256; d += 32) h[warp][d] = 0;
    __syncwarp();
    if (tile < a.n_tiles) {
        const u64 lo = (u64)tile * RS_TILE, hi = min(lo + RS_TILE, a.n);
        for (u64 i = lo + lane; i < hi; i += 32) {
            const u64 node = a.perm_in ? a.perm_in[i] : i;
            atomicAdd(&h[warp][key_byte<KW>(a.dense, node, a.byte)], 1u);
        }
        __syncwarp();
        for (int d = lane; d < 256; d += 32) a.hist[(size_t)d * a.n_tiles + tile] = h[warp][d];
    }
}

// exclusive scan of the 256 * n_tiles counts in place (digit-major order = the order of the sorted output); one CTA
static __global__ void __launch_bounds__(1024) sort_scan_kernel(u32* __restrict__ hist, u64 n) {
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    constexpr int PER = 8;
    for (u64 base = 0; base < n; base += 1024 * PER) {
        u32 v[PER];
        u64 s = 0;
        const u64 i0 = base + (u64)threadIdx.x * PER;
#pragma unroll
        for (int j = 0; j < PER; ++j) { v[j] = i0 + j < n ? hist[i0 + j] : 0u; s += v[j]; }
        u64 tot;
        u64 run = carry_s + block_scan_excl<1024>(s, &tot);
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (i0 + j < n) hist[i0 + j] = (u32)run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s += tot;
        __syncthreads();
    }
}

template <int KW>
__global__ void __launch_bounds__(RS_WARPS * 32) sort_scatter_kernel(SortArgs a) {
    __shared__ u32 off[RS_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 tile = blockIdx.x * RS_WARPS + warp;
    if (tile >= a.n_tiles) return;
    for (int d = lane; d < 256; d += 32) off[warp][d] = a.hist[(size_t)d * a.n_tiles + tile];
    __syncwarp();
    const u32 lane_lt = (1u << lane) - 1u;
    const u64 lo = (u64)tile * RS_TILE, hi = min(lo + RS_TILE, a.n);
    for (u64 i0 = lo; i0 < hi; i0 += 32) {   // chunks in order, lanes in order: stable
        const u64 i = i0 + lane;
        const bool live = i < hi;
        u32 node = 0, d = 256 + lane;        // dead lanes: distinct pseudo digits, matched with nobody
        if (live) {
            node = a.perm_in ? a.perm_in[i] : (u32)i;
            d = key_byte<KW>(a.dense, node, a.byte);
        }
        const u32 same = __match_any_sync(0xffffffffu, d);
        if (live) {
            a.perm_out[off[warp][d] + (u32)__popc(same & lane_lt)] = node;
        }
        __syncwarp();
        if (live && (same & lane_lt) == 0) off[warp][d] += (u32)__popc(same);   // the group's first lane advances the digit
        __syncwarp();
    }
}

// node list, head-group references and record sizes in the new order
template <int KW>
__global__ void __launch_bounds__(256) sort_gather_kernel(const u64* __restrict__ dense, const u32* __restrict__ dense_h,
                                                          const u64* __restrict__ rec_offsets, const u32* __restrict__ perm, u64 n,
                                                          u64* __restrict__ dense_out, u32* __restrict__ dense_h_out, u32* __restrict__ sizes) {
    constexpr int DW = KW + 1;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 src = perm[i];
#pragma unroll
    for (int j = 0; j < DW; ++j) dense_out[i * DW + j] = dense[src * DW + j];
    if (dense[src * DW + KW] & HEADS_FLAG) dense_h_out[i] = dense_h[src];
    sizes[i] = (u32)(rec_offsets[src + 1] - rec_offsets[src]);
}

// out[i] = tile_base[tile] + exclusive prefix of sizes inside the tile; out[n] = total (sizes: u32, offsets: u64)
static __global__ void __launch_bounds__(TS_THREADS) tile_scan_u32_to_u64_kernel(const u32* __restrict__ in, u64 n, const u64* __restrict__ tile_base,
                                                                          u64* __restrict__ out) {
    const u64 base = (u64)blockIdx.x * TS_TILE + (u64)threadIdx.x * 4;
    u32 v[4];
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (base + i < n) ? in[base + i] : 0u; s += v[i]; }
    u64 tot;
    u64 run = tile_base[blockIdx.x] + block_scan_excl<TS_THREADS>(s, &tot);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
        if (base + i + 1 == n) out[n] = run;
    }
}

}  // namespace gx
