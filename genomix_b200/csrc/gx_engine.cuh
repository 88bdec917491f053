// gx_engine.cuh -- the KW-templated kernels of the graph-build path and their launchers.
//
//   K1+K2  extract_kernel<KW,EX_UPSERT>  read -> canonical k-mers -> hash-table upsert          (single GPU: fused)
//   K1x    extract_kernel<KW,EX_ROUTE>   same, own keys upserted, the rest bucketed by owner GPU (multi GPU)
//          extract_kernel<KW,EX_FLAT> + partition_flat_kernel                                   (opt-in L2-blocked build)
//   K2x    insert_records_kernel         (key, mask) records -> upsert (received from peers / table regions / spills)
//   K3     heads_* + emit_size/compact/serialise  read-head grouping, sizing, dense node list, Node serialisation
//          graph_stats_kernel, partition_records_kernel, route/rebase_heads
//
// Reference semantics restated by each kernel are cited at the kernel.
#pragma once
#include "gx_internal.cuh"
#include "gx_parse.cuh"
#include "gx_scan.cuh"
#include "gx_table.cuh"

namespace gx {

static constexpr int EX_THREADS = 256;
static constexpr int EX_WARPS = EX_THREADS / 32;
static constexpr int WIN_BYTES = 512;                       // packed letters per warp window: 2048 letters
static constexpr int WIN_WORDS = WIN_BYTES / 8 + GX_MAX_KW + 2;
static constexpr int WIN_POSITIONS = 4 * WIN_BYTES - 160;   // k-mer start positions per window (k <= 128)
#ifndef GX_EX_BATCH
#define GX_EX_BATCH 1   // >1: prefetch-batched variant (measured slower on B200: L2 prefetch pulls whole 128 B lines)
#endif
#ifndef GX_EX_MIN_BLOCKS
#define GX_EX_MIN_BLOCKS 4
#endif
static constexpr int EX_BATCH = GX_EX_BATCH;                // groups of 30 positions in flight per warp
static constexpr int MAX_BUCKETS = 1024;                    // table regions of the L2-blocked build
static constexpr int ROUTE_MAXG = 32;                       // ranks handled by the parallel-reservation routing path
static constexpr int BUCKET_PAD = 16;                       // per-region counters live 128 B apart (one L2 line each)

// Region (bucket) of a key: regions are contiguous slot ranges because slot_of() is monotone in the hash too.
__host__ __device__ __forceinline__ u32 bucket_of(u64 h, u32 n_buckets) {
#ifdef __CUDA_ARCH__
    return (u32)__umul64hi(h, (u64)n_buckets);
#else
    return (u32)(((unsigned __int128)h * n_buckets) >> 64);
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct ExtractArgs {
    const uint8_t* text; u64 n_text;
    const LineDesc* desc; u64 n_lines;
    int k;
    u64* table; u64 capacity;
    void* heads;
    uint8_t* store;
    Counters* ctr;
    // routing (multi-GPU) -- unused by the fused kernel
    u32 n_ranks; u32 rank;
    u64* const* route_keys;             // [n_ranks] -> send bucket of key words (KW per record)
    unsigned short* const* route_meta;  // [n_ranks] -> send bucket of edge masks
    u64* route_count;                   // [n_ranks] records appended so far
    // L2-blocked build (EX_FLAT)
    u64* flat_keys; unsigned short* flat_meta;  // [chunk occurrences] in parse order
    u64* bucket_count; u32 n_buckets;           // records per table region
};

// four text bytes at the 4-byte aligned address `w` -> one packed quad; bytes outside [lo, hi) read as 'A'
__device__ __forceinline__ u32 load_quad(const uint8_t* w, const uint8_t* lo, const uint8_t* hi) {
    u32 x;
    if (w >= lo && w + 4 <= hi) {
        x = __ldg(reinterpret_cast<const u32*>(w));
    } else {
        x = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (w + i >= lo && w + i < hi) x |= (u32)__ldg(w + i) << (8 * i);
    }
    u32 q = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bool ok;
        q |= code_of((x >> (8 * i)) & 0xffu, ok) << (2 * i);
    }
    return q;
}

// letters [a, a+k) of the packed window -> f
template <int KW>
__device__ __forceinline__ void window_kmer(const u64* __restrict__ W, u32 a, int k, u64 (&f)[KW]) {
    const u32 wi = a >> 5;
    const u32 sh = (a & 31u) * 2u;
    u64 lo = W[wi];
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const u64 hi = W[wi + j + 1];
        f[j] = (lo >> sh) | ((hi << 1) << (63u - sh));
        lo = hi;
    }
    f[KW - 1] &= top_word_mask(k);
}

__device__ __forceinline__ u32 window_letter(const u64* __restrict__ W, u32 a) {
    return (u32)(W[a >> 5] >> ((a & 31u) * 2u)) & 3u;
}

// Pack a read's letters into the read store in VKmer byte order (VKmer.java:462-479: letter i at bits
// 2*(i%4) of byte nb-1-i/4); non-ACGT letters pack as A (GeneCode.java:29-50). Warp-cooperative.
__device__ __forceinline__ void pack_read_to_store(const uint8_t* __restrict__ src, u32 len, uint8_t* __restrict__ dst,
                                                   int lane) {
    const u32 nb = (len + 3) / 4;
    for (u32 q = lane; q < nb; q += 32) {
        u32 b = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 idx = 4 * q + i;
            if (idx < len) {
                bool ok;
                b |= code_of(__ldg(src + idx), ok) << (2 * i);
            }
        }
        dst[nb - 1 - q] = (uint8_t)b;
    }
}

// An upsert that ran out of probe budget: park the record; only if even the spill area is full is the job lost.
template <int KW>
__device__ __forceinline__ void spill_record(Counters* ctr, const u64 (&key)[KW], u32 mask) {
    const u64 idx = atomicAdd(&ctr->spill_count, 1ull);
    if (idx >= ctr->spill_cap) { atomicAdd(&ctr->table_overflow, 1ull); return; }
#pragma unroll
    for (int i = 0; i < KW; ++i) ctr->spill_keys[idx * KW + i] = key[i];
    ctr->spill_meta[idx] = (unsigned short)mask;
}

// K1 (+K2 when ROUTE == false).
// Restates ReadsKeyValueParserFactory.SplitReads (:150-196): for every position p of every split mate the
// forward and reverse-complement k-mers, dir = fwd <= rc ? FORWARD : REVERSE, key = the smaller, one
// tuple (key, Node{coverage 1, <=2 edges, read head on p == 0}). The tuple is never materialised: it is
// folded straight into the table (ROUTE == false) or appended to its owner GPU's bucket (ROUTE == true).
// One warp per input line; lanes 1..30 own consecutive positions, lanes 0 and 31 are halo lanes that
// only compute the direction of the neighbouring position.
enum ExtractMode { EX_UPSERT = 0, EX_ROUTE = 1, EX_FLAT = 2 };

template <int KW, int MODE>
__global__ void __launch_bounds__(EX_THREADS, GX_EX_MIN_BLOCKS) extract_kernel(ExtractArgs a) {
    constexpr bool ROUTE = MODE == EX_ROUTE;
    constexpr bool FLAT = MODE == EX_FLAT;
    // groups of 30 positions handled per batch: routing batches a whole short read so that one reservation per
    // (warp, destination) covers up to 120 records; the upsert path batches only in the prefetch variant
    constexpr int NB = ROUTE ? 4 : (MODE == EX_UPSERT ? EX_BATCH : 1);
    constexpr bool STASH = NB > 1;
    __shared__ u64 sq[EX_WARPS][WIN_WORDS];
    // per-lane private stash of a batch's keys, masks and destinations (lane-major: conflict-free)
    __shared__ u64 stash_k[STASH ? EX_WARPS : 1][NB][KW][32];
    __shared__ unsigned short stash_m[STASH ? EX_WARPS : 1][NB][32];
    __shared__ unsigned short stash_o[(STASH && ROUTE) ? EX_WARPS : 1][NB][32];
    __shared__ unsigned short route_cnt[ROUTE ? EX_WARPS : 1][ROUTE ? NB * ROUTE_MAXG : 1];  // records per (row, destination)
    __shared__ u64 route_base[ROUTE ? EX_WARPS : 1][ROUTE ? ROUTE_MAXG : 1];               // reserved start per destination
    __shared__ u32 bucket_hist[FLAT ? MAX_BUCKETS : 1];  // this CTA's records per table region (EX_FLAT)
    if constexpr (FLAT) {
        for (u32 i = threadIdx.x; i < a.n_buckets; i += EX_THREADS) bucket_hist[i] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lane_lt = (1u << lane) - 1u;
    u64* W = sq[warp];
    uint8_t* Wb = reinterpret_cast<uint8_t*>(W);
    Head<KW>* heads = reinterpret_cast<Head<KW>*>(a.heads);
    const int k = a.k;
    const uint8_t* text_lo = a.text;
    const uint8_t* text_hi = a.text + a.n_text;
    u32 new_slots = 0;

    for (u64 line = (u64)blockIdx.x * EX_WARPS + warp; line < a.n_lines; line += (u64)gridDim.x * EX_WARPS) {
        const LineDesc d = a.desc[line];
        if ((d.flags & 3u) == 0) continue;
#pragma unroll 1
        for (int mate = 0; mate < 2; ++mate) {
            const u32 len = d.len[mate];
            if (len == 0) continue;
            const uint8_t* rd = a.text + d.off[mate];
            pack_read_to_store(rd, len, a.store + d.store[mate], lane);
            if (!(d.flags & (1u << mate))) continue;
            const u32 npos = len - (u32)k + 1u;
            for (u32 pa = 0; pa < npos; pa += WIN_POSITIONS) {
                const u32 pb = min(pa + (u32)WIN_POSITIONS, npos);
                const u32 lo = pa > 0 ? pa - 1 : 0;
                const u32 hi = pb < npos ? pb + k : len;
                const uint8_t* src = rd + lo;
                const u32 m = (u32)((uintptr_t)src & 3u);
                const uint8_t* aligned = src - m;
                const u32 nwords = (m + (hi - lo) + 3u) / 4u;
                __syncwarp();
                for (u32 j = lane; j < nwords; j += 32) Wb[j] = (uint8_t)load_quad(aligned + 4 * j, text_lo, text_hi);
                __syncwarp();
                for (u32 g0 = pa; g0 < pb; g0 += 30 * NB) {
                    u32 acts = 0;  // bit j: this lane stashed a record in batch row j
                    // ---- phase A: canonical key, edge bits, read head; consume or stash
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const u32 g = g0 + 30 * j;
                        if (g >= pb) break;  // warp-uniform
                        const long long p = (long long)g - 1 + lane;
                        const bool comp = p >= 0 && p < (long long)npos && p <= (long long)pb;
                        u64 f[KW], rc[KW];
                        bool rev = false;
                        if (comp) {
                            window_kmer<KW>(W, (u32)(p - lo) + m, k, f);
                            revcomp_key<KW>(f, k, rc);
                            rev = !key_le<KW>(f, rc);
                        }
                        const u32 dirs = __ballot_sync(0xffffffffu, comp && rev);
                        const bool active = lane >= 1 && lane <= 30 && p < (long long)pb;
                        if (!active) continue;
                        u32 mask = 0;
                        if (p + 1 < (long long)npos)
                            mask |= edge_bit_next(rev, (dirs >> (lane + 1)) & 1u, window_letter(W, (u32)(p + k - lo) + m));
                        if (p > 0)
                            mask |= edge_bit_prev(rev, (dirs >> (lane - 1)) & 1u, window_letter(W, (u32)(p - 1 - lo) + m));
                        u64 key[KW];
#pragma unroll
                        for (int i = 0; i < KW; ++i) key[i] = rev ? rc[i] : f[i];
                        if (p == 0) {
                            Head<KW>& h = heads[d.head_idx[mate]];
#pragma unroll
                            for (int i = 0; i < KW; ++i) h.key[i] = key[i];
                            // offset 0 unflipped, K-1 flipped (:165-170); library always 0 (:98-106)
                            h.uuid = (rev ? ((u64)(k - 1) << 40) : 0ull) | ((u64)mate << 35) | d.read_id;
                            h.this_off = d.store[mate];
                            h.mate_off = d.store[1 - mate];
                            h.this_len = len;
                            h.mate_len = d.len[1 - mate];
                            h.flipped = rev ? 1u : 0u;
                            h.valid = 1u;
                        }
                        if constexpr (FLAT) {
                            // L2-blocked build, pass 1: the occurrence goes to its flat slot (no atomics: the parser
                            // reserved [occ_base, occ_base + positions) for this line) and its table region is counted
                            const u64 idx = d.occ_base + (mate && (d.flags & 1u) ? (u64)(d.len[0] - (u32)k + 1u) : 0ull) + (u64)p;
#pragma unroll
                            for (int i = 0; i < KW; ++i) a.flat_keys[idx * KW + i] = key[i];
                            a.flat_meta[idx] = (unsigned short)mask;
                            atomicAdd(&bucket_hist[bucket_of(hash_key<KW>(key), a.n_buckets)], 1u);
                        } else {
                            bool direct = !STASH;
                            if constexpr (ROUTE) {
                                // multi-GPU: own keys go straight into the table, the others wait for phase B
                                const u32 owner = owner_of(hash_key<KW>(key), a.n_ranks);
                                direct = owner == a.rank;
                                if (!direct) stash_o[warp][j][lane] = (unsigned short)owner;
                            }
                            if (direct) {
                                bool is_new;
                                if (table_upsert<KW>(a.table, a.capacity, key, 1ull, mask, is_new) == a.capacity)
                                    spill_record<KW>(a.ctr, key, mask);
                                new_slots += is_new ? 1u : 0u;
                            } else {
#pragma unroll
                                for (int i = 0; i < KW; ++i) stash_k[warp][j][i][lane] = key[i];
                                stash_m[warp][j][lane] = (unsigned short)mask;
                                if constexpr (!ROUTE)
                                    prefetch_l2(a.table + slot_of(hash_key<KW>(key), a.capacity) * SlotTraits<KW>::WORDS);
                                acts |= 1u << j;
                            }
                        }
                    }
                    // ---- phase B
                    if constexpr (STASH && !ROUTE) {  // prefetch variant: the upserts, now (mostly) L2 hits
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            if (!((acts >> j) & 1u)) continue;
                            u64 key[KW];
#pragma unroll
                            for (int i = 0; i < KW; ++i) key[i] = stash_k[warp][j][i][lane];
                            bool is_new;
                            if (table_upsert<KW>(a.table, a.capacity, key, 1ull, stash_m[warp][j][lane], is_new) == a.capacity)
                                spill_record<KW>(a.ctr, key, stash_m[warp][j][lane]);
                            new_slots += is_new ? 1u : 0u;
                        }
                    }
                    if constexpr (ROUTE) {
                        // Append the batch to the owners' send buckets. One reservation per (warp, destination) covers the
                        // whole batch, and the reservations of ALL destinations are issued together (lane d reserves for
                        // destination d), so the warp waits for one atomic round trip per batch, not one per destination.
                        if (__any_sync(0xffffffffu, acts != 0)) {
                            if (a.n_ranks <= ROUTE_MAXG) {
                                unsigned short* C = route_cnt[warp];
                                for (u32 i = lane; i < NB * ROUTE_MAXG; i += 32) C[i] = 0;
                                __syncwarp();
                                u32 rank_in_row[NB];
#pragma unroll
                                for (int j = 0; j < NB; ++j) {
                                    const bool has = (acts >> j) & 1u;
                                    const u32 rowmask = __ballot_sync(0xffffffffu, has);
                                    rank_in_row[j] = 0;
                                    if (has) {
                                        const u32 dst = stash_o[warp][j][lane];
                                        const u32 peers = __match_any_sync(rowmask, dst);
                                        rank_in_row[j] = __popc(peers & lane_lt);
                                        if (lane == __ffs(peers) - 1) C[j * ROUTE_MAXG + dst] = (unsigned short)__popc(peers);
                                    }
                                }
                                __syncwarp();
                                for (u32 dst = lane; dst < a.n_ranks; dst += 32) {
                                    u32 tot = 0;
#pragma unroll
                                    for (int j = 0; j < NB; ++j) tot += C[j * ROUTE_MAXG + dst];
                                    route_base[warp][dst] = tot ? atomicAdd(a.route_count + dst, (u64)tot) : 0ull;
                                }
                                __syncwarp();
#pragma unroll
                                for (int j = 0; j < NB; ++j) {
                                    if (!((acts >> j) & 1u)) continue;
                                    const u32 dst = stash_o[warp][j][lane];
                                    u64 idx = route_base[warp][dst] + rank_in_row[j];
#pragma unroll
                                    for (int jj = 0; jj < NB; ++jj)
                                        if (jj < j) idx += C[jj * ROUTE_MAXG + dst];
                                    u64* kd = a.route_keys[dst] + idx * KW;
#pragma unroll
                                    for (int i = 0; i < KW; ++i) kd[i] = stash_k[warp][j][i][lane];
                                    a.route_meta[dst][idx] = stash_m[warp][j][lane];
                                }
                                __syncwarp();
                            } else {
                                for (u32 dest = 0; dest < a.n_ranks; ++dest) {  // many ranks: one destination at a time
                                    if (dest == a.rank) continue;
                                    u32 mine[NB];
                                    u32 total = 0;
#pragma unroll
                                    for (int j = 0; j < NB; ++j) {
                                        mine[j] = __ballot_sync(0xffffffffu, ((acts >> j) & 1u) && stash_o[warp][j][lane] == dest);
                                        total += __popc(mine[j]);
                                    }
                                    if (total == 0) continue;
                                    u64 base = 0;
                                    if (lane == 0) base = atomicAdd(a.route_count + dest, (u64)total);
                                    base = __shfl_sync(0xffffffffu, base, 0);
                                    u64* kbase = a.route_keys[dest];
                                    unsigned short* mbase = a.route_meta[dest];
#pragma unroll
                                    for (int j = 0; j < NB; ++j) {
                                        if ((mine[j] >> lane) & 1u) {
                                            const u64 idx = base + __popc(mine[j] & lane_lt);
#pragma unroll
                                            for (int i = 0; i < KW; ++i) kbase[idx * KW + i] = stash_k[warp][j][i][lane];
                                            mbase[idx] = stash_m[warp][j][lane];
                                        }
                                        base += __popc(mine[j]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) new_slots += __shfl_xor_sync(0xffffffffu, new_slots, dlt);
    if (lane == 0 && new_slots) atomicAdd(&a.ctr->distinct, (u64)new_slots);
    if constexpr (FLAT) {
        __syncthreads();
        for (u32 i = threadIdx.x; i < a.n_buckets; i += EX_THREADS)
            if (bucket_hist[i]) atomicAdd(a.bucket_count + (size_t)i * BUCKET_PAD, (u64)bucket_hist[i]);
    }
}

// K2x: upsert pre-extracted (key, mask, count) records (received from other GPUs, or partial aggregates).
template <int KW>
__global__ void __launch_bounds__(256) insert_records_kernel(const u64* __restrict__ keys,
                                                             const unsigned short* __restrict__ meta,
                                                             const u32* __restrict__ counts, u64 n, u64* table,
                                                             u64 capacity, Counters* ctr) {
    u32 new_slots = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = keys[i * KW + j];
        bool is_new;
        if (table_upsert<KW>(table, capacity, key, counts ? (u64)counts[i] : 1ull, meta[i], is_new) == capacity)
            spill_record<KW>(ctr, key, meta[i]);
        new_slots += is_new ? 1u : 0u;
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) new_slots += __shfl_xor_sync(0xffffffffu, new_slots, dlt);
    if ((threadIdx.x & 31) == 0 && new_slots) atomicAdd(&ctr->distinct, (u64)new_slots);
}

// L2-blocked build: pull the next table region into L2 with sequential line prefetches while the current region is being
// upserted, so that the random first touches of a region are L2 hits too.
static __global__ void __launch_bounds__(256) prefetch_region_kernel(const uint8_t* __restrict__ base, u64 bytes) {
    for (u64 off = ((u64)blockIdx.x * 256 + threadIdx.x) * 128; off < bytes; off += (u64)gridDim.x * 256 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
}

// L2-blocked build, pass 2: scatter the flat (key, mask) records into per-region segments whose exact
// offsets come from the histogram taken in pass 1 (bucket_cursor starts at the segment offsets).
static constexpr int PT_THREADS = 256;
static constexpr int PT_ITEMS = 16;
template <int KW>
__global__ void __launch_bounds__(PT_THREADS) partition_flat_kernel(const u64* __restrict__ flat_keys,
                                                                    const unsigned short* __restrict__ flat_meta, u64 n,
                                                                    u32 n_buckets, u64* __restrict__ bucket_cursor,
                                                                    u64* __restrict__ out_keys,
                                                                    unsigned short* __restrict__ out_meta) {
    __shared__ u32 hist[MAX_BUCKETS];
    __shared__ u64 base[MAX_BUCKETS];
    for (u32 i = threadIdx.x; i < n_buckets; i += PT_THREADS) hist[i] = 0;
    __syncthreads();
    const u64 tile0 = (u64)blockIdx.x * (PT_THREADS * PT_ITEMS);
    u64 key[PT_ITEMS][KW];
    u32 bucket[PT_ITEMS], rank[PT_ITEMS];
#pragma unroll
    for (int it = 0; it < PT_ITEMS; ++it) {
        const u64 i = tile0 + (u64)it * PT_THREADS + threadIdx.x;
        bucket[it] = 0xffffffffu;
        if (i < n) {
#pragma unroll
            for (int j = 0; j < KW; ++j) key[it][j] = __ldcs(flat_keys + i * KW + j);
            bucket[it] = bucket_of(hash_key<KW>(key[it]), n_buckets);
            rank[it] = atomicAdd(&hist[bucket[it]], 1u);
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < n_buckets; i += PT_THREADS)
        base[i] = hist[i] ? atomicAdd(bucket_cursor + (size_t)i * BUCKET_PAD, (u64)hist[i]) : 0ull;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < PT_ITEMS; ++it) {
        if (bucket[it] == 0xffffffffu) continue;
        const u64 i = tile0 + (u64)it * PT_THREADS + threadIdx.x;
        const u64 o = base[bucket[it]] + rank[it];
#pragma unroll
        for (int j = 0; j < KW; ++j) out_keys[o * KW + j] = key[it][j];
        out_meta[o] = __ldcs(flat_meta + i);
    }
}

// ---------------------------------------------------------------------------------------------
template <int KW>
__global__ void __launch_bounds__(256) init_table_kernel(u64* __restrict__ table, u64 capacity) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    const u64 n = capacity * SW;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const int w = (int)(i % SW);
        table[i] = (KW <= 2 && w < KW) ? EMPTY_WORD : 0ull;
    }
}

// grow: re-insert every occupied slot of the old table (count and mask carried over)
template <int KW>
__global__ void __launch_bounds__(256) rehash_kernel(const u64* __restrict__ old_table, u64 old_capacity,
                                                     u64* __restrict__ table, u64 capacity) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < old_capacity; s += (u64)gridDim.x * blockDim.x) {
        const u64* p = old_table + s * SW;
        if (!slot_occupied<KW>(p)) continue;
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = p[j];
        const u64 v = p[KW];
        bool is_new;
        table_upsert<KW>(table, capacity, key, v & COUNT_MASK, (u32)(v >> MASK_SHIFT), is_new);
    }
}

// ---------------------------------------------------------------------------------------------
// K3a: read heads -> owning slot (the reference carries the ReadHeadInfo inside the first k-mer's tuple and
// unions TreeSets per key, AggregateKmerAggregateFactory.java:120-123,141-143)
template <int KW>
__global__ void __launch_bounds__(256) heads_count_kernel(const Head<KW>* __restrict__ heads, u64 n_heads,
                                                          const u64* __restrict__ table, u64 capacity,
                                                          u64* __restrict__ hslot, u32* __restrict__ hcount,
                                                          Counters* ctr) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const Head<KW>& h = heads[i];
    u64 slot = capacity;
    if (h.valid == 1u) {
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = h.key[j];
        slot = table_find<KW>(table, capacity, key);
    }
    hslot[i] = slot;
    if (slot == capacity) { if (h.valid != 2u) atomicAdd(&ctr->heads_missing, 1ull); return; }
    atomicAdd(hcount + slot, 1u);
}

static __global__ void __launch_bounds__(256) heads_scatter_kernel(const u64* __restrict__ hslot, u64 n_heads, u64 capacity,
                                                            const u32* __restrict__ hstart, u32* __restrict__ hfill,
                                                            u32* __restrict__ hperm) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const u64 slot = hslot[i];
    if (slot == capacity) return;
    const u32 pos = hstart[slot] + atomicAdd(hfill + slot, 1u);
    hperm[pos] = (u32)i;
}

// order of ReadHeadInfo.compareTo (ReadHeadInfo.java:247-264): offset, library, mate, readId == numeric order of
// the uuid for the non-negative offsets graph build produces; unflipped set before flipped set; ties (same uuid
// from two input lines) resolved to the earlier line, which the TreeSet keeps.
template <int KW>
__device__ __forceinline__ bool head_less(const Head<KW>* __restrict__ heads, u32 x, u32 y) {
    const Head<KW>& a = heads[x];
    const Head<KW>& b = heads[y];
    if (a.flipped != b.flipped) return a.flipped < b.flipped;
    if (a.uuid != b.uuid) return a.uuid < b.uuid;
    return x < y;
}

template <int KW>
__global__ void __launch_bounds__(256) heads_sort_kernel(const Head<KW>* __restrict__ heads, const u64* __restrict__ hslot,
                                                         u64 n_heads, u64 capacity, const u32* __restrict__ hstart,
                                                         u32* __restrict__ hcount, u32* __restrict__ hperm,
                                                         Counters* ctr) {
    const u64 pos = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 kept = 0;
    if (pos < n_heads) {
        // hperm is dense over [0, n_found); positions past it are unused
        const u32 hi = hperm[pos];
        if (hi != 0xffffffffu) {
            const u64 slot = hslot[hi];
            if (slot != capacity && hstart[slot] == (u32)pos) {  // group leader
                const u32 n = hcount[slot];
                u32* v = hperm + pos;
                if (n > 1) {
                    if (n <= 16) {
                        for (u32 i = 1; i < n; ++i) {
                            const u32 x = v[i];
                            u32 j = i;
                            while (j > 0 && head_less<KW>(heads, x, v[j - 1])) { v[j] = v[j - 1]; --j; }
                            v[j] = x;
                        }
                    } else {  // heapsort
                        auto sift = [&](u32 start, u32 end) {
                            u32 root = start;
                            for (;;) {
                                u32 child = 2 * root + 1;
                                if (child >= end) break;
                                if (child + 1 < end && head_less<KW>(heads, v[child], v[child + 1])) ++child;
                                if (head_less<KW>(heads, v[root], v[child])) {
                                    const u32 t = v[root]; v[root] = v[child]; v[child] = t;
                                    root = child;
                                } else break;
                            }
                        };
                        for (u32 s = n / 2; s-- > 0;) sift(s, n);
                        for (u32 e = n - 1; e > 0; --e) {
                            const u32 t = v[0]; v[0] = v[e]; v[e] = t;
                            sift(0, e);
                        }
                    }
                }
                // TreeSet de-duplication
                kept = n ? 1u : 0u;
                for (u32 i = 1; i < n; ++i) {
                    const Head<KW>& p = heads[v[kept - 1]];
                    const Head<KW>& c = heads[v[i]];
                    if (p.flipped == c.flipped && p.uuid == c.uuid) continue;
                    v[kept++] = v[i];
                }
                hcount[slot] = kept;
            }
        }
    }
    const u64 tot = block_reduce_sum<256>((u64)kept);
    if (threadIdx.x == 0 && tot) atomicAdd(&ctr->read_heads, tot);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU read-head routing: a ReadHeadInfo belongs to the node of the read's first k-mer, so it follows
// that key to its owner GPU together with the packed read and mate sequences it will serialise.
struct HeadRouteArgs {
    void* heads; u64 first, n;              // local heads [first, first+n) created since the last exchange
    const uint8_t* store;                   // local read store
    u32 n_ranks, rank;
    void* const* send_heads;                // [n_ranks] -> Head<KW> send buckets
    uint8_t* const* send_store;             // [n_ranks] -> packed sequence bytes that go with them
    u64* send_head_count; u64* send_store_bytes;  // [n_ranks]
};

template <int KW>
__global__ void __launch_bounds__(256) route_heads_kernel(HeadRouteArgs a) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Head<KW>& h = reinterpret_cast<Head<KW>*>(a.heads)[a.first + i];
    if (h.valid != 1u) return;
    u64 key[KW];
#pragma unroll
    for (int j = 0; j < KW; ++j) key[j] = h.key[j];
    const u32 owner = owner_of(hash_key<KW>(key), a.n_ranks);
    if (owner == a.rank) return;
    const u32 tb = (h.this_len + 3u) / 4u, mb = (h.mate_len + 3u) / 4u;
    const u64 idx = atomicAdd(a.send_head_count + owner, 1ull);
    const u64 off = atomicAdd(a.send_store_bytes + owner, (u64)(tb + mb));
    uint8_t* dst = a.send_store[owner] + off;
    for (u32 j = 0; j < tb; ++j) dst[j] = a.store[h.this_off + j];
    for (u32 j = 0; j < mb; ++j) dst[tb + j] = a.store[h.mate_off + j];
    Head<KW> out = h;
    out.this_off = off;        // relative to the segment this rank sends; the receiver rebases
    out.mate_off = off + tb;
    reinterpret_cast<Head<KW>*>(a.send_heads[owner])[idx] = out;
    h.valid = 2u;              // moved away: ignored by this rank's emit
}

template <int KW>
__global__ void __launch_bounds__(256) rebase_heads_kernel(void* heads, u64 first, u64 n, u64 store_base) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Head<KW>& h = reinterpret_cast<Head<KW>*>(heads)[first + i];
    h.this_off += store_base;
    h.mate_off += store_base;
}

static __global__ void bump_cursors_kernel(Counters* ctr, u64 heads, u64 store_bytes) {
    ctr->head_cursor += heads;
    ctr->store_cursor += store_bytes;
}

// ---------------------------------------------------------------------------------------------
// K3b: sizes and serialisation of `VKmer key | Node` records.
static constexpr int EM_THREADS = 256;
static constexpr int EM_PER_THREAD = 4;
static constexpr int EM_TILE = EM_THREADS * EM_PER_THREAD;  // slots per tile (CTA)
static constexpr int EM_MAX_STAGE_BYTES = 160 * 1024;        // upper bound of the serialise kernel's staging area

struct EmitArgs {
    const u64* table; u64 capacity; int k;
    const void* heads; const u32* hstart; const u32* hcount; const u32* hperm;
    const uint8_t* store;
    u64* tile_bytes; u64* tile_nodes;   // per tile: sums (size pass) then exclusive bases (after the scan)
    uint8_t* out; u64* rec_offsets;
    u64* dense; u64 n_nodes;            // dense node list: (KW key words, value word, slot) per node, slot order
    u32 stage_bytes;                    // dynamic shared memory staging area of the serialise kernel
};

__device__ __forceinline__ u32 head_bytes(u32 this_len, u32 mate_len) {
    // ReadHeadInfo.write (ReadHeadInfo.java:205-212): flags, long, VKmer this, [VKmer mate]
    return 1u + 8u + 4u + (this_len + 3u) / 4u + (mate_len ? 4u + (mate_len + 3u) / 4u : 0u);
}

template <int KW>
__device__ __forceinline__ u32 node_record_bytes(const EmitArgs& a, u64 slot, u64 val, u32& n_unflipped, u32& n_flipped) {
    const u32 nb = (u32)(a.k + 3) / 4u;
    const u32 mask = (u32)(val >> MASK_SHIFT);
    u32 sz = 8u + 4u + nb + 1u + 4u;  // recLen, keyLen, VKmer key, active byte, coverage float
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const u32 c = __popc((mask >> (4 * t)) & 0xfu);
        if (c) sz += 4u + c * (4u + nb);
    }
    n_unflipped = n_flipped = 0;
    const u32 n = a.hcount ? a.hcount[slot] : 0u;
    if (n) {
        const Head<KW>* heads = reinterpret_cast<const Head<KW>*>(a.heads);
        const u32* v = a.hperm + a.hstart[slot];
        for (u32 i = 0; i < n; ++i) {
            const Head<KW>& h = heads[v[i]];
            sz += head_bytes(h.this_len, h.mate_len);
            if (h.flipped) ++n_flipped; else ++n_unflipped;
        }
        if (n_unflipped) sz += 5u;  // boolean wholeBody + int size (ExternalableTreeSet.java:236-253)
        if (n_flipped) sz += 5u;
    }
    return sz;
}

// Tile = EM_TILE consecutive slots per CTA; thread t owns slots 4t..4t+3 of the tile (contiguous 16-byte loads).
template <int KW>
__global__ void __launch_bounds__(EM_THREADS) emit_size_kernel(EmitArgs a) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    const u64 slot0 = (u64)blockIdx.x * EM_TILE + (u64)threadIdx.x * EM_PER_THREAD;
    u64 sz = 0, occ = 0;
#pragma unroll
    for (int i = 0; i < EM_PER_THREAD; ++i) {
        const u64 slot = slot0 + i;
        if (slot < a.capacity) {
            const u64* s = a.table + slot * SW;
            if (slot_occupied<KW>(s)) {
                u32 nu, nf;
                sz += node_record_bytes<KW>(a, slot, s[KW], nu, nf);
                occ += 1;
            }
        }
    }
    const u64 tb = block_reduce_sum<EM_THREADS>(sz);
    const u64 tn = block_reduce_sum<EM_THREADS>(occ);
    if (threadIdx.x == 0) { a.tile_bytes[blockIdx.x] = tb; a.tile_nodes[blockIdx.x] = tn; }
}

// Byte sink that assembles the stream in a 32-bit register and stores whole aligned words; only the first
// and last (partial) words of a record, which it shares with its neighbours, go out as byte stores.
struct WordWriter {
    uint8_t* base;  // 4-byte aligned origin (shared-memory stage or the global record buffer)
    u32 pos;        // byte offset from base of the next byte
    u32 acc;        // bytes of the current word gathered so far (first stream byte in the low lane)
    u32 first;      // != 0 only while in the record's first word: index of our first byte inside it

    __device__ __forceinline__ void init(uint8_t* dst) {
        const u32 mis = (u32)((uintptr_t)dst & 3u);
        base = dst - mis;
        pos = mis;
        acc = 0;
        first = mis;
    }
    __device__ __forceinline__ void flush_word(u32 end) {  // the word [end-4, end) is complete
        uint8_t* w = base + end - 4;
        if (first) {
            for (u32 i = first; i < 4; ++i) w[i] = (uint8_t)(acc >> (8 * i));
            first = 0;
        } else {
            *reinterpret_cast<u32*>(w) = acc;
        }
        acc = 0;
    }
    __device__ __forceinline__ void put8(u32 v) {
        acc |= (v & 0xffu) << (8u * (pos & 3u));
        ++pos;
        if ((pos & 3u) == 0) flush_word(pos);
    }
    __device__ __forceinline__ void put32be(u32 v) {
        const u32 le = __byte_perm(v, 0, 0x0123);  // byte-swapped: first stream byte in the low lane
        const u32 sh = 8u * (pos & 3u);
        acc |= le << sh;
        const u32 keep = sh ? (le >> (32u - sh)) : 0u;
        pos += 4;
        flush_word(pos & ~3u);
        acc = keep;
    }
    __device__ __forceinline__ void put64be(u64 v) { put32be((u32)(v >> 32)); put32be((u32)v); }
    __device__ __forceinline__ void finish() {  // bytes of a last, incomplete word
        const u32 n = pos & 3u;
        uint8_t* w = base + (pos & ~3u);
        for (u32 i = first; i < n; ++i) w[i] = (uint8_t)(acc >> (8 * i));
    }
};

// big-endian bytes of the k-letter value = the reference's Kmer byte array (Kmer.java:225-242)
template <int KW>
__device__ __forceinline__ void put_kmer_bytes(WordWriter& w, const u64 (&x)[KW], u32 nb) {
    // most significant byte first: the (nb & 3) bytes of the partial top 32-bit chunk, then whole chunks.
    // Fully unrolled with predicates so that x[] stays in registers.
    const u32 full = nb >> 2, part = nb & 3u;
    u32 top = 0;
#pragma unroll
    for (int c = 0; c < 2 * KW; ++c)
        if ((u32)c == full) top = (u32)(x[c >> 1] >> (32 * (c & 1)));
    for (u32 i = part; i-- > 0;) w.put8(top >> (8 * i));
#pragma unroll
    for (int c = 2 * KW - 1; c >= 0; --c)
        if ((u32)c < full) w.put32be((u32)(x[c >> 1] >> (32 * (c & 1))));
}

// Node.write (Node.java:408-427) + getActiveFields (:466-487) behind the SequenceFile record framing
// (recordLength, keyLength, VKmer.write VKmer.java:389-391).
template <int KW>
__device__ void serialise_node(const EmitArgs& a, u64 slot, const u64 (&key)[KW], u64 val, u32 rec_bytes, u32 n_unflipped,
                               u32 n_flipped, uint8_t* dst) {
    const u32 nb = (u32)(a.k + 3) / 4u;
    const u32 mask = (u32)(val >> MASK_SHIFT);
    const u64 count = val & COUNT_MASK;
    WordWriter w;
    w.init(dst);
    w.put32be(rec_bytes - 8u);
    w.put32be(4u + nb);
    w.put32be((u32)a.k);
    put_kmer_bytes<KW>(w, key, nb);
    u32 active = 0x80u;  // AVERAGE_COVERAGE always present
#pragma unroll
    for (int t = 0; t < 4; ++t)
        if ((mask >> (4 * t)) & 0xfu) active |= 1u << t;
    if (n_unflipped) active |= 1u << 4;
    if (n_flipped) active |= 1u << 5;
    w.put8(active);
    // all 16 possible neighbours are one-letter shifts of X or of rc(X) (gx_internal.cuh header):
    //   FF b: X[1:]+b   FR b: rc(X[1:]+b) = (3-b)+rc(X)[:-1]   RF b: rc(b+X[:-1]) = rc(X)[1:]+(3-b)   RR b: b+X[:-1]
    u64 rcx[KW];
    revcomp_key<KW>(key, a.k, rcx);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const u32 bits = (mask >> (4 * t)) & 0xfu;
        if (!bits) continue;
        w.put32be((u32)__popc(bits));
        // walk the set bits (not all four bases): lanes of a warp stay converged on "my next edge of this type"
#pragma unroll 1
        for (u32 rest = bits; rest; rest &= rest - 1u) {
            const u32 b = (u32)__ffs(rest) - 1u;
            u64 nk[KW];
            if (t == 0) key_append<KW>(key, a.k, b, nk);
            else if (t == 1) key_prepend<KW>(rcx, a.k, 3u - b, nk);
            else if (t == 2) key_append<KW>(rcx, a.k, 3u - b, nk);
            else key_prepend<KW>(key, a.k, b, nk);
            w.put32be((u32)a.k);
            put_kmer_bytes<KW>(w, nk, nb);
        }
    }
    if (n_unflipped | n_flipped) {
        const Head<KW>* heads = reinterpret_cast<const Head<KW>*>(a.heads);
        const u32* v = a.hperm + a.hstart[slot];
        u32 i = 0;
        for (int set = 0; set < 2; ++set) {
            const u32 n = set ? n_flipped : n_unflipped;
            if (!n) continue;
            w.put8(1);  // wholeBodyInStream
            w.put32be(n);
            for (u32 e = 0; e < n; ++e, ++i) {
                const Head<KW>& h = heads[v[i]];
                w.put8(h.mate_len ? 1 : 0);
                w.put64be(h.uuid);
                w.put32be(h.this_len);
                const u32 tb = (h.this_len + 3u) / 4u;
                for (u32 j = 0; j < tb; ++j) w.put8(a.store[h.this_off + j]);
                if (h.mate_len) {
                    w.put32be(h.mate_len);
                    const u32 mb = (h.mate_len + 3u) / 4u;
                    for (u32 j = 0; j < mb; ++j) w.put8(a.store[h.mate_off + j]);
                }
            }
        }
    }
    w.put32be(__float_as_uint((float)count));  // coverage = float sum of 1.0s (exact to 2^24)
    w.finish();
}

// Pass 2 (after the tile sums are scanned): compact the occupied slots into a dense node list in slot order and
// give every node its byte offset in the record stream. Streaming, every lane busy.
template <int KW>
__global__ void __launch_bounds__(EM_THREADS) emit_compact_kernel(EmitArgs a) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    constexpr int DW = KW + 2;
    const u64 tile0 = (u64)blockIdx.x * EM_TILE + (u64)threadIdx.x * EM_PER_THREAD;
    u32 sz[EM_PER_THREAD];
    u32 my_bytes = 0, my_nodes = 0;
#pragma unroll
    for (int i = 0; i < EM_PER_THREAD; ++i) {
        sz[i] = 0;
        const u64 slot = tile0 + i;
        if (slot < a.capacity) {
            const u64* s = a.table + slot * SW;
            u32 nu, nf;
            if (slot_occupied<KW>(s)) sz[i] = node_record_bytes<KW>(a, slot, s[KW], nu, nf);
        }
        my_bytes += sz[i];
        my_nodes += sz[i] ? 1u : 0u;
    }
    u64 tile_total, tile_nodes;
    u64 ex = a.tile_bytes[blockIdx.x] + block_scan_excl<EM_THREADS>((u64)my_bytes, &tile_total);
    u64 nex = a.tile_nodes[blockIdx.x] + block_scan_excl<EM_THREADS>((u64)my_nodes, &tile_nodes);
#pragma unroll
    for (int i = 0; i < EM_PER_THREAD; ++i) {
        if (!sz[i]) continue;
        const u64 slot = tile0 + i;
        const u64* s = a.table + slot * SW;
        u64* d = a.dense + nex * DW;
#pragma unroll
        for (int j = 0; j <= KW; ++j) d[j] = s[j];  // key words and the value word
        d[KW + 1] = slot;
        a.rec_offsets[nex] = ex;
        ex += sz[i];
        ++nex;
    }
}

// Pass 3: one thread per node of the dense list; a CTA's EM_THREADS consecutive nodes cover one contiguous byte
// range of the stream, staged in shared memory and copied out with aligned 16-byte stores.
template <int KW>
__global__ void __launch_bounds__(EM_THREADS) emit_serialise_kernel(EmitArgs a) {
    constexpr int DW = KW + 2;
    extern __shared__ __align__(16) uint8_t stage[];
    const u64 n0 = (u64)blockIdx.x * EM_THREADS;
    const u64 n1 = min(n0 + (u64)EM_THREADS, a.n_nodes);
    const u64 gbase = a.rec_offsets[n0];
    const u64 tile_total = a.rec_offsets[n1] - gbase;  // rec_offsets[n_nodes] = total bytes
    const u32 skew = (u32)(((uintptr_t)(a.out + gbase)) & 15u);
    const bool staged = tile_total + skew <= (u64)a.stage_bytes;
    const u64 n = n0 + threadIdx.x;
    if (n < n1) {
        const u64* d = a.dense + n * DW;
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = d[j];
        const u64 val = d[KW], slot = d[KW + 1];
        const u64 off = a.rec_offsets[n];
        const u32 sz = (u32)(a.rec_offsets[n + 1] - off);
        u32 nu = 0, nf = 0;
        if (a.hcount && a.hcount[slot]) (void)node_record_bytes<KW>(a, slot, val, nu, nf);
        uint8_t* dst = staged ? (stage + skew + (off - gbase)) : (a.out + off);
        serialise_node<KW>(a, slot, key, val, sz, nu, nf, dst);
    }
    if (!staged) return;
    __syncthreads();
    // coalesced copy-out: stage[skew .. skew+tile_total) -> out[gbase ..), 16-byte body, byte edges
    uint8_t* g0 = a.out + gbase;
    const u64 head = min((u64)((16u - skew) & 15u), tile_total);
    const u64 body = (tile_total - head) / 16u;
    const u64 tail = tile_total - head - body * 16u;
    if (threadIdx.x < head) g0[threadIdx.x] = stage[skew + threadIdx.x];
    const uint4* sv = reinterpret_cast<const uint4*>(stage + skew + head);
    uint4* gv = reinterpret_cast<uint4*>(g0 + head);
    for (u64 i = threadIdx.x; i < body; i += EM_THREADS) gv[i] = sv[i];
    if (threadIdx.x < tail) g0[head + body * 16u + threadIdx.x] = stage[skew + head + body * 16u + threadIdx.x];
}

// Fused graph statistics over the dense node list (GraphStatistics.java:78-131; Node.java:820-848).
struct GraphStatsDev {
    u64 nodes, degree_total, degree_max, degree_bins[17], coverage_total, coverage_max, coverage_bins[257];
    u64 unflipped, flipped, self_edges[4], path_nodes, tips_forward, tips_reverse, tips_both, tips_one;
};

template <int KW>
__global__ void __launch_bounds__(256) graph_stats_kernel(EmitArgs a, GraphStatsDev* __restrict__ out) {
    constexpr int DW = KW + 2;
    __shared__ u32 s_deg[17];
    __shared__ u32 s_cov[257];
    for (int i = threadIdx.x; i < 17; i += 256) s_deg[i] = 0;
    for (int i = threadIdx.x; i < 257; i += 256) s_cov[i] = 0;
    __syncthreads();
    u64 deg_tot = 0, cov_tot = 0, unfl = 0, fl = 0, path = 0, tf = 0, tr = 0, tb = 0, to = 0, nodes = 0;
    u64 deg_max = 0, cov_max = 0, self[4] = {0, 0, 0, 0};
    for (u64 n = (u64)blockIdx.x * 256 + threadIdx.x; n < a.n_nodes; n += (u64)gridDim.x * 256) {
        const u64* d = a.dense + n * DW;
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = d[j];
        const u64 val = d[KW], slot = d[KW + 1];
        const u32 mask = (u32)(val >> MASK_SHIFT);
        const u64 cov = val & COUNT_MASK;
        const u32 out_deg = __popc(mask & 0xffu), in_deg = __popc(mask >> 8);
        const u32 deg = in_deg + out_deg;
        ++nodes;
        deg_tot += deg; deg_max = max(deg_max, (u64)deg);
        atomicAdd(&s_deg[deg], 1u);
        cov_tot += cov; cov_max = max(cov_max, cov);
        atomicAdd(&s_cov[cov > 256 ? 256 : (u32)cov], 1u);
        if (in_deg == 1 && out_deg == 1) ++path;
        if (out_deg == 0) ++tf;
        if (in_deg == 0) ++tr;
        if (in_deg == 0 && out_deg == 0) ++tb;
        if ((in_deg == 0) != (out_deg == 0)) ++to;
        u64 rcx[KW];
        revcomp_key<KW>(key, a.k, rcx);
        for (u32 rest = mask; rest; rest &= rest - 1u) {
            const u32 bit = (u32)__ffs(rest) - 1u, t = bit >> 2, b = bit & 3u;
            u64 nk[KW];
            if (t == 0) key_append<KW>(key, a.k, b, nk);
            else if (t == 1) key_prepend<KW>(rcx, a.k, 3u - b, nk);
            else if (t == 2) key_append<KW>(rcx, a.k, 3u - b, nk);
            else key_prepend<KW>(key, a.k, b, nk);
            if (key_eq<KW>(nk, key)) ++self[t];
        }
        if (a.hcount) {
            const u32 nh = a.hcount[slot];
            if (nh) {
                const Head<KW>* heads = reinterpret_cast<const Head<KW>*>(a.heads);
                const u32* v = a.hperm + a.hstart[slot];
                for (u32 i = 0; i < nh; ++i) { if (heads[v[i]].flipped) ++fl; else ++unfl; }
            }
        }
    }
    // block reductions, then one atomic per counter and CTA
    const u64 r_nodes = block_reduce_sum<256>(nodes), r_deg = block_reduce_sum<256>(deg_tot), r_cov = block_reduce_sum<256>(cov_tot);
    const u64 r_unfl = block_reduce_sum<256>(unfl), r_fl = block_reduce_sum<256>(fl), r_path = block_reduce_sum<256>(path);
    const u64 r_tf = block_reduce_sum<256>(tf), r_tr = block_reduce_sum<256>(tr), r_tb = block_reduce_sum<256>(tb), r_to = block_reduce_sum<256>(to);
    u64 r_self[4];
    for (int t = 0; t < 4; ++t) r_self[t] = block_reduce_sum<256>(self[t]);
    atomicMax(&out->degree_max, deg_max);
    atomicMax(&out->coverage_max, cov_max);
    if (threadIdx.x == 0) {
        atomicAdd(&out->nodes, r_nodes); atomicAdd(&out->degree_total, r_deg); atomicAdd(&out->coverage_total, r_cov);
        atomicAdd(&out->unflipped, r_unfl); atomicAdd(&out->flipped, r_fl); atomicAdd(&out->path_nodes, r_path);
        atomicAdd(&out->tips_forward, r_tf); atomicAdd(&out->tips_reverse, r_tr); atomicAdd(&out->tips_both, r_tb);
        atomicAdd(&out->tips_one, r_to);
        for (int t = 0; t < 4; ++t) atomicAdd(&out->self_edges[t], r_self[t]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 17; i += 256) if (s_deg[i]) atomicAdd(&out->degree_bins[i], (u64)s_deg[i]);
    for (int i = threadIdx.x; i < 257; i += 256) if (s_cov[i]) atomicAdd(&out->coverage_bins[i], (u64)s_cov[i]);
}

// R3: KmerPartitionComputerFactory.partition over emitted records (KmerPartitionComputerFactory.java:28-52):
// h = 1; h = 31*h + (signed byte) over the Kmer field bytes; h < 0 -> -(h+1); h % nParts
static __global__ void __launch_bounds__(256) partition_records_kernel(const uint8_t* __restrict__ records,
                                                                const u64* __restrict__ rec_offsets, u64 n_nodes,
                                                                int n_parts, int* __restrict__ parts) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const uint8_t* r = records + rec_offsets[i];
    const u32 key_len = ((u32)r[4] << 24) | ((u32)r[5] << 16) | ((u32)r[6] << 8) | (u32)r[7];
    int h = 1;
    for (u32 j = 4; j < key_len; ++j) h = 31 * h + (int)(signed char)r[8 + j];  // skip the VKmer length header
    if (h < 0) h = -(h + 1);
    parts[i] = h % n_parts;
}

// ---------------------------------------------------------------------------------------------
// Host-callable launch table, one instance per KW (instantiated in gx_kw<N>.cu).
struct EngineOps {
    int kw;
    size_t slot_bytes;
    size_t head_bytes;
    void (*init_table)(u64* table, u64 capacity, cudaStream_t st);
    void (*extract_insert)(const ExtractArgs& a, cudaStream_t st);
    void (*extract_route)(const ExtractArgs& a, cudaStream_t st);
    void (*extract_flat)(const ExtractArgs& a, cudaStream_t st);
    void (*partition_flat)(const u64* flat_keys, const unsigned short* flat_meta, u64 n, u32 n_buckets, u64* bucket_cursor,
                           u64* out_keys, unsigned short* out_meta, cudaStream_t st);
    void (*insert_records)(const u64* keys, const unsigned short* meta, const u32* counts, u64 n, u64* table,
                           u64 capacity, Counters* ctr, cudaStream_t st);
    void (*rehash)(const u64* old_table, u64 old_capacity, u64* table, u64 capacity, cudaStream_t st);
    void (*heads_count)(const void* heads, u64 n_heads, const u64* table, u64 capacity, u64* hslot, u32* hcount,
                        Counters* ctr, cudaStream_t st);
    void (*heads_sort)(const void* heads, const u64* hslot, u64 n_heads, u64 capacity, const u32* hstart, u32* hcount,
                       u32* hperm, Counters* ctr, cudaStream_t st);
    void (*emit_size)(const EmitArgs& a, cudaStream_t st);
    void (*emit_compact)(const EmitArgs& a, cudaStream_t st);
    void (*emit_serialise)(const EmitArgs& a, cudaStream_t st);
    void (*graph_stats)(const EmitArgs& a, GraphStatsDev* out, cudaStream_t st);
    void (*route_heads)(const HeadRouteArgs& a, cudaStream_t st);
    void (*rebase_heads)(void* heads, u64 first, u64 n, u64 store_base, cudaStream_t st);
    int (*prepare)();  // one-time function attributes (dynamic shared memory opt-in)
};

const EngineOps* engine_ops(int kw);
const EngineOps* engine_ops_kw1();
const EngineOps* engine_ops_kw2();
const EngineOps* engine_ops_kw3();
const EngineOps* engine_ops_kw4();

static inline unsigned grid_for(u64 n, unsigned threads, unsigned max_blocks = 148u * 16u) {
    u64 b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (unsigned)b;
}

}  // namespace gx
