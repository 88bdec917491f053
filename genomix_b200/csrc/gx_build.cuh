// gx_build.cuh -- k-mer extraction helpers shared by the build kernels, and the table life-cycle kernels.
//
//   insert_records_kernel<KW>  (key, mask, count) records -> upsert: records that spilled, rehash leftovers
//   init_table / rehash        table life cycle
// The build itself (region-sorted k-mer records, L2-resident upserts) lives in gx_split.cuh.
//
// Reference semantics restated by each kernel are cited at the kernel.
#pragma once
#include "gx_internal.cuh"
#include "gx_parse.cuh"
#include "gx_scan.cuh"
#include "gx_table.cuh"

namespace gx {

static constexpr int WIN_BYTES = 512;                       // packed letters per warp window: 2048 letters
static constexpr int WIN_WORDS = WIN_BYTES / 8 + GX_MAX_KW + 2;

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
    // A 0x41 C 0x43 G 0x47 T 0x54 (either case): code = ((c >> 1) ^ (c >> 2)) & 3; anything else packs as A
    // (GeneCode.java:29-50). The parser already rejected non-ACGT mates that get split; mate sequences that are only
    // stored may hold other letters, so the check stays.
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

// f <- letters [p+1, p+1+k) and rc <- its reverse complement, from those of position p and the letter c at p+k
template <int KW>
__device__ __forceinline__ void roll_kmer(u64 (&f)[KW], u64 (&rc)[KW], int k, u32 c) {
#pragma unroll
    for (int i = 0; i < KW; ++i) f[i] = (f[i] >> 2) | ((i + 1 < KW) ? (f[i + 1] << 62) : 0ull);
    const int pos = 2 * (k - 1);
#pragma unroll
    for (int i = 0; i < KW; ++i)
        if (i == (pos >> 6)) f[i] |= (u64)c << (pos & 63);
#pragma unroll
    for (int i = KW - 1; i >= 0; --i) rc[i] = (rc[i] << 2) | ((i > 0) ? (rc[i - 1] >> 62) : (u64)(3u - c));
    rc[KW - 1] &= top_word_mask(k);
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

// An upsert that ran out of probe budget: park the record with its count; only if even the spill area is full is
// the job lost (the host reports GX_ERR_NOMEM).
template <int KW>
__device__ __forceinline__ void spill_record(Counters* ctr, const u64 (&key)[KW], u32 mask, u32 count = 1u) {
    const u64 idx = atomicAdd(&ctr->spill_count, 1ull);
    if (idx >= ctr->spill_cap) { atomicAdd(&ctr->table_overflow, 1ull); return; }
#pragma unroll
    for (int i = 0; i < KW; ++i) ctr->spill_keys[idx * KW + i] = key[i];
    ctr->spill_meta[idx] = (unsigned short)mask;
    ctr->spill_counts[idx] = count;
}

// The ReadHeadInfo of a read whose first k-mer has canonical key `key` and direction `rev`
// (ReadsKeyValueParserFactory.java:165-170: offset 0 unflipped, K-1 flipped; library always 0, :98-106).
template <int KW>
__device__ __forceinline__ void write_head(Head<KW>* heads, const LineDesc& d, u64 order, int mate, const u64 (&key)[KW], bool rev, int k) {
    Head<KW>& h = heads[d.head_idx[mate]];
    h.order = order;
#pragma unroll
    for (int i = 0; i < KW; ++i) h.key[i] = key[i];
    h.uuid = (rev ? ((u64)(k - 1) << 40) : 0ull) | ((u64)mate << 35) | d.read_id;
    h.this_off = d.store[mate];
    h.mate_off = d.store[1 - mate];
    h.this_len = d.len[mate];
    h.mate_len = d.len[1 - mate];
    h.flipped = rev ? 1u : 0u;
    h.valid = 1u;
}

// Upsert pre-extracted (key, mask, count) records: the spill area, and gx_merge-style partial aggregates.
template <int KW>
__global__ void __launch_bounds__(256) insert_records_kernel(const u64* __restrict__ keys,
                                                             const unsigned short* __restrict__ meta,
                                                             const u32* __restrict__ counts, u64 n, u64* table, u64 capacity,
                                                             u64 hash_mul, Counters* ctr) {
    u32 new_slots = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = keys[i * KW + j];
        const u32 cnt = counts ? counts[i] : 1u;
        bool is_new;
        if (table_upsert<KW>(table, capacity, local_hash(hash_key<KW>(key), hash_mul), key, (u64)cnt, meta[i], is_new) == capacity)
            spill_record<KW>(ctr, key, meta[i], cnt);
        new_slots += is_new ? 1u : 0u;
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) new_slots += __shfl_xor_sync(0xffffffffu, new_slots, dlt);
    if ((threadIdx.x & 31) == 0 && new_slots) atomicAdd(&ctr->distinct, (u64)new_slots);
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

// grow: re-insert every occupied slot of the old table (count and mask carried over) under the new table's mapping
// (`hash_mul`; the old table is only scanned). Home slots are monotone in the hash, so when both tables use the same
// mapping the old table read in slot order lands in the new one in (almost) slot order: the pass streams through both. A record whose probe budget runs out is spilled with its count.
template <int KW>
__global__ void __launch_bounds__(256) rehash_kernel(const u64* __restrict__ old_table, u64 old_capacity,
                                                     u64* __restrict__ table, u64 capacity, u64 hash_mul, Counters* ctr) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < old_capacity; s += (u64)gridDim.x * blockDim.x) {
        const u64* p = old_table + s * SW;
        if (!slot_occupied<KW>(p)) continue;
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = p[j];
        const u64 v = p[KW];
        bool is_new;
        if (table_upsert<KW>(table, capacity, local_hash(hash_key<KW>(key), hash_mul), key, v & COUNT_MASK, (u32)(v >> MASK_SHIFT),
                             is_new) == capacity) {
            // cannot happen while capacity > old occupancy; kept lossless anyway (counts above 2^32-1 are split)
            u64 left = v & COUNT_MASK;
            while (left) {
                const u32 part = left > 0xffffffffull ? 0xffffffffu : (u32)left;
                spill_record<KW>(ctr, key, (u32)(v >> MASK_SHIFT), part);
                left -= part;
            }
            atomicAdd(&ctr->distinct, ~0ull);  // the key left the table: it is counted again when re-inserted
        }
    }
}

}  // namespace gx
