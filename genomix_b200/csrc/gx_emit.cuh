// gx_emit.cuh -- the finish phase: read heads -> nodes, record sizing, Node serialisation, graph statistics.
//
//   K3a  heads_count/scatter/sort      ReadHeadInfo grouping per node (TreeSet order and de-duplication)
//   K3b  emit_size/compact/serialise   `VKmer key | Node` records in SequenceFile record framing
//        graph_stats_kernel, partition_records_kernel, route/rebase_heads (multi-GPU)
//
// Reference semantics restated by each kernel are cited at the kernel.
#pragma once
#include "gx_build.cuh"

namespace gx {

// ---------------------------------------------------------------------------------------------
// K3a: read heads -> nodes (the reference carries the ReadHeadInfo inside the first k-mer's tuple and unions TreeSets
// per key, AggregateKmerAggregateFactory.java:120-123,141-143).
//
// Only ~1 % of a job's k-mer occurrences carry a head, so nothing here is sized by the table: a head finds its node's
// slot, marks the node (HEADS_FLAG in the value word) and joins the node's *group* in a small open-addressing map keyed
// by slot (2 entries per head). Groups are then laid out contiguously (count -> scan -> scatter), put in TreeSet order,
// de-duplicated, and every kept head gets its byte offset inside the node's serialised head sets.
static constexpr u32 HG_SMALL = 32;                // groups up to this size are sorted by one thread
static constexpr u64 HG_EMPTY = ~0ull;

struct HeadGroup {
    u32 start;     // first position of the group in hperm / hoff
    u32 nu, nf;    // heads kept in startReads (unflipped) and endReads (flipped)
    u32 bytes;     // serialised bytes of all kept heads (without the two set headers)
};

__device__ __forceinline__ u32 head_group_find(const u64* __restrict__ ht_key, u32 ht_mask, u64 slot) {
    u32 h = (u32)mix64(slot) & ht_mask;
    while (ht_key[h] != slot) h = (h + 1u) & ht_mask;   // the group exists: HEADS_FLAG is only set after it was created
    return h;
}

template <int KW>
__global__ void __launch_bounds__(256) heads_lookup_kernel(const Head<KW>* __restrict__ heads, u64 n_heads, u64* table, u64 capacity,
                                                           u32 n_ranks, u64* __restrict__ ht_key, u32* __restrict__ ht_count, u32 ht_mask,
                                                           u32* __restrict__ hentry, Counters* ctr) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const Head<KW>& h = heads[i];
    u64 slot = capacity;
    if (h.valid == 1u) {
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = h.key[j];
        slot = table_find<KW>(table, capacity, local_hash(hash_key<KW>(key), n_ranks), key);
    }
    if (slot == capacity) {
        hentry[i] = 0xffffffffu;
        if (h.valid != 2u) atomicAdd(&ctr->heads_missing, 1ull);
        return;
    }
    u64* val = table + slot * SW + KW;
    if (!(*val & HEADS_FLAG)) atomicOr(val, HEADS_FLAG);
    u32 e = (u32)mix64(slot) & ht_mask;
    for (;;) {
        const u64 prev = atomicCAS(ht_key + e, HG_EMPTY, slot);
        if (prev == HG_EMPTY || prev == slot) break;
        e = (e + 1u) & ht_mask;
    }
    atomicAdd(ht_count + e, 1u);
    hentry[i] = e;
}

static __global__ void __launch_bounds__(256) heads_scatter_kernel(const u32* __restrict__ hentry, u64 n_heads,
                                                            const u32* __restrict__ ht_start, u32* __restrict__ ht_fill,
                                                            u32* __restrict__ hperm) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const u32 e = hentry[i];
    if (e == 0xffffffffu) return;
    hperm[ht_start[e] + atomicAdd(ht_fill + e, 1u)] = (u32)i;
}

// order of ReadHeadInfo.compareTo (ReadHeadInfo.java:247-264): offset, library, mate, readId == numeric order of
// the uuid for the non-negative offsets graph build produces; unflipped set before flipped set; ties (same uuid
// from two input lines) resolved to the earlier line, which the TreeSet keeps.
template <int KW>
__device__ __forceinline__ u64 head_sort_key(const Head<KW>& h) { return ((u64)h.flipped << 63) | h.uuid; }   // uuid < 2^48

__device__ __forceinline__ u32 head_bytes(u32 this_len, u32 mate_len) {
    // ReadHeadInfo.write (ReadHeadInfo.java:205-212): flags, long, VKmer this, [VKmer mate]
    return 1u + 8u + 4u + (this_len + 3u) / 4u + (mate_len ? 4u + (mate_len + 3u) / 4u : 0u);
}

// One thread per group: small groups are sorted, de-duplicated and laid out here; large ones go to the big list.
template <int KW>
__global__ void __launch_bounds__(256) heads_group_kernel(const Head<KW>* __restrict__ heads, const u64* __restrict__ ht_key, u32 ht_size,
                                                          const u32* __restrict__ ht_start, const u32* __restrict__ ht_count,
                                                          u32* __restrict__ hperm, u32* __restrict__ hoff, HeadGroup* __restrict__ group,
                                                          u32* __restrict__ big_list, Counters* ctr) {
    const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
    u32 kept = 0;
    if (e < ht_size && ht_key[e] != HG_EMPTY) {
        const u32 n = ht_count[e], start = ht_start[e];
        if (n > HG_SMALL) {
            big_list[atomicAdd(&ctr->big_groups, 1ull)] = e;
        } else {
            u32* v = hperm + start;
            u64 kx[HG_SMALL];   // sort keys, local memory (this kernel touches ~1 % of the job's data)
            for (u32 i = 0; i < n; ++i) kx[i] = head_sort_key<KW>(heads[v[i]]);
            for (u32 i = 1; i < n; ++i) {   // insertion sort by (key, input order)
                const u32 x = v[i];
                const u64 k = kx[i];
                u32 j = i;
                while (j > 0 && (kx[j - 1] > k || (kx[j - 1] == k && heads[v[j - 1]].order > heads[x].order))) {
                    v[j] = v[j - 1]; kx[j] = kx[j - 1]; --j;
                }
                v[j] = x; kx[j] = k;
            }
            u32 nu = 0, nf = 0, bytes = 0;
            for (u32 i = 0; i < n; ++i) {
                if (i && kx[i] == kx[i - 1]) continue;   // TreeSet de-duplication: the earlier line stays
                const Head<KW>& h = heads[v[i]];
                v[kept] = v[i];
                hoff[start + kept] = bytes;
                bytes += head_bytes(h.this_len, h.mate_len);
                if (kx[i] >> 63) ++nf; else ++nu;
                ++kept;
            }
            group[e] = HeadGroup{start, nu, nf, bytes};
        }
    }
    const u64 tot = block_reduce_sum<256>((u64)kept);
    if (threadIdx.x == 0 && tot) atomicAdd(&ctr->read_heads, tot);
}

// Large groups (a repeat that thousands of reads start with; the shape of the reference's EdgeSizePressureTest): one CTA
// per group. Bitonic sort with ascending comparators only (flip + disperse steps), so the virtual +inf padding up to the
// next power of two never moves and the network works in place for any n; then a chunked scan de-duplicates and lays out.
static constexpr int HB_THREADS = 1024;

template <int KW>
__global__ void __launch_bounds__(HB_THREADS) heads_group_big_kernel(const Head<KW>* __restrict__ heads, const u32* __restrict__ ht_start,
                                                                     const u32* __restrict__ ht_count, u32* __restrict__ hperm,
                                                                     u32* __restrict__ hoff, u64* __restrict__ bkey,
                                                                     HeadGroup* __restrict__ group, const u32* __restrict__ big_list,
                                                                     Counters* ctr) {
    __shared__ u64 s_carry[4];   // kept so far, bytes so far, nf so far, last key of the previous chunk
    const u32 n_big = (u32)ctr->big_groups;
    for (u32 b = blockIdx.x; b < n_big; b += gridDim.x) {
        const u32 e = big_list[b];
        const u32 n = ht_count[e], start = ht_start[e];
        u32* v = hperm + start;
        u64* kx = bkey + start;
        for (u32 i = threadIdx.x; i < n; i += HB_THREADS) kx[i] = head_sort_key<KW>(heads[v[i]]);
        __syncthreads();
        auto cmpxchg = [&](u32 i, u32 l) {   // i < l < n: smaller (key, input order) to the lower index
            const u64 ki = kx[i], kl = kx[l];
            const u32 vi = v[i], vl = v[l];
            if (ki > kl || (ki == kl && heads[vi].order > heads[vl].order)) { kx[i] = kl; kx[l] = ki; v[i] = vl; v[l] = vi; }
        };
        u32 np2 = 1;
        while (np2 < n) np2 <<= 1;
        for (u32 k = 2; k <= np2; k <<= 1) {
            for (u32 i = threadIdx.x; i < n; i += HB_THREADS) {   // flip
                const u32 l = i ^ (k - 1u);
                if (l > i && l < n) cmpxchg(i, l);
            }
            __syncthreads();
            for (u32 j = k >> 2; j > 0; j >>= 1) {                // disperse
                for (u32 i = threadIdx.x; i < n; i += HB_THREADS) {
                    const u32 l = i ^ j;
                    if (l > i && l < n) cmpxchg(i, l);
                }
                __syncthreads();
            }
        }
        // de-duplicate + byte offsets, HB_THREADS elements at a time, compacting in place (writes never overtake reads)
        if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; s_carry[2] = 0; s_carry[3] = ~0ull; }
        __syncthreads();
        for (u32 base = 0; base < n; base += HB_THREADS) {
            const u32 i = base + threadIdx.x;
            u64 k = 0, prev = 0;
            u32 hv = 0, hb = 0;
            bool keep = false;
            if (i < n) {
                k = kx[i];
                prev = (threadIdx.x == 0) ? s_carry[3] : kx[i - 1];
                hv = v[i];
                keep = (i == 0) || k != prev;
                if (keep) { const Head<KW>& h = heads[hv]; hb = head_bytes(h.this_len, h.mate_len); }
            }
            u64 tot_k, tot_b, tot_f;
            const u64 ex_k = block_scan_excl<HB_THREADS>(keep ? 1ull : 0ull, &tot_k);
            const u64 ex_b = block_scan_excl<HB_THREADS>((u64)hb, &tot_b);
            const u64 my_f = (keep && (k >> 63)) ? 1ull : 0ull;
            (void)block_scan_excl<HB_THREADS>(my_f, &tot_f);
            const u64 c_k = s_carry[0], c_b = s_carry[1];
            const u64 last = (base + HB_THREADS <= n) ? kx[base + HB_THREADS - 1] : 0ull;
            __syncthreads();   // every read of this chunk (and of the carries) is done
            if (keep) { v[c_k + ex_k] = hv; hoff[start + c_k + ex_k] = (u32)(c_b + ex_b); }
            if (threadIdx.x == 0) { s_carry[0] = c_k + tot_k; s_carry[1] = c_b + tot_b; s_carry[2] += tot_f; s_carry[3] = last; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const u32 kept = (u32)s_carry[0], nf = (u32)s_carry[2];
            group[e] = HeadGroup{start, kept - nf, nf, (u32)s_carry[1]};
            atomicAdd(&ctr->read_heads, (u64)kept);
        }
        __syncthreads();
    }
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
// K3b: `VKmer key | Node` records.
//   emit_scan_kernel   ONE pass over the table: record size per occupied slot, single-pass (decoupled look-back) scan
//                      over tiles, dense node list (key words + value word) and the byte offset of every record.
//   emit_write_kernel  item-parallel serialiser: a record is a header item, one item per edge and one per read head;
//                      every thread writes one item, so lanes do equal work whatever the node degrees are. A CTA's
//                      records cover one contiguous byte range, staged in shared memory, copied out in 16-byte rows.
static constexpr int EM_THREADS = 256;
static constexpr int EM_PER_THREAD = 4;
static constexpr int EM_TILE = EM_THREADS * EM_PER_THREAD;  // slots per tile (CTA) of the scan pass
static constexpr int EW_NODES = 32;                          // nodes per warp tile of the write pass
static constexpr int EM_MAX_STAGE_BYTES = 160 * 1024;        // upper bound of the write kernel's staging area (CTA)

struct EmitArgs {
    const u64* table; u64 capacity; int k;
    // read heads (null / 0 when the job has none)
    const void* heads; const u32* hperm; const u32* hoff; const uint8_t* store;
    const u64* ht_key; u32 ht_mask; const HeadGroup* group;
    // scan pass
    u64* tile_state;        // [n_tiles][2]: {flag << 62 | bytes, nodes}, one 16-byte word per tile (zeroed)
    u32* tile_counter;      // dynamic tile ids (zeroed)
    u64* totals;            // [2] record bytes, nodes
    u64* dense;             // dense node list: (KW key words, value word) per node, slot order
    u32* dense_h;           // group index of a node with read heads (written only for those)
    u64* rec_offsets;       // [n_nodes + 1] byte offset of every record
    // write pass: nodes [n_first, n_last) -> out[rec_offsets[n] - out_base ...]
    u64 n_nodes, n_first, n_last, out_base;
    uint8_t* out;
    u32 stage_bytes;        // staging window per warp of the write kernel (dynamic shared memory, multiple of 16)
    u64* big_tiles;         // tiles of more than EW_MAX_WINDOWS windows, left to emit_write_big_kernel
    u64* big_tile_count;
};

// non-empty edge lists of a mask: bit 4t set when type t has an edge
__device__ __forceinline__ u32 list_bits(u32 mask) { return (mask | (mask >> 1) | (mask >> 2) | (mask >> 3)) & 0x1111u; }

// bytes of a record without read heads: recLen, keyLen, VKmer key, active byte, edge lists, coverage float
__device__ __forceinline__ u32 plain_record_bytes(u32 mask, u32 nb) {
    return 8u + 4u + nb + 1u + 4u + 4u * (u32)__popc(list_bits(mask)) + (u32)__popc(mask) * (4u + nb);
}
// bytes the read heads add: both sets' heads plus, per non-empty set, boolean wholeBody + int size
// (ExternalableTreeSet.java:236-253)
__device__ __forceinline__ u32 heads_record_bytes(const HeadGroup& g) { return g.bytes + (g.nu ? 5u : 0u) + (g.nf ? 5u : 0u); }

__device__ __forceinline__ void st_state(u64* p, u64 x, u64 y) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}

// Scan pass, warp-autonomous (no CTA barriers). A warp takes a ticket for the next tile of ES_TILE slots (tickets are
// handed out in start order, so look-back never waits on a tile that has not started), sums the tile's record bytes and
// nodes, publishes the aggregate, looks back over the tiles before it for its exclusive prefix, and then walks the tile
// a second time (L1/L2 hits) to write the dense node list and the record offsets.
static constexpr int ES_SUB = 4;                 // sub-tiles of 32 lanes x EM_PER_THREAD slots per tile
static constexpr int ES_TILE = 32 * EM_PER_THREAD * ES_SUB;

template <int KW>
struct ScanSlots {
    u64 w[EM_PER_THREAD][KW + 1];
    u32 sz[EM_PER_THREAD], he[EM_PER_THREAD];
    u32 bytes, nodes;
};

// this lane's EM_PER_THREAD slots slot0, slot0 + 32, ... (a warp reads 32 consecutive slots per load): words, record
// sizes, head groups
template <int KW>
__device__ __forceinline__ void scan_load(const EmitArgs& a, u64 slot0, u32 nb, ScanSlots<KW>& r) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    constexpr int DW = KW + 1;
    r.bytes = r.nodes = 0;
#pragma unroll
    for (int i = 0; i < EM_PER_THREAD; ++i) {
        r.sz[i] = 0; r.he[i] = 0xffffffffu;
        const u64 slot = slot0 + 32u * i;
        if (slot < a.capacity) {
            const u64* s = a.table + slot * SW;
            if constexpr (KW == 1) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(s);
                r.w[i][0] = v.x; r.w[i][1] = v.y;
            } else {
#pragma unroll
                for (int j = 0; j < DW; ++j) r.w[i][j] = s[j];
            }
            bool occ;
            if constexpr (KW <= 2) {
                occ = false;
#pragma unroll
                for (int j = 0; j < KW; ++j) occ = occ || r.w[i][j] != EMPTY_WORD;
            } else {
                occ = r.w[i][KW] != 0;
            }
            if (occ) {
                const u64 val = r.w[i][KW];
                r.sz[i] = plain_record_bytes((u32)(val >> MASK_SHIFT), nb);
                if (val & HEADS_FLAG) {
                    r.he[i] = head_group_find(a.ht_key, a.ht_mask, slot);
                    r.sz[i] += heads_record_bytes(a.group[r.he[i]]);
                }
                r.bytes += r.sz[i];
                r.nodes += 1u;
            }
        }
    }
}

template <int KW>
__global__ void __launch_bounds__(EM_THREADS) emit_scan_kernel(EmitArgs a) {
    constexpr int DW = KW + 1;
    constexpr u64 VALUE_BITS = (1ull << 62) - 1;
    const int lane = threadIdx.x & 31;
    const u32 nb = (u32)(a.k + 3) / 4u;
    const u64 n_tiles = (a.capacity + ES_TILE - 1) / ES_TILE;
    for (;;) {
        u32 tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if ((u64)tile >= n_tiles) break;
        const u64 slot_base = (u64)tile * ES_TILE + (u64)lane;
        const u32 lane_lt = (1u << lane) - 1u;
        // ---- pass 1: the tile's aggregate. Only the value words are needed (after the build an occupied slot's value word is
        // non-zero for every key width): all of the lane's loads are in flight together
        u64 tile_bytes = 0, tile_nodes = 0;
        {
            constexpr int SW = SlotTraits<KW>::WORDS;
            constexpr int HALF = ES_SUB * EM_PER_THREAD / 2;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                u64 v[HALF];
#pragma unroll
                for (int q = 0; q < HALF; ++q) {
                    const u64 slot = slot_base + 32u * (h * HALF + q);
                    v[q] = slot < a.capacity ? a.table[slot * SW + KW] : 0ull;
                }
#pragma unroll
                for (int q = 0; q < HALF; ++q) {
                    if (!v[q]) continue;
                    u32 sz = plain_record_bytes((u32)(v[q] >> MASK_SHIFT), nb);
                    if (v[q] & HEADS_FLAG)
                        sz += heads_record_bytes(a.group[head_group_find(a.ht_key, a.ht_mask, slot_base + 32u * (h * HALF + q))]);
                    tile_bytes += sz; tile_nodes += 1;
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            tile_bytes += __shfl_xor_sync(0xffffffffu, tile_bytes, d);
            tile_nodes += __shfl_xor_sync(0xffffffffu, tile_nodes, d);
        }
        // ---- decoupled look-back over the tiles before this one: exclusive prefix of (bytes, nodes)
        u64 eb = 0, en = 0;
        u64* st = a.tile_state + 2 * (u64)tile;
        if (tile == 0) {
            if (lane == 0) st_state(st, (2ull << 62) | tile_bytes, tile_nodes);
        } else {
            if (lane == 0) st_state(st, (1ull << 62) | tile_bytes, tile_nodes);
            long long look = (long long)tile - 1;
            for (;;) {
                const long long idx = look - lane;
                u64 x = 2ull << 62, y = 0;   // before the first tile: an inclusive prefix of zero
                do {
                    if (idx >= 0) ld_relaxed_v2(a.tile_state + 2 * idx, x, y);
                } while (__any_sync(0xffffffffu, (x >> 62) == 0));
                const u32 incl = __ballot_sync(0xffffffffu, (x >> 62) == 2);
                u64 vb = x & VALUE_BITS, vn = y;
                if (incl && lane > __ffs(incl) - 1) { vb = 0; vn = 0; }   // tiles before the nearest inclusive prefix
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    vb += __shfl_xor_sync(0xffffffffu, vb, d);
                    vn += __shfl_xor_sync(0xffffffffu, vn, d);
                }
                eb += vb; en += vn;
                if (incl) break;
                look -= 32;
            }
            if (lane == 0) st_state(st, (2ull << 62) | (eb + tile_bytes), en + tile_nodes);
        }
        if (lane == 0 && (u64)tile + 1 == n_tiles) {   // last tile: totals
            a.totals[0] = eb + tile_bytes;
            a.totals[1] = en + tile_nodes;
            a.rec_offsets[en + tile_nodes] = eb + tile_bytes;
        }
        // ---- pass 2: dense node list and record offsets, slot order (slot = sub-tile base + 32 * i + lane)
#pragma unroll 1
        for (int sub = 0; sub < ES_SUB; ++sub) {
            ScanSlots<KW> r;
            scan_load<KW>(a, slot_base + (u64)sub * 32 * EM_PER_THREAD, nb, r);
#pragma unroll
            for (int i = 0; i < EM_PER_THREAD; ++i) {
                u64 ib = r.sz[i];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u64 t = __shfl_up_sync(0xffffffffu, ib, d);
                    if (lane >= d) ib += t;
                }
                const u32 occ = __ballot_sync(0xffffffffu, r.sz[i] != 0);
                const u64 off = eb + ib - r.sz[i], rank = en + (u32)__popc(occ & lane_lt);
                eb += __shfl_sync(0xffffffffu, ib, 31);
                en += (u32)__popc(occ);
                if (!r.sz[i]) continue;
                u64* d = a.dense + rank * DW;
                if constexpr (KW == 1) {
                    *reinterpret_cast<ulonglong2*>(d) = make_ulonglong2(r.w[i][0], r.w[i][1]);
                } else {
#pragma unroll
                    for (int j = 0; j < DW; ++j) d[j] = r.w[i][j];
                }
                a.rec_offsets[rank] = off;
                if (r.he[i] != 0xffffffffu) a.dense_h[rank] = r.he[i];   // the few nodes with read heads
            }
        }
    }
}

// Byte sink of the write pass. The warp's staging window is zeroed shared memory; an item's bytes are gathered in a
// 32-bit register in stream order and leave as whole 32-bit words: the item's first word and its last, partial one --
// the (at most two) words it shares with its neighbours -- are OR-ed in, all others are plain stores, so items can be
// written in any order by any lane without byte stores. With CLIP, words outside the window [0, win_words) are dropped
// (tiles larger than the window are written window by window).
template <bool CLIP>
struct OrWriter {
    u32* win;        // staging window (word 0 = window start)
    u32 win_words;
    u32 widx;        // window-relative index of the word being gathered (wraps below 0 for words before the window)
    u32 sh;          // bits gathered in acc: 0, 8, 16 or 24
    u32 acc;
    bool first;      // no word has left yet

    // v = byte position of the item's first byte relative to the window start (may be negative: two's complement)
    __device__ __forceinline__ void init(u32* window, u32 window_words, int v) {
        win = window; win_words = window_words;
        widx = (u32)(v >> 2);          // arithmetic shift: floor
        sh = 8u * ((u32)v & 3u);
        acc = 0;
        first = true;
    }
    __device__ __forceinline__ bool inside() const { return !CLIP || widx < win_words; }
    // four stream bytes (first byte in the low lane); sh does not change
    __device__ __forceinline__ void put32le_first(u32 le) {   // the item's first word: shared with the item before
        const u32 out = acc | (le << sh);
        acc = __funnelshift_l(le, 0u, sh);     // le >> (32 - sh); 0 when sh == 0
        if (inside()) atomicOr(win + widx, out);
        first = false;
        ++widx;
    }
    __device__ __forceinline__ void put32le(u32 le) {         // any later word: entirely this item's
        const u32 out = acc | (le << sh);
        acc = __funnelshift_l(le, 0u, sh);
        if (inside()) win[widx] = out;
        ++widx;
    }
    __device__ __forceinline__ void put32be_first(u32 v) { put32le_first(__byte_perm(v, 0, 0x0123)); }
    __device__ __forceinline__ void put32be(u32 v) { put32le(__byte_perm(v, 0, 0x0123)); }
    // general form: 1..4 bytes, first word or not
    __device__ __forceinline__ void put(u32 le, u32 nbytes) {
        const u64 t = (u64)acc | ((u64)le << sh);
        sh += 8u * nbytes;
        if (sh >= 32u) {
            if (inside()) {
                if (first) atomicOr(win + widx, (u32)t); else win[widx] = (u32)t;
            }
            first = false;
            acc = (u32)(t >> 32);
            sh -= 32u;
            ++widx;
        } else {
            acc = (u32)t;
        }
    }
    __device__ __forceinline__ void put8(u32 v) { put(v & 0xffu, 1); }
    __device__ __forceinline__ void put32be_any(u32 v) { put(__byte_perm(v, 0, 0x0123), 4); }
    __device__ __forceinline__ void put64be_any(u64 v) { put32be_any((u32)(v >> 32)); put32be_any((u32)v); }
    __device__ __forceinline__ void finish() {
        if (sh && inside()) atomicOr(win + widx, acc);
    }
};

// big-endian bytes of the k-letter value = the reference's Kmer byte array (Kmer.java:225-242); not the item's first bytes
template <int KW, bool CLIP>
__device__ __forceinline__ void put_kmer_bytes(OrWriter<CLIP>& w, const u64 (&x)[KW], u32 nb) {
    // most significant byte first: the (nb & 3) bytes of the partial top 32-bit chunk, then whole chunks.
    // Fully unrolled with predicates so that x[] stays in registers.
    const u32 full = nb >> 2, part = nb & 3u;
    if (part) {
        u32 top = 0;
#pragma unroll
        for (int c = 0; c < 2 * KW; ++c)
            if ((u32)c == full) top = (u32)(x[c >> 1] >> (32 * (c & 1)));
        w.put(__byte_perm(top, 0, 0x0123) >> (8u * (4u - part)), part);
    }
#pragma unroll
    for (int c = 2 * KW - 1; c >= 0; --c)
        if ((u32)c < full) w.put32be((u32)(x[c >> 1] >> (32 * (c & 1))));
}

// The packed letters of a read in the read store (VKmer byte order) -> the stream, four bytes at a time.
template <bool CLIP>
__device__ __forceinline__ void put_store_bytes(OrWriter<CLIP>& w, const uint8_t* __restrict__ src, u32 n) {
    u32 j = 0;
    const u32 mis = (u32)((uintptr_t)src & 3u);
    if (mis) for (; j < min(n, 4u - mis); ++j) w.put8(src[j]);
    for (; j + 4u <= n; j += 4u) w.put(*reinterpret_cast<const u32*>(src + j), 4);   // stream order = memory order
    for (; j < n; ++j) w.put8(src[j]);
}

// A node's record without its read heads, written by ONE thread at window-relative byte position v (Node.write,
// Node.java:408-427, getActiveFields :466-487, behind the SequenceFile record framing: recordLength, keyLength,
// VKmer.write VKmer.java:389-391):
//   recLen keyLen | k, key bytes | active byte | per non-empty edge list: count, (k, neighbour bytes)* | [heads] | coverage
// All 16 possible neighbours are one-letter shifts of X or of rc(X) (gx_internal.cuh header):
//   FF b: X[1:]+b   FR b: rc(X[1:]+b) = (3-b)+rc(X)[:-1]   RF b: rc(b+X[:-1]) = rc(X)[1:]+(3-b)   RR b: b+X[:-1]
template <int KW, bool CLIP>
__device__ __forceinline__ void write_node(const EmitArgs& a, u32* win, u32 win_words, int v, const u64 (&key)[KW], u64 val,
                                           u32 rec_bytes, u32 heads_active) {
    const u32 nb = (u32)(a.k + 3) / 4u;
    const u32 mask = (u32)(val >> MASK_SHIFT);
    const u32 lists = list_bits(mask);
    const u32 active = 0x80u | heads_active | ((lists * 0x1248u >> 12) & 0xfu);  // AVERAGE_COVERAGE always present; bit t: list t
    OrWriter<CLIP> w;
    w.init(win, win_words, v);
    w.put32be_first(rec_bytes - 8u);
    w.put32be(4u + nb);
    w.put32be((u32)a.k);
    put_kmer_bytes<KW, CLIP>(w, key, nb);
    w.put8(active);
    u64 rcx[KW];
    revcomp_key<KW>(key, a.k, rcx);
    // walk the set bits: lanes of a warp stay converged on "my next edge"; the body is the same straight-line code
    // for every edge type (selects, no branches)
#pragma unroll 1
    for (u32 rest = mask; rest; rest &= rest - 1u) {
        const u32 bit = (u32)__ffs(rest) - 1u;
        const u32 t = bit >> 2, b = bit & 3u;
        const u32 nibble = (mask >> (4u * t)) & 0xfu;
        if ((nibble & ((1u << b) - 1u)) == 0) w.put32be((u32)__popc(nibble));   // first edge of its list: the count
        const bool from_rc = (t == 1u) || (t == 2u);
        const u32 base = from_rc ? 3u - b : b;
        u64 src[KW], app[KW], pre[KW], nk[KW];
#pragma unroll
        for (int i = 0; i < KW; ++i) src[i] = from_rc ? rcx[i] : key[i];
        key_append<KW>(src, a.k, base, app);
        key_prepend<KW>(src, a.k, base, pre);
        const bool use_app = (t == 0u) || (t == 2u);
#pragma unroll
        for (int i = 0; i < KW; ++i) nk[i] = use_app ? app[i] : pre[i];
        w.put32be((u32)a.k);
        put_kmer_bytes<KW, CLIP>(w, nk, nb);
    }
    w.finish();
    // coverage = float sum of 1.0s (exact to 2^24): the record's last four bytes (behind the read heads, if any)
    w.init(win, win_words, v + (int)rec_bytes - 4);
    w.put32be_first(__float_as_uint((float)(val & COUNT_MASK)));
    w.finish();
}

// One ReadHeadInfo of a node's startReads / endReads (ReadHeadInfo.write, ReadHeadInfo.java:205-212), preceded by the
// set header (boolean wholeBody + int size, ExternalableTreeSet.java:236-253) if it is the first of its set.
// v_heads = window-relative byte position of the node's head section (right behind the edge lists).
template <int KW>
__device__ __forceinline__ void write_head_item(const EmitArgs& a, u32* win, u32 win_words, int v_heads, const HeadGroup& g, u32 hi) {
    const u32 p = g.start + hi;
    const Head<KW>& h = reinterpret_cast<const Head<KW>*>(a.heads)[a.hperm[p]];
    const bool second = g.nu && hi >= g.nu;            // in endReads, behind a non-empty startReads
    const bool first = hi == 0 || hi == g.nu;          // first head of its set
    const int v = v_heads + (int)(5u + (second ? 5u : 0u) + a.hoff[p]);
    const int len = (int)head_bytes(h.this_len, h.mate_len);
    if (v + len <= 0 || v - 5 >= (int)(win_words * 4u)) return;   // not in this window
    OrWriter<true> w;
    w.init(win, win_words, v - (first ? 5 : 0));
    if (first) {
        w.put8(1);  // wholeBodyInStream
        w.put32be_any((hi == 0 && g.nu) ? g.nu : g.nf);
    }
    w.put8(h.mate_len ? 1 : 0);
    w.put64be_any(h.uuid);
    w.put32be_any(h.this_len);
    put_store_bytes(w, a.store + h.this_off, (h.this_len + 3u) / 4u);
    if (h.mate_len) {
        w.put32be_any(h.mate_len);
        put_store_bytes(w, a.store + h.mate_off, (h.mate_len + 3u) / 4u);
    }
    w.finish();
}

// Write pass. Every WARP is on its own (no CTA barriers: the warps of an SM sit in different phases, which is what hides
// the load and shared-memory latencies): it owns 32 consecutive nodes = one contiguous byte range of the stream, produced
// in windows of `stage_bytes` of the warp's shared memory (almost always one): zero the window, every lane writes its
// node's record (header, edges, coverage), then the read heads that fall into the window are written one lane per head,
// and the window leaves as aligned 16-byte rows (byte stores at the two ragged ends, which neighbouring warps fill from
// their side). A tile of many windows (a node with a huge read-head set) goes to a list instead, and
// emit_write_big_kernel spreads its windows over all warps of the GPU.
static constexpr int EW_WARPS = EM_THREADS / 32;
static constexpr u32 EW_MAX_WINDOWS = 4;   // tiles with more windows than this are written by emit_write_big_kernel

struct WriteSmem {
    u32 hitem[EW_WARPS][33];   // exclusive prefix of the read heads (of each node of the warp's tile) inside the window
    u32 hlo[EW_WARPS][32];     // first such head of each node
    u32 grp[EW_WARPS][32];
    u32 hpos[EW_WARPS][32];    // tile-relative byte position of each node's head section
};

// One tile of up to 32 nodes, as seen by one lane.
template <int KW>
struct WriteTile {
    u64 key[KW];
    u64 val;
    u32 off, rec_bytes, n_heads, heads_active, grp, hpos;   // this lane's node
    u32 tn, skew;                                           // the tile
    u64 span;                                               // skew + bytes of the tile
    uint8_t* g0;                                            // global address of the tile's first byte
};

template <int KW>
__device__ __forceinline__ void write_tile_load(const EmitArgs& a, u64 tile, int lane, WriteTile<KW>& t) {
    constexpr int DW = KW + 1;
    const u32 nb = (u32)(a.k + 3) / 4u;
    const u64 n0 = a.n_first + tile * 32;
    t.tn = (u32)min((u64)32, a.n_last - n0);
    u64 o0 = 0, o1 = 0;
    t.val = 0; t.n_heads = 0; t.heads_active = 0; t.grp = 0xffffffffu;
    if ((u32)lane < t.tn) {
        const u64* d = a.dense + (n0 + lane) * DW;
        if constexpr (KW == 1) {
            const ulonglong2 x = __ldcs(reinterpret_cast<const ulonglong2*>(d));
            t.key[0] = x.x; t.val = x.y;
        } else {
#pragma unroll
            for (int j = 0; j < KW; ++j) t.key[j] = __ldcs(d + j);
            t.val = __ldcs(d + KW);
        }
        o0 = a.rec_offsets[n0 + lane];
        o1 = a.rec_offsets[n0 + lane + 1];
        if (t.val & HEADS_FLAG) {
            t.grp = a.dense_h[n0 + lane];
            const HeadGroup g = a.group[t.grp];
            t.n_heads = g.nu + g.nf;
            t.heads_active = (g.nu ? 1u << 4 : 0u) | (g.nf ? 1u << 5 : 0u);
        }
    }
    const u64 gbase = __shfl_sync(0xffffffffu, o0, 0);
    const u64 tile_total = __shfl_sync(0xffffffffu, o1, (int)t.tn - 1) - gbase;
    t.off = (u32)(o0 - gbase);
    t.rec_bytes = (u32)(o1 - o0);
    t.g0 = a.out + (gbase - a.out_base);
    t.skew = (u32)(((uintptr_t)t.g0) & 15u);     // stage byte s <-> global byte g0 - skew + s
    t.span = (u64)t.skew + tile_total;
    const u32 mask = (u32)(t.val >> MASK_SHIFT);
    t.hpos = t.off + 13u + nb + (u32)__popc(mask) * (4u + nb) + 4u * (u32)__popc(list_bits(mask));
}

// window [wlo, wlo + win_bytes) of the tile's span: zero, write, copy out
template <int KW>
__device__ __forceinline__ void write_tile_window(const EmitArgs& a, const WriteTile<KW>& t, u64 wlo, uint8_t* stage, WriteSmem& S,
                                                  int lane, int warp) {
    const u32 win_bytes = a.stage_bytes;
    u32* win = reinterpret_cast<u32*>(stage);
    const u32 wb = (u32)min((u64)win_bytes, t.span - wlo);           // bytes of this window that belong to the tile
    const u32 wwords = (wb + 3u) / 4u;
    __syncwarp();   // previous window copied out
    for (u32 i = lane; i < (wwords + 3u) / 4u; i += 32) reinterpret_cast<uint4*>(stage)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    if ((u32)lane < t.tn) {
        const long long v = (long long)t.skew + t.off - (long long)wlo;   // window-relative position of my record
        if (t.span <= (u64)win_bytes) write_node<KW, false>(a, win, wwords, (int)v, t.key, t.val, t.rec_bytes, t.heads_active);
        else if (v < (long long)wb && v + (long long)t.rec_bytes > 0)
            write_node<KW, true>(a, win, wwords, (int)v, t.key, t.val, t.rec_bytes, t.heads_active);
    }
    // Read heads. The common case -- a node with one or two heads in a single-window tile -- is written by the node's own
    // lane right away (its group is in registers; the dependent loads of several such lanes overlap). Everything else
    // (many heads, or a tile of several windows) goes through the lane-per-head path below.
    const bool own_heads = t.n_heads != 0 && t.n_heads <= 2u && t.span <= (u64)win_bytes;
    if (own_heads) {
        const HeadGroup g = a.group[t.grp];
        const int vh = (int)(t.skew + t.hpos);
        for (u32 hi = 0; hi < t.n_heads; ++hi) write_head_item<KW>(a, win, wwords, vh, g, hi);
    }
    if (__any_sync(0xffffffffu, t.n_heads != 0 && !own_heads)) {
        // the heads of my node that touch the window: [h_lo, h_hi). A head's bytes start at
        // vh + 5 + (5 if it is in the second of two sets) + hoff, monotone in the head's position.
        u32 h_lo = 0, h_hi = 0;
        const long long vh = (long long)t.skew + t.hpos - (long long)wlo;
        if (t.n_heads && !own_heads) {
            const HeadGroup g = a.group[t.grp];
            const long long hb = (long long)heads_record_bytes(g);
            if (vh < (long long)wb && vh + hb > 0) {
                if (vh >= 0 && vh + hb <= (long long)wb) {
                    h_hi = t.n_heads;   // the whole head section lies inside
                } else {
                    auto start = [&](u32 hi) { return vh + 5 + ((g.nu && hi >= g.nu) ? 5 : 0) + (long long)a.hoff[g.start + hi]; };
                    u32 lo = 0, hi = t.n_heads;          // first head that starts behind the window's first byte
                    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (start(mid) > 0) hi = mid; else lo = mid + 1; }
                    h_lo = lo ? lo - 1u : 0u;            // the one before it may reach into the window
                    lo = h_lo; hi = t.n_heads;           // first head that starts (set header included) behind the window
                    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (start(mid) - 5 >= (long long)wb) hi = mid; else lo = mid + 1; }
                    h_hi = lo;
                }
            }
        }
        const u32 cnt = h_hi - h_lo;
        u32 incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 x = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += x;
        }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        S.hitem[warp][lane] = incl - cnt;
        if (lane == 31) S.hitem[warp][32] = total;
        S.hlo[warp][lane] = h_lo;
        S.grp[warp][lane] = t.grp;
        S.hpos[warp][lane] = t.hpos;
        __syncwarp();
        for (u32 it = lane; it < total; it += 32) {
            u32 lo = 0, hi = 31u;   // largest node with hitem[node] <= it (nodes without heads in the window repeat the prefix)
            while (lo < hi) {
                const u32 mid = (lo + hi + 1u) >> 1;
                if (S.hitem[warp][mid] <= it) lo = mid; else hi = mid - 1u;
            }
            const long long vhn = (long long)t.skew + S.hpos[warp][lo] - (long long)wlo;
            write_head_item<KW>(a, win, wwords, (int)vhn, a.group[S.grp[warp][lo]], S.hlo[warp][lo] + (it - S.hitem[warp][lo]));
        }
    }
    __syncwarp();
    // copy-out: window bytes [max(wlo, skew), wlo + wb) -> global, 16-byte rows, byte stores at the ragged ends
    uint8_t* gw = t.g0 - t.skew + wlo;                             // global address of window byte 0 (16-byte aligned)
    const u32 lo_b = wlo == 0 ? t.skew : 0u;                       // first valid byte of the window
    const u32 body_lo = (lo_b + 15u) & ~15u, body_hi = wb & ~15u;
    if (body_lo <= body_hi) {
        for (u32 i = lo_b + lane; i < body_lo; i += 32) gw[i] = stage[i];
        const uint4* sv = reinterpret_cast<const uint4*>(stage);
        uint4* gv = reinterpret_cast<uint4*>(gw);
        for (u32 i = body_lo / 16u + lane; i < body_hi / 16u; i += 32) __stcs(gv + i, sv[i]);   // streamed once
        for (u32 i = body_hi + lane; i < wb; i += 32) gw[i] = stage[i];
    } else {
        for (u32 i = lo_b + lane; i < wb; i += 32) gw[i] = stage[i];
    }
}

template <int KW>
__global__ void __launch_bounds__(EM_THREADS, KW == 1 ? 4 : (KW == 2 ? 3 : 2)) emit_write_kernel(EmitArgs a) {
    extern __shared__ __align__(16) uint8_t stage_all[];
    __shared__ WriteSmem S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all + (size_t)warp * a.stage_bytes;
    const u64 n_tiles = (a.n_last - a.n_first + 31) / 32;
    const u64 stride = (u64)gridDim.x * EW_WARPS;
    for (u64 tile = (u64)blockIdx.x * EW_WARPS + warp; tile < n_tiles; tile += stride) {
        if (tile + stride < n_tiles) {   // this warp's next tile: its node list and offsets into L2 while this one is written
            const u64 nn = a.n_first + (tile + stride) * 32;
            constexpr int DENSE_LINES = (32 * (KW + 1) * 8 + 127) / 128 + 1;   // 32 nodes, not line-aligned
            if (lane < DENSE_LINES) prefetch_l2(reinterpret_cast<const char*>(a.dense + nn * (KW + 1)) + 128 * lane);
            else if (lane < DENSE_LINES + 3) prefetch_l2(reinterpret_cast<const char*>(a.rec_offsets + nn) + 128 * (lane - DENSE_LINES));
        }
        WriteTile<KW> t;
        write_tile_load<KW>(a, tile, lane, t);
        if (t.span > (u64)EW_MAX_WINDOWS * a.stage_bytes) {
            if (lane == 0) a.big_tiles[atomicAdd(a.big_tile_count, 1ull)] = tile;
            continue;
        }
        for (u64 wlo = 0; wlo < t.span; wlo += a.stage_bytes) write_tile_window<KW>(a, t, wlo, stage, S, lane, warp);
        __syncwarp();
    }
}

// The listed tiles, one after the other, each with its windows dealt to all warps of the grid.
template <int KW>
__global__ void __launch_bounds__(EM_THREADS) emit_write_big_kernel(EmitArgs a) {
    extern __shared__ __align__(16) uint8_t stage_all[];
    __shared__ WriteSmem S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all + (size_t)warp * a.stage_bytes;
    const u64 n_big = *a.big_tile_count;
    const u64 gw = (u64)blockIdx.x * EW_WARPS + warp, n_gw = (u64)gridDim.x * EW_WARPS;
    for (u64 b = 0; b < n_big; ++b) {
        WriteTile<KW> t;
        write_tile_load<KW>(a, a.big_tiles[b], lane, t);
        for (u64 wlo = gw * a.stage_bytes; wlo < t.span; wlo += n_gw * a.stage_bytes) write_tile_window<KW>(a, t, wlo, stage, S, lane, warp);
        __syncwarp();
    }
}

// Fused graph statistics over the dense node list (GraphStatistics.java:78-131; Node.java:820-848).
struct GraphStatsDev {
    u64 nodes, degree_total, degree_max, degree_bins[17], coverage_total, coverage_max, coverage_bins[257];
    u64 unflipped, flipped, self_edges[4], path_nodes, tips_forward, tips_reverse, tips_both, tips_one;
    // GraphStatistics.java:87-119: kmerLength, <x>-with-FORWARD / -with-REVERSE (nodes with degree(dir) != 0), scaffoldSeedScore
    u64 kmer_length_total, kmer_length_max;
    u64 nodes_with_dir[2], coverage_with_dir_total[2], coverage_with_dir_max[2];
    u64 seed_nodes, seed_score_total, seed_score_max;
    u64 seed_nodes_with_dir[2], seed_score_with_dir_total[2], seed_score_with_dir_max[2];
};
static constexpr u64 GS_COVERAGE_WINDOW = 1000;   // COVERAGE_DIST_MEAN 0 +- COVERAGE_DIST_STD 1000 (GraphStatistics.java:75-76)

template <int KW>
__global__ void __launch_bounds__(256) graph_stats_kernel(EmitArgs a, GraphStatsDev* __restrict__ out) {
    constexpr int DW = KW + 1;
    __shared__ u32 s_deg[17];
    __shared__ u32 s_cov[257];
    for (int i = threadIdx.x; i < 17; i += 256) s_deg[i] = 0;
    for (int i = threadIdx.x; i < 257; i += 256) s_cov[i] = 0;
    __syncthreads();
    u64 deg_tot = 0, cov_tot = 0, unfl = 0, fl = 0, path = 0, tf = 0, tr = 0, tb = 0, to = 0, nodes = 0;
    u64 deg_max = 0, cov_max = 0, self[4] = {0, 0, 0, 0};
    u64 wd_nodes[2] = {0, 0}, wd_cov[2] = {0, 0}, wd_cov_max[2] = {0, 0};
    u64 seed_n = 0, seed_tot = 0, seed_max = 0, wd_seed_n[2] = {0, 0}, wd_seed_tot[2] = {0, 0}, wd_seed_max[2] = {0, 0};
    for (u64 n = (u64)blockIdx.x * 256 + threadIdx.x; n < a.n_nodes; n += (u64)gridDim.x * 256) {
        const u64* d = a.dense + n * DW;
        u64 key[KW];
#pragma unroll
        for (int j = 0; j < KW; ++j) key[j] = d[j];
        const u64 val = d[KW];
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
        u64 n_heads = 0;
        if (val & HEADS_FLAG) {
            const HeadGroup g = a.group[a.dense_h[n]];
            unfl += g.nu; fl += g.nf;
            n_heads = (u64)g.nu + g.nf;
        }
        // calculateSeedScore (Node.java:859-862): kmer length * read heads, counted for coverage within the window
        const u64 seed = (u64)a.k * n_heads;
        const bool in_window = cov <= GS_COVERAGE_WINDOW;
        if (in_window) { ++seed_n; seed_tot += seed; seed_max = max(seed_max, seed); }
        const u32 dir_deg[2] = {out_deg, in_deg};   // DIR.FORWARD = {FF, FR}, DIR.REVERSE = {RF, RR}
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            if (dir_deg[d] == 0) continue;
            ++wd_nodes[d]; wd_cov[d] += cov; wd_cov_max[d] = max(wd_cov_max[d], cov);
            if (in_window) { ++wd_seed_n[d]; wd_seed_tot[d] += seed; wd_seed_max[d] = max(wd_seed_max[d], seed); }
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
    atomicMax(&out->seed_score_max, seed_max);
    for (int d = 0; d < 2; ++d) {
        atomicMax(&out->coverage_with_dir_max[d], wd_cov_max[d]);
        atomicMax(&out->seed_score_with_dir_max[d], wd_seed_max[d]);
    }
    {
        const u64 r_sn = block_reduce_sum<256>(seed_n), r_st = block_reduce_sum<256>(seed_tot);
        u64 r_wn[2], r_wc[2], r_wsn[2], r_wst[2];
        for (int d = 0; d < 2; ++d) {
            r_wn[d] = block_reduce_sum<256>(wd_nodes[d]); r_wc[d] = block_reduce_sum<256>(wd_cov[d]);
            r_wsn[d] = block_reduce_sum<256>(wd_seed_n[d]); r_wst[d] = block_reduce_sum<256>(wd_seed_tot[d]);
        }
        if (threadIdx.x == 0) {
            atomicAdd(&out->seed_nodes, r_sn); atomicAdd(&out->seed_score_total, r_st);
            atomicAdd(&out->kmer_length_total, r_nodes * (u64)a.k);
            if (r_nodes) atomicMax(&out->kmer_length_max, (u64)a.k);
            for (int d = 0; d < 2; ++d) {
                atomicAdd(&out->nodes_with_dir[d], r_wn[d]); atomicAdd(&out->coverage_with_dir_total[d], r_wc[d]);
                atomicAdd(&out->seed_nodes_with_dir[d], r_wsn[d]); atomicAdd(&out->seed_score_with_dir_total[d], r_wst[d]);
            }
        }
    }
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

// Unclipped "coverage-bins" (GraphStatistics.java:89: one counter per Math.round(coverage) value), the input of the
// driver's FittingMixture cut-off (GenomixDriver.java:120-137). Low bins are gathered per CTA in shared memory.
static constexpr int CH_SMEM_BINS = 4096;

template <int KW>
__global__ void __launch_bounds__(256) coverage_histogram_kernel(EmitArgs a, u64* __restrict__ bins, u64 n_bins) {
    constexpr int DW = KW + 1;
    __shared__ u32 s_bins[CH_SMEM_BINS];
    for (int i = threadIdx.x; i < CH_SMEM_BINS; i += 256) s_bins[i] = 0;
    __syncthreads();
    for (u64 n = (u64)blockIdx.x * 256 + threadIdx.x; n < a.n_nodes; n += (u64)gridDim.x * 256) {
        const u64 cov = a.dense[n * DW + KW] & COUNT_MASK;
        if (cov < CH_SMEM_BINS) atomicAdd(&s_bins[cov], 1u);
        else if (cov < n_bins) atomicAdd(bins + cov, 1ull);
    }
    __syncthreads();
    for (u64 i = threadIdx.x; i < CH_SMEM_BINS && i < n_bins; i += 256)
        if (s_bins[i]) atomicAdd(bins + i, (u64)s_bins[i]);
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

}  // namespace gx
