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
// K3a: read heads -> owning slot (the reference carries the ReadHeadInfo inside the first k-mer's tuple and
// unions TreeSets per key, AggregateKmerAggregateFactory.java:120-123,141-143)
template <int KW>
__global__ void __launch_bounds__(256) heads_count_kernel(const Head<KW>* __restrict__ heads, u64 n_heads,
                                                          const u64* __restrict__ table, u64 capacity,
                                                          u32 n_ranks, u64* __restrict__ hslot,
                                                          u32* __restrict__ hcount, Counters* ctr) {
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

}  // namespace gx
