// gx_merge.cuh -- folding serialised `VKmer key | Node` records back into the table (gx_push_records).
//
// Restates the merge half of the reference's aggregator, AggregateKmerAggregateFactory.aggregate
// (genomix-hyracks/.../graph/dataflow/AggregateKmerAggregateFactory.java:128-144): the accumulated Node of a key and
// another serialised Node of the same key become one Node -- per edge type VKmerList.unionUpdate, ReadHeadSet.unionUpdate
// (TreeSet.addAll), coverage added. Here the "accumulated" side is the hash table: a record turns back into what the
// build kernels produce from reads -- (key, 16-bit edge mask, occurrence count) plus Head entries and packed sequences --
// and goes through the same upsert and the same read-head grouping as everything else.
//
//   merge_scan_kernel   pass 1: validate the record framing, count read heads / packed-sequence bytes, reserve their room
//   merge_apply_kernel  pass 2: key words, edge mask, count; Head entries and their sequences
//
// Only graph-build Nodes are representable (every edge is a one-letter shift of the key, no internal kmer, read-head sets
// stored by value); anything else is reported as a format error, never folded in wrongly.
#pragma once
#include "gx_build.cuh"

namespace gx {

struct MergeArgs {
    const uint8_t* rec; const u64* rec_off; u64 n_rec; int k;
    u64* head_base; u64* store_base;        // per record: first Head index / read-store offset reserved for it
    Counters* ctr;
    void* heads; uint8_t* store;            // pass 2
    u64* keys; unsigned short* meta; u32* counts;
    u64 order_base;                         // input order of the first record (TreeSet ties: earlier input wins)
};

__device__ __forceinline__ u32 rd32be(const uint8_t* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3]; }
__device__ __forceinline__ u64 rd64be(const uint8_t* p) { return ((u64)rd32be(p) << 32) | (u64)rd32be(p + 4); }

// nb big-endian bytes (the reference's Kmer byte array, Kmer.java:225-242) -> little-endian 64-bit words
template <int KW>
__device__ __forceinline__ void key_from_bytes(const uint8_t* p, u32 nb, u64 (&w)[KW]) {
#pragma unroll
    for (int i = 0; i < KW; ++i) w[i] = 0;
    for (u32 j = 0; j < nb; ++j) {
        const u32 bit = 8u * (nb - 1u - j);
#pragma unroll
        for (int i = 0; i < KW; ++i)
            if ((u32)i == (bit >> 6)) w[i] |= (u64)p[j] << (bit & 63u);
    }
}

template <int KW>
__device__ __forceinline__ u32 key_letter(const u64 (&w)[KW], u32 pos) {
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < KW; ++i)
        if ((u32)i == (pos >> 5)) c = (u32)(w[i] >> (2u * (pos & 31u))) & 3u;
    return c;
}

// Walk one record. Returns false if it is malformed / not a graph-build Node. With APPLY the fields are handed to `f`.
// f.edge(type, entry bytes)  f.head(set, flags, uuid, this_len, this bytes, mate_len, mate bytes)  f.coverage(float bits)
template <class F>
__device__ __forceinline__ bool walk_record(const uint8_t* r, u64 len, int k, F& f) {
    const u32 nb = (u32)(k + 3) / 4u;
    if (len < 8u + 4u + nb + 1u) return false;
    const u32 rec_len = rd32be(r), key_len = rd32be(r + 4);
    if ((u64)rec_len + 8u != len || key_len != 4u + nb || rd32be(r + 8) != (u32)k) return false;
    const uint8_t* p = r + 12 + nb;
    const uint8_t* end = r + len;
    const u32 active = *p++;
    if (active & (1u << 6)) return false;   // internal kmer: not a graph-build Node
    for (u32 t = 0; t < 4; ++t) {
        if (!(active & (1u << t))) continue;
        if (p + 4 > end) return false;
        const u32 cnt = rd32be(p);
        p += 4;
        if (cnt == 0 || cnt > 4u) return false;
        for (u32 e = 0; e < cnt; ++e) {
            if (p + 4 + nb > end || rd32be(p) != (u32)k) return false;
            if (!f.edge(t, p + 4)) return false;
            p += 4 + nb;
        }
    }
    for (u32 set = 0; set < 2; ++set) {
        if (!(active & (1u << (4 + set)))) continue;
        if (p + 5 > end || p[0] != 1) return false;   // stored by path reference: not produced by graph build
        const u32 n = rd32be(p + 1);
        p += 5;
        for (u32 e = 0; e < n; ++e) {
            if (p + 13 > end) return false;
            const u32 flags = p[0];
            const u64 uuid = rd64be(p + 1);
            const u32 this_len = rd32be(p + 9);
            const uint8_t* this_bytes = p + 13;
            const u32 tb = (this_len + 3u) / 4u;
            p += 13 + tb;
            u32 mate_len = 0;
            const uint8_t* mate_bytes = nullptr;
            if (flags & 1u) {
                if (p + 4 > end) return false;
                mate_len = rd32be(p);
                mate_bytes = p + 4;
                p += 4 + (mate_len + 3u) / 4u;
            }
            if (p > end) return false;
            f.head(set, uuid, this_len, this_bytes, mate_len, mate_bytes);
        }
    }
    if (!(active & (1u << 7)) || p + 4 != end) return false;
    f.coverage(rd32be(p));
    return true;
}

struct MergeCounter {
    u32 heads = 0, store_bytes = 0;
    __device__ __forceinline__ bool edge(u32, const uint8_t*) { return true; }
    __device__ __forceinline__ void head(u32, u64, u32 this_len, const uint8_t*, u32 mate_len, const uint8_t*) {
        ++heads;
        store_bytes += (this_len + 3u) / 4u + (mate_len + 3u) / 4u;
    }
    __device__ __forceinline__ void coverage(u32) {}
};

static constexpr u32 LE_RECORD = 8;   // Counters::error code: malformed / unsupported Node record (record index as "line")

template <int KW>
__global__ void __launch_bounds__(256) merge_scan_kernel(MergeArgs a) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_rec) return;
    MergeCounter mc;
    const bool ok = walk_record(a.rec + a.rec_off[i], a.rec_off[i + 1] - a.rec_off[i], a.k, mc);
    if (!ok) { report_line_error(a.ctr, i, LE_RECORD); a.head_base[i] = 0; a.store_base[i] = 0; return; }
    a.head_base[i] = mc.heads ? atomicAdd(&a.ctr->head_cursor, (u64)mc.heads) : 0ull;
    a.store_base[i] = mc.store_bytes ? atomicAdd(&a.ctr->store_cursor, (u64)mc.store_bytes) : 0ull;
}

template <int KW>
struct MergeApplier {
    const MergeArgs& a;
    u64 key[KW], rcx[KW];
    u32 mask = 0;
    u64 count = 0;
    u64 head_idx, store_off, order;
    bool bad = false;
    __device__ __forceinline__ MergeApplier(const MergeArgs& args) : a(args) {}
    // which one-letter shift of the key is this neighbour? (gx_internal.cuh header)
    //   FF b: X[1:]+b   FR b: (3-b)+rc(X)[:-1]   RF b: rc(X)[1:]+(3-b)   RR b: b+X[:-1]
    __device__ __forceinline__ bool edge(u32 t, const uint8_t* bytes) {
        const u32 nb = (u32)(a.k + 3) / 4u;
        u64 y[KW], nk[KW];
        key_from_bytes<KW>(bytes, nb, y);
        u32 b;
        if (t == 0) { b = key_letter<KW>(y, (u32)a.k - 1u); key_append<KW>(key, a.k, b, nk); }
        else if (t == 1) { b = 3u - key_letter<KW>(y, 0); key_prepend<KW>(rcx, a.k, 3u - b, nk); }
        else if (t == 2) { b = 3u - key_letter<KW>(y, (u32)a.k - 1u); key_append<KW>(rcx, a.k, 3u - b, nk); }
        else { b = key_letter<KW>(y, 0); key_prepend<KW>(key, a.k, b, nk); }
        if (!key_eq<KW>(nk, y)) return false;   // not a one-letter shift: a cleaned / merged graph, not a graph-build Node
        mask |= 1u << (4u * t + b);
        return true;
    }
    __device__ __forceinline__ void head(u32 set, u64 uuid, u32 this_len, const uint8_t* this_bytes, u32 mate_len, const uint8_t* mate_bytes) {
        Head<KW>& h = reinterpret_cast<Head<KW>*>(a.heads)[head_idx++];
        const u32 tb = (this_len + 3u) / 4u, mb = (mate_len + 3u) / 4u;
#pragma unroll
        for (int i = 0; i < KW; ++i) h.key[i] = key[i];
        h.uuid = uuid;
        h.order = order;
        h.this_off = store_off;
        h.mate_off = store_off + tb;
        h.this_len = this_len;
        h.mate_len = mate_len;
        h.flipped = set;
        for (u32 j = 0; j < tb; ++j) a.store[store_off + j] = this_bytes[j];
        for (u32 j = 0; j < mb; ++j) a.store[store_off + tb + j] = mate_bytes[j];
        store_off += tb + mb;
        h.valid = 1u;
    }
    __device__ __forceinline__ void coverage(u32 bits) {
        const float c = __uint_as_float(bits);
        if (!(c >= 1.0f) || c > 4.0e9f) { bad = true; return; }   // a node stands for at least one occurrence; counts travel as u32
        count = (u64)llrintf(c);
    }
};

template <int KW>
__global__ void __launch_bounds__(256) merge_apply_kernel(MergeArgs a) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 count = 0;
    if (i < a.n_rec) {
        const u32 nb = (u32)(a.k + 3) / 4u;
        const uint8_t* r = a.rec + a.rec_off[i];
        MergeApplier<KW> m(a);
        key_from_bytes<KW>(r + 12, nb, m.key);
        revcomp_key<KW>(m.key, a.k, m.rcx);
        m.head_idx = a.head_base[i];
        m.store_off = a.store_base[i];
        m.order = a.order_base + i;
        bool ok = walk_record(r, a.rec_off[i + 1] - a.rec_off[i], a.k, m) && !m.bad;
        // the key of a Node is canonical (ReadsKeyValueParserFactory.java:163,181): anything else was not written by graph build
        ok = ok && key_le<KW>(m.key, m.rcx);
        if constexpr (KW <= 2) {   // the all-ones word marks a free slot: only k = 32 * KW poly-T could collide, and that is not canonical
            bool empty = true;
#pragma unroll
            for (int j = 0; j < KW; ++j) empty = empty && m.key[j] == EMPTY_WORD;
            ok = ok && !empty;
        }
        if (!ok) { report_line_error(a.ctr, i, LE_RECORD); m.count = 0; }
#pragma unroll
        for (int j = 0; j < KW; ++j) a.keys[i * KW + j] = m.key[j];
        a.meta[i] = (unsigned short)m.mask;
        a.counts[i] = (u32)m.count;
        count = m.count;
    }
    const u64 tot = block_reduce_sum<256>(count);   // one barrier site for every thread of the CTA
    if (threadIdx.x == 0 && tot) atomicAdd(&a.ctr->occurrences, tot);
}

}  // namespace gx
