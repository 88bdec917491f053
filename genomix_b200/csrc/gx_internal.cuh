// gx_internal.cuh -- shared device/host definitions of libgenomix_gb (sm_100a only).
//
// Data model:
//   * A k-mer is a little-endian multi-word integer V = sum_i code(letter i) * 4^i, KW = ceil(k/32)
//     64-bit words. Serialising V big-endian into ceil(k/4) bytes gives exactly the reference's
//     Kmer byte layout (genomix-data/.../types/Kmer.java:225-242), and unsigned integer order on V
//     equals the reference's byte-wise compareTo, so canonical = min(V_fwd, V_rc), tie -> FORWARD
//     (ReadsKeyValueParserFactory.java:163,181).
//   * A node's four VKmerLists are a pure function of (canonical key, 16-bit mask): bit 4*type+base,
//     type in {FF,FR,RF,RR} (EDGETYPE.java:6-9). F* lists hold Y = X[1:]+base (FF) or rc(Y) (FR);
//     R* lists hold Y = base+X[:-1] (RR) or rc(Y) (RF). Set exactly as setEdgesForCurAndNext does
//     (ReadsKeyValueParserFactory.java:209-233).
//   * Table value word: bits 0..46 occurrence count (the reference's float coverage sum), bit 47 "has read
//     heads" (set at finish), bits 48..63 edge mask.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef GX_MAX_KW
#define GX_MAX_KW 4
#endif

namespace gx {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 EMPTY_WORD = ~0ull;           // key words of a free slot (never a canonical key)
static constexpr u64 VAL_LOCK = ~0ull;             // value word of a slot being claimed (KW >= 3 protocol)
static constexpr u64 COUNT_MASK = (1ull << 47) - 1;  // value word bits 0..46: occurrence count
static constexpr u64 HEADS_FLAG = 1ull << 47;        // value word bit 47: the node has read heads (set by gx_finish)
static constexpr int MASK_SHIFT = 48;

// ---------------------------------------------------------------------------------------------
// line descriptors produced by the parse kernel, consumed by the extract kernel (per chunk)
struct LineDesc {
    u32 off[2];      // text offset (within the chunk) of mate 0 / mate 1 letters
    u32 len[2];      // letters (0 = field absent or empty)
    u64 read_id;     // Long.parseLong(field 0)
    u64 store[2];    // byte offset of each mate's packed sequence in the read store
    u64 head_idx[2]; // global index into the heads array for each split mate (valid only if flag set)
    u32 flags;       // bit0/bit1: mate 0/1 passes [ACGTacgt]+ (is split); bit2: line has 3 fields
    u32 pad;
    u64 occ_base;    // chunk-relative index of this line's first k-mer occurrence (mate 0 first, then mate 1)
};

// One ReadHeadInfo to be (ReadHeadInfo.java:17-28): lives until emit.
template <int KW>
struct Head {
    u64 key[KW];   // canonical first k-mer of the read (node that holds this head)
    u64 uuid;      // offset<<40 | library<<36 | mate<<35 | readId  (ReadHeadInfo.java:100-127)
    u64 order;     // input order of the line the head comes from: rank << 48 | line number (TreeSet ties: first line wins)
    u64 this_off;  // read store offsets of the packed sequences (VKmer byte order, no header)
    u64 mate_off;
    u32 this_len;  // letters
    u32 mate_len;  // letters; 0 = no mate sequence
    u32 flipped;   // 1: goes to flippedReadIds (first k-mer was REVERSE)
    u32 valid;     // 0 until the extract kernel filled it
};

// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 mix64(u64 x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

template <int KW>
__host__ __device__ __forceinline__ u64 hash_key(const u64 (&w)[KW]) {
    u64 h = mix64(w[0] + 0x9e3779b97f4a7c15ull);
#pragma unroll
    for (int i = 1; i < KW; ++i) h = mix64(h ^ (w[i] + 0x9e3779b97f4a7c15ull * (u64)(i + 1)));
    return h;
}

// One 64-bit hash h of the key drives every placement decision, most significant bits first:
//   owner rank   = floor(h * n_ranks / 2^64)                     (which GPU's table holds the key)
//   local hash   = h * n_ranks mod 2^64                          (uniform again inside the owner's range)
//   table region = floor(local * n_regions / 2^64)               (bucket of the region-sorted build, gx_split.cuh)
//   home slot    = floor(local * capacity / 2^64)
// so regions are contiguous slot ranges and one multisplit pass serves both the exchange and the L2-blocked insert.
__host__ __device__ __forceinline__ u64 mulhi64(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
__host__ __device__ __forceinline__ u32 owner_of(u64 h, u32 n_ranks) { return (u32)mulhi64(h, (u64)n_ranks); }
__host__ __device__ __forceinline__ u64 local_hash(u64 h, u64 n_ranks) { return h * n_ranks; }
// (the table kernels take the multiplier as `hash_mul`: n_ranks for the job's table; n_ranks * n_regions for the pilot
//  table of gx_api.cu, which holds region 0 only and so spreads that region's hash range over all of its slots)
__host__ __device__ __forceinline__ u32 region_of(u64 hl, u32 n_regions) { return (u32)mulhi64(hl, (u64)n_regions); }
__host__ __device__ __forceinline__ u64 slot_of(u64 hl, u64 capacity) { return mulhi64(hl, capacity); }

// reverse the order of the 32 2-bit groups of x and complement them (A<->T, C<->G: 3 - code)
__host__ __device__ __forceinline__ u64 revcomp_word(u64 x) {
#ifdef __CUDA_ARCH__
    u64 r = __brevll(x);
#else
    u64 r = x;
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
    r = ((r >> 4) & 0x0f0f0f0f0f0f0f0full) | ((r & 0x0f0f0f0f0f0f0f0full) << 4);
    r = __builtin_bswap64(r);
#endif
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);  // un-swap bits inside each group
    return ~r;
}

// rc = reverse complement of the k-letter value f
template <int KW>
__host__ __device__ __forceinline__ void revcomp_key(const u64 (&f)[KW], int k, u64 (&rc)[KW]) {
    u64 t[KW];
#pragma unroll
    for (int i = 0; i < KW; ++i) t[i] = revcomp_word(f[KW - 1 - i]);
    // t holds the reversed letters left-aligned in 64*KW bits: shift right by 64*KW - 2k
    const int sh = 64 * KW - 2 * k;  // 0 <= sh < 64
    if (sh == 0) {
#pragma unroll
        for (int i = 0; i < KW; ++i) rc[i] = t[i];
    } else {
#pragma unroll
        for (int i = 0; i < KW; ++i) {
            u64 lo = t[i] >> sh;
            u64 hi = (i + 1 < KW) ? (t[i + 1] << (64 - sh)) : 0ull;
            rc[i] = lo | hi;
        }
    }
}

// a <= b as multi-word unsigned integers (word KW-1 most significant)
template <int KW>
__host__ __device__ __forceinline__ bool key_le(const u64 (&a)[KW], const u64 (&b)[KW]) {
#pragma unroll
    for (int i = KW - 1; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] < b[i];
    }
    return true;
}

template <int KW>
__host__ __device__ __forceinline__ bool key_eq(const u64 (&a)[KW], const u64 (&b)[KW]) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < KW; ++i) e = e && (a[i] == b[i]);
    return e;
}

__host__ __device__ __forceinline__ u64 top_word_mask(int k) {
    const int r = (2 * k) & 63;
    return r == 0 ? ~0ull : ((1ull << r) - 1);
}

// Y = X[1:] + base  (drop letter 0, append at position k-1)
template <int KW>
__host__ __device__ __forceinline__ void key_append(const u64 (&x)[KW], int k, u32 base, u64 (&y)[KW]) {
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        u64 lo = x[i] >> 2;
        u64 hi = (i + 1 < KW) ? (x[i + 1] << 62) : 0ull;
        y[i] = lo | hi;
    }
    const int pos = 2 * (k - 1);
#pragma unroll
    for (int i = 0; i < KW; ++i)  // static indices keep y[] in registers
        if (i == (pos >> 6)) y[i] |= (u64)base << (pos & 63);
}

// Y = base + X[:-1]  (prepend at position 0, drop letter k-1)
template <int KW>
__host__ __device__ __forceinline__ void key_prepend(const u64 (&x)[KW], int k, u32 base, u64 (&y)[KW]) {
#pragma unroll
    for (int i = KW - 1; i >= 0; --i) {
        u64 hi = x[i] << 2;
        u64 lo = (i > 0) ? (x[i - 1] >> 62) : (u64)base;
        y[i] = hi | lo;
    }
    y[KW - 1] &= top_word_mask(k);
}

// neighbour key for edge bit (type, base) of node X (see header comment)
template <int KW>
__host__ __device__ __forceinline__ void neighbour_key(const u64 (&x)[KW], int k, int type, u32 base, u64 (&n)[KW]) {
    u64 y[KW];
    if (type < 2) key_append<KW>(x, k, base, y); else key_prepend<KW>(x, k, base, y);
    if (type == 1 || type == 2) revcomp_key<KW>(y, k, n);
    else {
#pragma unroll
        for (int i = 0; i < KW; ++i) n[i] = y[i];
    }
}

// Edge-mask bits of one occurrence, exactly as setEdgesForCurAndNext assigns them
// (ReadsKeyValueParserFactory.java:209-233) seen from the canonical key (gx_internal.cuh header):
//   to next  : cur F -> FF/FR with base b;      cur R -> RF/RR with base 3-b
//   from prev: cur F -> RR/RF with base a;      cur R -> FR/FF with base 3-a
__host__ __device__ __forceinline__ u32 edge_bit_next(bool cur_rev, bool next_rev, u32 b) {
    const u32 type = cur_rev ? (next_rev ? 3u : 2u) : (next_rev ? 1u : 0u);
    return 1u << (type * 4u + (cur_rev ? 3u - b : b));
}
__host__ __device__ __forceinline__ u32 edge_bit_prev(bool cur_rev, bool prev_rev, u32 a) {
    const u32 type = cur_rev ? (prev_rev ? 0u : 1u) : (prev_rev ? 2u : 3u);
    return 1u << (type * 4u + (cur_rev ? 3u - a : a));
}

// ASCII -> 2-bit code (A0 C1 G2 T3, either case; GeneCode.java:29-50); ok=false for anything else (code 0)
__host__ __device__ __forceinline__ u32 code_of(u32 c, bool& ok) {
    const u32 u = c | 0x20u;
    ok = (u == 'a') | (u == 'c') | (u == 'g') | (u == 't');
    return ok ? (((c >> 1) ^ (c >> 2)) & 3u) : 0u;
}

// ---------------------------------------------------------------------------------------------
// hash-table slot layouts
template <int KW> struct SlotTraits;
template <> struct SlotTraits<1> { static constexpr int WORDS = 2; };  // key, val                      16 B
template <> struct SlotTraits<2> { static constexpr int WORDS = 4; };  // k0 k1 val pad                 32 B (128-bit CAS on k0:k1)
template <> struct SlotTraits<3> { static constexpr int WORDS = 4; };  // k0 k1 k2 val                  32 B (claim on val)
template <> struct SlotTraits<4> { static constexpr int WORDS = 6; };  // k0..k3 val pad                48 B (claim on val)

template <int KW> __host__ __device__ __forceinline__ constexpr int slot_words() { return SlotTraits<KW>::WORDS; }
template <int KW> __host__ __device__ __forceinline__ constexpr int val_index() { return KW; }

}  // namespace gx
