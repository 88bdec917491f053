// gx_split.cuh -- the build: region-sorted k-mer records, L2-resident upserts.
//
// Replaces the reference's sort + pre-clustered group (ExternalSortOperatorDescriptor.java:119-194,
// PreclusteredGroupWriter.java:76-136) with ONE radix pass on the top bits of the key hash instead of a comparison sort,
// and the sender side of its M:N hash connector (MToNPartitioningMergingConnectorDescriptor.java:65-87) with the same pass:
//
//   split_count_kernel<KW>    read -> canonical k-mers -> histogram over buckets (owner GPU, table region), over every
//                             16th line: the room each bucket gets in the record arena (estimate + slack; exact counts
//                             for small chunks and for the rare chunk whose estimate a bucket overflows).
//   split_place_kernel<KW>    read -> canonical k-mers + edge masks + read heads -> CTA-level multisplit in shared memory
//                             -> bucket-sorted runs appended to each bucket's room. No table access: a pure function of
//                             the text.
//   upsert_regions_kernel<KW> persistent warps walk the arenas region by region (work items dealt by ticket), so that at
//                             any time all warps upsert into the same few MB of the table: L2 hits instead of one HBM
//                             row activation per k-mer occurrence.
//
// A random 16-byte table access that misses L2 costs a whole 128-byte line and a DRAM row activation (15.6 G upserts/s on
// B200, profiles/r01_microbench_random_access.txt); the same access L2-resident runs 4x faster. Sorting the records by
// region first costs one coalesced write + read of (8*KW + 2) bytes per occurrence and 1/16 of a second extraction pass.
#pragma once
#include "gx_build.cuh"

namespace gx {

#ifndef GX_SP_THREADS
#define GX_SP_THREADS 512
#endif
#ifndef GX_SP_MIN1
#define GX_SP_MIN1 2
#endif
#ifndef GX_SP_MIN2
#define GX_SP_MIN2 2
#endif
static constexpr int SP_THREADS = GX_SP_THREADS;
static constexpr int SP_WARPS = SP_THREADS / 32;
static constexpr int SP_NPL = 4;                // consecutive positions per lane and round
static constexpr int SP_MAX_BUCKETS = 1024;     // owners x regions handled by one multisplit
static constexpr int SP_MAX_RANKS = 64;
static constexpr int CURSOR_PAD = 16;           // bucket cursors live 128 B apart (one L2 line each)

struct SplitArgs {
    const uint8_t* text; u64 n_text;
    const LineDesc* desc; u64 n_lines;
    u64 first_line;         // rank << 48 | job-wide number of the chunk's first line (input order of read heads)
    int k;
    void* heads;
    uint8_t* store;
    Counters* ctr;
    u32 n_ranks;            // owner = owner_of(h, n_ranks)
    u32 n_regions;          // table regions per owner; bucket = owner * n_regions + region_of(local hash)
    u32 sample;             // count pass: every sample-th line is counted (1 = exact counts)
    u64* bucket_count;      // count pass: [n_buckets] records per bucket (of the sampled lines)
    u64* cursor;            // place pass: [n_buckets * CURSOR_PAD] next record index of each bucket
    const u64* limit;       // place pass: [n_buckets] end of each bucket's room in the arena
    u64* owner_keys[SP_MAX_RANKS];             // record area of each owner (KW words per record): the local arena or,
    unsigned short* owner_meta[SP_MAX_RANKS];  // with direct NVLink delivery, this rank's share of the owner's inbox
};

// CTAs per SM (profiles/r02_tune_split.txt): k <= 64 runs two 512-thread CTAs (KW = 2 at 64 registers with a few spilled
// words is 18 % faster than one CTA at 102 registers: the rounds are barrier-separated, a second CTA fills the gaps)
template <int KW> struct SplitBlocks { static constexpr int MIN = (KW == 1) ? GX_SP_MIN1 : (KW == 2) ? GX_SP_MIN2 : 1; };

template <int KW>
struct SplitSmem {
    static constexpr int T = SP_THREADS * SP_NPL;  // records per CTA round (upper bound)
    u64 win[SP_WARPS][WIN_WORDS];
    u64 keys[T * KW];
    u64 gbase[SP_MAX_BUCKETS];   // record index (in the owner's area) that stage position start[b] goes to
    u64 glimit[SP_MAX_BUCKETS];  // end of the bucket's room
    u32 cnt[SP_MAX_BUCKETS];     // records of this round per bucket (zero between rounds)
    u32 start[SP_MAX_BUCKETS];   // first stage position of the bucket's run
    unsigned short meta[T];
    unsigned short bkt[T];
    u32 warp_sums[SP_WARPS];
};

// A warp's walk over the split mates of its lines: line = line0 + i * stride, mate 0 then mate 1.
struct ReadCursor {
    u64 line0, stride, cand;    // next candidate: line0 + (cand >> 1) * stride, mate cand & 1
    u64 cur_line;
    u32 flags, len[2], off[2];  // hot fields of the current line's descriptor
};

// (Re)load the packed 2-bit window of the read so that letters [need_lo, need_hi) are present. Warp-uniform.
__device__ __forceinline__ void ensure_window(u64* W, const uint8_t* rd, u32 len, u32 need_lo, u32 need_hi, u32& win_lo, u32& win_hi,
                                              u32& win_m, const uint8_t* text_lo, const uint8_t* text_hi, int lane) {
    constexpr u32 WIN_LETTERS = 4u * WIN_BYTES - 4u;   // letters a window can hold whatever the source alignment
    if (need_lo >= win_lo && need_hi <= win_hi) return;
    uint8_t* Wb = reinterpret_cast<uint8_t*>(W);
    win_lo = need_lo;
    win_hi = min(len, win_lo + WIN_LETTERS);
    const uint8_t* src = rd + win_lo;
    win_m = (u32)((uintptr_t)src & 3u);
    const uint8_t* aligned = src - win_m;
    const u32 nwords = (win_m + (win_hi - win_lo) + 3u) / 4u;
    __syncwarp();
    for (u32 j = lane; j < nwords; j += 32) Wb[j] = (uint8_t)load_quad(aligned + 4 * j, text_lo, text_hi);
    __syncwarp();
}

// Canonical keys of positions q .. q+NPL-1 (those below pend) from the packed window: key = min(fwd, rc), bit j of
// revbits set when position q+j is REVERSE (ReadsKeyValueParserFactory.java:163,181: ties are FORWARD).
template <int KW>
__device__ __forceinline__ void lane_keys(const u64* W, u32 wbase, u32 q, u32 pend, int k, u64 (&key)[SP_NPL][KW], u32& revbits) {
    u64 f[KW], rc[KW];
    revbits = 0;
    if (q < pend) {
        window_kmer<KW>(W, q + wbase, k, f);
        revcomp_key<KW>(f, k, rc);
    }
#pragma unroll
    for (int j = 0; j < SP_NPL; ++j) {
        if (q + j < pend) {
            const bool rev = !key_le<KW>(f, rc);
            revbits |= (u32)rev << j;
#pragma unroll
            for (int i = 0; i < KW; ++i) key[j][i] = rev ? rc[i] : f[i];
            if (j + 1 < SP_NPL && q + j + 1 < pend) roll_kmer<KW>(f, rc, k, window_letter(W, q + j + (u32)k + wbase));
        }
    }
}

__device__ __forceinline__ u32 bucket_of_hash(u64 h, u32 n_ranks, u32 n_regions) {
    return owner_of(h, n_ranks) * n_regions + region_of(local_hash(h, n_ranks), n_regions);
}

// K1a. Histogram of the chunk's k-mer occurrences over buckets. Same extraction as the placement pass, no output but the
// counts: warps run independently (no CTA rounds), the CTA's histogram is flushed once at the end.
template <int KW>
__global__ void __launch_bounds__(SP_THREADS, SplitBlocks<KW>::MIN) split_count_kernel(SplitArgs a) {
    __shared__ u64 win[SP_WARPS][WIN_WORDS];
    __shared__ u32 hist[SP_MAX_BUCKETS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 b = tid; b < SP_MAX_BUCKETS; b += SP_THREADS) hist[b] = 0;
    __syncthreads();
    u64* W = win[warp];
    const int k = a.k;
    const uint8_t* text_lo = a.text;
    const uint8_t* text_hi = a.text + a.n_text;
    constexpr u32 CHUNK = 32u * SP_NPL;
    for (u64 line = ((u64)blockIdx.x * SP_WARPS + warp) * a.sample; line < a.n_lines; line += (u64)gridDim.x * SP_WARPS * a.sample) {
        const LineDesc& dd = a.desc[line];
        const u32 flags = dd.flags;
        if ((flags & 3u) == 0) continue;
#pragma unroll 1
        for (int mate = 0; mate < 2; ++mate) {
            if (!(flags & (1u << mate))) continue;
            const u32 len = dd.len[mate];
            const uint8_t* rd = a.text + dd.off[mate];
            const u32 npos = len - (u32)k + 1u;
            u32 win_lo = 0, win_hi = 0, win_m = 0;
            for (u32 p0 = 0; p0 < npos; p0 += CHUNK) {
                const u32 pend = min(p0 + CHUNK, npos);
                ensure_window(W, rd, len, p0, min(pend + (u32)k, len), win_lo, win_hi, win_m, text_lo, text_hi, lane);
                const u32 wbase = win_m - win_lo;
                const u32 q = p0 + (u32)SP_NPL * (u32)lane;
                u64 key[SP_NPL][KW];
                u32 revbits;
                lane_keys<KW>(W, wbase, q, pend, k, key, revbits);
#pragma unroll
                for (int j = 0; j < SP_NPL; ++j)
                    if (q + j < pend) atomicAdd(&hist[bucket_of_hash(hash_key<KW>(key[j]), a.n_ranks, a.n_regions)], 1u);
            }
        }
    }
    __syncthreads();
    const u32 nb = a.n_ranks * a.n_regions;
    for (u32 b = tid; b < nb; b += SP_THREADS)
        if (hist[b]) atomicAdd(a.bucket_count + b, (u64)hist[b]);
}

// Room for every bucket in the arena, from the (sampled) counts: bucket b gets cap_b = count_b * sample, plus -- when the
// counts are an estimate (sample > 1) -- a sixteenth and 2048 records of slack (the estimate of a 2.8 M record bucket from
// every 16th line is good to a few thousand records; a bucket that still overflows makes the host redo the chunk with
// exact counts). Per owner o the table tab[o] = { start of its R regions, end of its block, R counts (split_finish_kernel) }
// is what the owner's upsert reads -- locally or, with its block, on another GPU.
static __global__ void __launch_bounds__(1024) split_prefix_kernel(const u64* __restrict__ bucket_count, u32 n_buckets, u32 n_regions,
                                                            u32 sample, u64 arena_cap, u64* __restrict__ tab, u64* __restrict__ owner_off,
                                                            u64* __restrict__ cursor, u64* __restrict__ limit, Counters* ctr) {
    __shared__ u64 sh_start[SP_MAX_BUCKETS + 1];
    const u32 b = threadIdx.x;
    u64 v = b < n_buckets ? bucket_count[b] * sample : 0ull;
    if (b < n_buckets && sample > 1) v += v / 16 + 2048;
    u64 tot;
    const u64 ex = block_scan_excl<1024>(v, &tot);
    if (b < n_buckets) sh_start[b] = ex;
    if (b == 0) { sh_start[n_buckets] = tot; ctr->split_overflow = tot > arena_cap ? 1ull : 0ull; }
    __syncthreads();
    if (b < n_buckets) {
        const u32 o = b / n_regions, r = b % n_regions;
        u64* t = tab + (size_t)o * (2 * n_regions + 1);
        t[r] = sh_start[b];
        if (r + 1 == n_regions) t[n_regions] = sh_start[b + 1];
        if (r == 0) owner_off[o] = sh_start[b];
        cursor[(size_t)b * CURSOR_PAD] = sh_start[b];
        limit[b] = sh_start[b + 1];
    }
    if (b == 0) owner_off[n_buckets / n_regions] = tot;
}

// after the placement: records per bucket
static __global__ void __launch_bounds__(1024) split_finish_kernel(const u64* __restrict__ cursor, const u64* __restrict__ limit,
                                                            u32 n_buckets, u32 n_regions, u64* __restrict__ tab) {
    const u32 b = threadIdx.x;
    if (b >= n_buckets) return;
    const u32 o = b / n_regions, r = b % n_regions;
    u64* t = tab + (size_t)o * (2 * n_regions + 1);
    t[n_regions + 1 + r] = min(cursor[(size_t)b * CURSOR_PAD], limit[b]) - t[r];
}

// K1b. Restates ReadsKeyValueParserFactory.SplitReads (:150-196) and setEdgesForCurAndNext (:209-233): for every position p
// of every split mate the forward and reverse-complement k-mers, dir = fwd <= rc ? FORWARD : REVERSE, key = the smaller, one
// tuple (key, Node{coverage 1, <=2 edges, read head on p == 0}). The tuples leave the kernel as (key, 16-bit edge mask)
// records sorted by bucket; read heads go to the heads array, packed reads to the read store.
// A warp owns one read at a time and handles 32*NPL consecutive positions per CTA round: lane l computes the k-mer at
// position p0 + NPL*l from the packed window and rolls it forward NPL-1 times; the directions of the neighbouring
// positions come from the adjacent lanes (shuffles). The last position of a full chunk is a halo (computed for its
// direction, emitted by the next chunk).
template <int KW>
__global__ void __launch_bounds__(SP_THREADS, SplitBlocks<KW>::MIN) split_place_kernel(SplitArgs a) {
    constexpr int NPL = SP_NPL;
    extern __shared__ __align__(16) unsigned char split_smem_raw[];
    SplitSmem<KW>& S = *reinterpret_cast<SplitSmem<KW>*>(split_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 b = tid; b < SP_MAX_BUCKETS; b += SP_THREADS) S.cnt[b] = 0;
    u64* W = S.win[warp];
    Head<KW>* heads = reinterpret_cast<Head<KW>*>(a.heads);
    const int k = a.k;
    const uint8_t* text_lo = a.text;
    const uint8_t* text_hi = a.text + a.n_text;
    constexpr u32 CHUNK = 32u * NPL;

    // this warp's walk over (line, mate) candidates
    const u64 line0 = (u64)blockIdx.x * SP_WARPS + warp;
    const u64 line_stride = (u64)gridDim.x * SP_WARPS;
    u64 cand = 0;               // next candidate: line0 + (cand >> 1) * line_stride, mate cand & 1
    bool in_read = false;
    u64 cur_line = 0;           // only the hot fields of the line descriptor stay in registers
    u32 d_flags = 0, d_len[2] = {0, 0}, d_off[2] = {0, 0};
    int mate = 0;
    u32 len = 0, npos = 0, p0 = 0, win_lo = 0, win_hi = 0, win_m = 0, carry_rev = 0;
    const uint8_t* rd = nullptr;

    for (;;) {
        // ---- 1. a read with positions left
        while (!in_read) {
            const u64 ln = line0 + (cand >> 1) * line_stride;
            if (ln >= a.n_lines) break;
            const int mt = (int)(cand & 1);
            ++cand;
            if (mt == 0) {
                const LineDesc& dd = a.desc[ln];
                cur_line = ln;
                d_flags = dd.flags; d_len[0] = dd.len[0]; d_len[1] = dd.len[1]; d_off[0] = dd.off[0]; d_off[1] = dd.off[1];
                if ((d_flags & 3u) == 0) { ++cand; continue; }   // nothing of this line is split: skip both mates
            }
            const u32 l = mt ? d_len[1] : d_len[0];
            if (l == 0) continue;
            const uint8_t* letters = a.text + (mt ? d_off[1] : d_off[0]);
            pack_read_to_store(letters, l, a.store + a.desc[ln].store[mt], lane);
            if (!(d_flags & (1u << mt))) continue;
            mate = mt; len = l; npos = l - (u32)k + 1u; p0 = 0; rd = letters;
            win_lo = win_hi = 0; carry_rev = 0;
            in_read = true;
        }
        if (!__syncthreads_or(in_read ? 1 : 0)) break;   // also separates the rounds' use of the stage

        // ---- 2. this warp's chunk: positions [p0, pend) computed, [p0, pemit) emitted
        u64 key[NPL][KW];
        u32 msk[NPL];
        u32 emit = 0;   // bit j: record j of this lane is emitted
        if (in_read) {
            const u32 pend = min(p0 + CHUNK, npos);
            const u32 pemit = (p0 + CHUNK < npos) ? p0 + CHUNK - 1u : npos;
            ensure_window(W, rd, len, p0 > 0 ? p0 - 1u : 0u, min(pend + (u32)k, len), win_lo, win_hi, win_m, text_lo, text_hi, lane);
            const u32 wbase = win_m - win_lo;   // letter x sits at packed index x + wbase (mod 2^32)
            const u32 q = p0 + (u32)NPL * (u32)lane;
            u32 revbits;
            lane_keys<KW>(W, wbase, q, pend, k, key, revbits);
            u32 prev_first = __shfl_up_sync(0xffffffffu, (revbits >> (NPL - 1)) & 1u, 1);
            if (lane == 0) prev_first = carry_rev;
            const u32 next_last = __shfl_down_sync(0xffffffffu, revbits & 1u, 1);
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const u32 p = q + j;
                msk[j] = 0;
                if (p < pemit) {
                    emit |= 1u << j;
                    const bool rev = (revbits >> j) & 1u;
                    const u32 prev_rev = j ? (revbits >> (j - 1)) & 1u : prev_first;
                    const u32 next_rev = (j + 1 < NPL) ? (revbits >> (j + 1)) & 1u : next_last;
                    u32 m = 0;
                    if (p + 1u < npos) m |= edge_bit_next(rev, next_rev, window_letter(W, p + (u32)k + wbase));
                    if (p > 0u) m |= edge_bit_prev(rev, prev_rev, window_letter(W, p - 1u + wbase));
                    msk[j] = m;
                    if (p == 0u) write_head<KW>(heads, a.desc[cur_line], a.first_line + cur_line, mate, key[j], rev, k);
                }
            }
            // direction of position pemit - 1 for the next chunk (only meaningful after a full chunk)
            carry_rev = __shfl_sync(0xffffffffu, (revbits >> (NPL - 2)) & 1u, 31);
            p0 = pemit;
            if (p0 >= npos) in_read = false;
        }

        // ---- 3. multisplit of the round's records by bucket
        u32 bk[NPL], rk[NPL];
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            if ((emit >> j) & 1u) {
                bk[j] = bucket_of_hash(hash_key<KW>(key[j]), a.n_ranks, a.n_regions);
                rk[j] = atomicAdd(&S.cnt[bk[j]], 1u);
            }
        }
        __syncthreads();
        // exclusive scan over the buckets (SP_BPT consecutive ones per thread) and ONE global reservation per non-empty bucket
        constexpr int SP_BPT = SP_MAX_BUCKETS / SP_THREADS;
        static_assert(SP_BPT * SP_THREADS == SP_MAX_BUCKETS, "whole buckets per thread");
        u32 cb[SP_BPT];
        u32 incl = 0;
#pragma unroll
        for (int q = 0; q < SP_BPT; ++q) { cb[q] = S.cnt[SP_BPT * tid + q]; incl += cb[q]; }
        const u32 mine = incl;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) S.warp_sums[warp] = incl;
        __syncthreads();
        u32 wsum = lane < SP_WARPS ? S.warp_sums[lane] : 0u;   // every warp scans the warp sums itself
#pragma unroll
        for (int d = 1; d < SP_WARPS; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, wsum, d);
            if (lane >= d) wsum += t;
        }
        const u32 total = __shfl_sync(0xffffffffu, wsum, SP_WARPS - 1);
        const u32 wex = warp ? __shfl_sync(0xffffffffu, wsum, warp - 1) : 0u;
        u32 ex = wex + incl - mine;
#pragma unroll
        for (int q = 0; q < SP_BPT; ++q) {
            if (cb[q]) {   // b < n_buckets because only such buckets were counted
                const u32 b = SP_BPT * tid + q;
                S.start[b] = ex;
                S.gbase[b] = atomicAdd(a.cursor + (size_t)b * CURSOR_PAD, (u64)cb[q]);
                S.glimit[b] = a.limit[b];
                S.cnt[b] = 0;
                ex += cb[q];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            if ((emit >> j) & 1u) {
                const u32 pos = S.start[bk[j]] + rk[j];
#pragma unroll
                for (int i = 0; i < KW; ++i) S.keys[pos * KW + i] = key[j][i];
                S.meta[pos] = (unsigned short)msk[j];
                S.bkt[pos] = (unsigned short)bk[j];
            }
        }
        __syncthreads();
        // copy-out: consecutive stage positions of a bucket go to consecutive records of its owner's area
        for (u32 i = tid; i < total; i += SP_THREADS) {
            const u32 b = S.bkt[i];
            const u64 g = S.gbase[b] + (i - S.start[b]);
            if (g >= S.glimit[b]) { a.ctr->split_overflow = 1ull; continue; }   // estimated room exceeded: the host redoes the chunk
            const u32 o = a.n_ranks > 1 ? b / a.n_regions : 0u;
            u64* kd = a.owner_keys[o] + g * KW;
            if constexpr (KW == 2) {
                *reinterpret_cast<ulonglong2*>(kd) = *reinterpret_cast<const ulonglong2*>(&S.keys[i * 2]);
            } else {
#pragma unroll
                for (int w = 0; w < KW; ++w) kd[w] = S.keys[i * KW + w];
            }
            a.owner_meta[o][g] = S.meta[i];
        }
    }
}

// Debug aid (GENOMIX_GB_DEBUG=1): number of records inside the buckets' ranges that do not belong to the bucket their hash names.
template <int KW>
__global__ void __launch_bounds__(256) check_arena_kernel(const u64* __restrict__ keys, const u64* __restrict__ tab, u32 n_ranks,
                                                          u32 n_regions, u64* __restrict__ bad) {
    const u32 n_buckets = n_ranks * n_regions;
    for (u32 b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        const u64* t = tab + (size_t)(b / n_regions) * (2 * n_regions + 1);
        const u64 lo = t[b % n_regions], n = t[n_regions + 1 + b % n_regions];
        for (u64 i = threadIdx.x; i < n; i += blockDim.x) {
            u64 key[KW];
#pragma unroll
            for (int j = 0; j < KW; ++j) key[j] = keys[(lo + i) * KW + j];
            if (bucket_of_hash(hash_key<KW>(key), n_ranks, n_regions) != b) atomicAdd(bad, 1ull);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2. Region-by-region upsert of bucket-sorted record areas.
//
// Work item = UP_ITEM consecutive records of one (region, source) pair, handled by ONE warp on its own (no CTA barriers:
// the 16-32 warps of an SM sit in different stages, which is what hides the L2 latency). Items are dealt round-robin, so
// at any time all warps of the GPU work inside the same region or two of the table. Per item:
//   stage 1  every lane loads its 4 records and their 4 home slots (all loads in flight together);
//   stage 2  from the slot snapshots: key present -> fold (fire-and-forget RED), slot empty -> claim it (CAS; all of a
//            lane's claims in flight together), slot taken by another key -> slow queue;
//   stage 3  claims that succeeded get their value folded in, claims lost to another key -> slow queue;
//   slow     the few queued records (about 6 % at load 0.4) are compacted into the low lanes and walk their probe
//            sequences with the general table_upsert. Without the queue every warp would wait for the longest probe
//            sequence among its 32 lanes, four times per item (measured: 3.7 dependent L2 round trips per record).
// Measured on cfg2 (profiles/r02_tune_upsert.txt): 256 x 2 CTAs/SM without spills beats 256 x 3 with spills (a spilled
// register makes the warp wait for the load that fills it); L2 prefetches of the next table region and of the next
// ticket's records did not pay (the region is L2-resident after its first touches either way).
#ifndef GX_UP_KW1_THREADS
#define GX_UP_KW1_THREADS 256
#define GX_UP_KW1_BLOCKS 2
#endif
// CTA shape per key width: what the register budget of the staged fast path allows
template <int KW> struct UpsertCfg {
    static constexpr int THREADS = (KW == 1) ? GX_UP_KW1_THREADS : 512;
    static constexpr int MIN_BLOCKS = (KW == 1) ? GX_UP_KW1_BLOCKS : 1;
    static constexpr int WARPS = THREADS / 32;
};
static constexpr int UP_PER_LANE = 4;
static constexpr int UP_ITEM = 32 * UP_PER_LANE;             // records per work item
#ifndef GX_UP_TICKET
#define GX_UP_TICKET 4
#endif
static constexpr u32 UP_TICKET = GX_UP_TICKET;               // consecutive work items per ticket
static constexpr int UP_PUBLISH = 512;                       // a warp publishes its new-key count at the latest after this many
static constexpr int UP_MAX_SRC = 64;                        // record areas walked by one launch (own + received from peers)

struct UpsertSrc {
    const u64* keys; const unsigned short* meta;   // the owner's record area (block start)
    const u64* seg;        // [n_regions + 1] record index of each region's first record (and the end of the block)
    const u64* cnt;        // [n_regions] records of each region (its room in the block may be larger)
    u64 rebase;            // 1: indices in seg count from seg[0] (a block cut out of a larger arena), 0: from the area start
};

struct UpsertArgs {
    UpsertSrc src[UP_MAX_SRC];
    u32 n_src;
    u32 r0, r1;            // regions [r0, r1) of this launch, walked in order; pair index i = (region - r0) * n_src + src
    u32 n_regions;
    u32 n_ranks;
    u32 active_warps;      // warps per CTA that take work (fewer for tiny tables, so that the in-flight margin stays small)
    u64* table; u64 capacity;
    u64 hash_mul;              // home slot = slot_of(hash * hash_mul, capacity): n_ranks, or n_ranks * n_regions (pilot table)
    Counters* ctr;
    const u32* item_prefix;    // [(r1 - r0) * n_src + 1] exclusive prefix of work items per pair
    u64 hard_limit;            // work items are deferred (not applied) once ctr->distinct exceeds this: with the margin the
                               // host leaves for everything in flight, the table as a whole can never fill up
    u32* region_new;           // [n_regions] keys this chunk has added to each table region so far
    u64 region_room;           // work items of a region are deferred once region_new[region], plus half of what the warps in
                               // flight are adding at this warp's current rate, exceeds this
    u32 sample_shift;          // region_new is fed by every 2^sample_shift-th work item, its new keys scaled up accordingly:
                               // one hot counter per region cannot take an atomic from every item of every warp
    u32* deferred_out;         // deferred work item ids -> ctr->deferred_count
    const u32* deferred_in;    // != nullptr: apply exactly these n_deferred_in work items
    u32 n_deferred_in;
};

// exclusive prefix of work items per (region, source) pair; one CTA
static __global__ void __launch_bounds__(1024) upsert_prefix_kernel(UpsertArgs a, u32* __restrict__ item_prefix) {
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const u32 n_pairs = (a.r1 - a.r0) * a.n_src;
    for (u32 base = 0; base < n_pairs; base += 1024) {
        const u32 i = base + threadIdx.x;
        u64 items = 0;
        if (i < n_pairs) {
            const u32 region = a.r0 + i / a.n_src, s = i % a.n_src;
            items = (a.src[s].cnt[region] + UP_ITEM - 1) / UP_ITEM;
        }
        u64 tot;
        const u64 ex = block_scan_excl<1024>(items, &tot);
        const u64 carry = carry_s;
        if (i < n_pairs) item_prefix[i] = (u32)(carry + ex);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) item_prefix[n_pairs] = (u32)carry_s;
}

static constexpr int UP_MAX_PAIRS = 2048;   // (region, source) pairs of one launch: their descriptors live in shared memory

template <int KW>
struct UpsertSmem {
    u64 qkeys[UpsertCfg<KW>::WARPS][UP_ITEM * KW];       // per-warp slow queue
    u64 first[UP_MAX_PAIRS];                 // record index (in the pair's source area) of the pair's first record
    u32 prefix[UP_MAX_PAIRS + 1];            // exclusive prefix of work items per pair
    u32 count[UP_MAX_PAIRS];                 // records of the pair
    unsigned short qmeta[UpsertCfg<KW>::WARPS][UP_ITEM];
};

template <int KW>
__global__ void __launch_bounds__(UpsertCfg<KW>::THREADS, UpsertCfg<KW>::MIN_BLOCKS) upsert_regions_kernel(UpsertArgs a) {
    constexpr int UP_THREADS = UpsertCfg<KW>::THREADS;
    constexpr int SW = SlotTraits<KW>::WORDS;
    constexpr bool PREFETCH = KW <= 2;   // the next item's records are loaded while the current one is applied
    extern __shared__ __align__(16) unsigned char upsert_smem_raw[];
    UpsertSmem<KW>& Q = *reinterpret_cast<UpsertSmem<KW>*>(upsert_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n_pairs = (a.r1 - a.r0) * a.n_src;
    for (u32 p = threadIdx.x; p <= n_pairs; p += UP_THREADS) {
        Q.prefix[p] = a.item_prefix[p];
        if (p < n_pairs) {
            const UpsertSrc& S = a.src[p % a.n_src];
            const u32 region = a.r0 + p / a.n_src;
            Q.first[p] = S.seg[region] - (S.rebase ? S.seg[0] : 0ull);
            Q.count[p] = (u32)S.cnt[region];
        }
    }
    __syncthreads();
    if ((u32)warp >= a.active_warps) return;
    const u32 lane_lt = (1u << lane) - 1u;
    u64* qk = Q.qkeys[warp];
    unsigned short* qm = Q.qmeta[warp];
    const u32 total = a.deferred_in ? a.n_deferred_in : Q.prefix[n_pairs];
    const u32 n_gw = gridDim.x * a.active_warps;
    // Work items are handed out by a ticket counter, one ticket ahead of use. A static round-robin deal lets the warps
    // drift apart (an SM far from the L2 slices it talks to is several per cent slower, which adds up to dozens of items
    // over a launch), and then the table regions in flight no longer fit in L2. With tickets all warps of the GPU work
    // inside a window of about n_gw consecutive items: one region, or the seam between two.
    // A ticket is worth UP_TICKET consecutive items: the counter is ONE address, and the L2 serialises the atomics on it
    // (a ticket per item made the counter the pace-maker of the whole kernel).
    auto take = [&]() {   // lane 0's copy is the ticket; it is broadcast where it is used, UP_TICKET items later
        u32 t = 0;
        if (lane == 0) t = (u32)atomicAdd(&a.ctr->upsert_ticket, 1ull) * UP_TICKET;
        return t;
    };
    u32 pair = 0;
    u32 unpublished = 0;                                 // slots this warp created and has not added to ctr->distinct yet
    u32 last_new = 0;                                    // new keys of this warp's previous item
    u64 next_distinct = ld_relaxed(&a.ctr->distinct);    // refreshed one item ahead, so the load never stalls the warp

    // work item w -> its id t, its pair, its records and how many there are (n == 0: no such item)
    u64 nkey[UP_PER_LANE][KW];
    u32 nm[UP_PER_LANE];
    u32 nn = 0, nt = 0, npair = 0;
    auto fetch = [&](u32 w) {
        nn = 0;
        if (w >= total) return;
        u32 t = w;
        if (a.deferred_in) {
            t = a.deferred_in[w];
            u32 lo = 0, hi = n_pairs - 1;   // largest pair with prefix[pair] <= t
            while (lo < hi) {
                const u32 mid = (lo + hi + 1) >> 1;
                if (Q.prefix[mid] <= t) lo = mid; else hi = mid - 1;
            }
            pair = lo;
        }
        while (t >= Q.prefix[pair + 1]) ++pair;   // skips empty pairs
        const u32 item_off = (t - Q.prefix[pair]) * UP_ITEM;
        const u64 first = Q.first[pair] + item_off;
        const UpsertSrc& S = a.src[pair % a.n_src];
        const u64* keys = S.keys + first * KW;
        const unsigned short* meta = S.meta + first;
        nt = t;
        npair = pair;
        nn = min((u32)UP_ITEM, Q.count[pair] - item_off);
#pragma unroll
        for (int i = 0; i < UP_PER_LANE; ++i) {
            const u32 r = lane + 32u * i;
            if (r < nn) {
#pragma unroll
                for (int j = 0; j < KW; ++j) nkey[i][j] = __ldcs(keys + (u64)r * KW + j);   // streamed once: evict first
                nm[i] = __ldcs(meta + r);
            }
        }
    };
    u32 w = __shfl_sync(0xffffffffu, take(), 0);
    fetch(w);
    u32 ticket = take(), w_next = 0;
    for (; w < total; w = w_next) {
        if ((w + 1u) % UP_TICKET) {
            w_next = w + 1u;                                // next item of this ticket
        } else {
            w_next = __shfl_sync(0xffffffffu, ticket, 0);   // taken UP_TICKET items ago
            ticket = take();                                // for the ticket after the next
        }
        u64 key[UP_PER_LANE][KW];
        u32 m[UP_PER_LANE];
        const u32 n = nn, t = nt;
        u32* fill_p = a.region_new + (a.r0 + npair / a.n_src);   // keys this chunk has added to the item's table region
#pragma unroll
        for (int i = 0; i < UP_PER_LANE; ++i) {
            m[i] = nm[i];
#pragma unroll
            for (int j = 0; j < KW; ++j) key[i][j] = nkey[i][j];
        }
        if constexpr (PREFETCH) fetch(w_next);
        // The item is applied only while the table has room. (1) Hard guarantee: ctr->distinct plus everything all warps in
        // flight can still add (the margin in hard_limit) stays below the table's capacity. (2) Per region: a region is a
        // contiguous slot range that this chunk fills while the regions after it still wait, so the load limit has to
        // hold per region, not on average; region_new is a sampled count, the warps in flight are assumed to find new
        // keys at this warp's current rate. The counter loads travel with the slot loads below.
        const u64 distinct_now = next_distinct;
        next_distinct = ld_relaxed(&a.ctr->distinct);
        const u32 fill = ld_relaxed_u32(fill_p);

        // ---- stage 1: home-slot snapshots of the item's records
        u64* sp[UP_PER_LANE];
        Probe<KW> pr[UP_PER_LANE];
#pragma unroll
        for (int i = 0; i < UP_PER_LANE; ++i) {
            if (lane + 32u * i < n) {
                sp[i] = a.table + slot_of(local_hash(hash_key<KW>(key[i]), a.hash_mul), a.capacity) * SW;
                probe_load<KW>(sp[i], pr[i]);
            }
        }
        if (distinct_now > a.hard_limit || (u64)fill + (((u64)last_new * n_gw) >> 1) > a.region_room) {
            // full: hand the item back to the host, which grows the table
            if (lane == 0) a.deferred_out[atomicAdd(&a.ctr->deferred_count, 1ull)] = t;
            if constexpr (!PREFETCH) fetch(w_next);
            continue;
        }
        // ---- stage 2: fold what is there, claim what is empty
        u32 slow = 0, claim = 0, n_new = 0;   // bit i: record i of this lane
        u64 old0[UP_PER_LANE], old1[UP_PER_LANE];
        if constexpr (KW <= 2) {
#pragma unroll
            for (int i = 0; i < UP_PER_LANE; ++i) {
                if (lane + 32u * i >= n) continue;
                bool eq = true, empty = true;
#pragma unroll
                for (int j = 0; j < KW; ++j) { eq = eq && pr[i].w[j] == key[i][j]; empty = empty && pr[i].w[j] == EMPTY_WORD; }
                if (eq) fold_value(sp[i] + KW, 1ull, m[i], pr[i].w[KW]);
                else if (empty) {
                    claim |= 1u << i;
                    // KW == 1: key and value word share one 16-byte slot, so ONE 128-bit CAS claims the slot and stores the
                    // first occurrence's count and edge bits (an empty slot's value word is 0): 1 atomic instead of 3
                    if constexpr (KW == 1) (void)cas128(sp[i], EMPTY_WORD, 0ull, key[i][0], 1ull | ((u64)m[i] << MASK_SHIFT), old0[i], old1[i]);
                    else (void)cas128(sp[i], EMPTY_WORD, EMPTY_WORD, key[i][0], key[i][1], old0[i], old1[i]);
                } else slow |= 1u << i;
            }
            // ---- stage 3: outcome of the claims
#pragma unroll
            for (int i = 0; i < UP_PER_LANE; ++i) {
                if (!((claim >> i) & 1u)) continue;
                if constexpr (KW == 1) {
                    if (old0[i] == EMPTY_WORD) ++n_new;                                                 // won: nothing left to do
                    else if (old0[i] == key[i][0]) fold_value(sp[i] + 1, 1ull, m[i], old1[i]);         // lost to the same key
                    else slow |= 1u << i;
                } else {
                    const bool won = old0[i] == EMPTY_WORD && old1[i] == EMPTY_WORD;
                    const bool same = old0[i] == key[i][0] && old1[i] == key[i][1];   // lost to the same key
                    if (won || same) { fold_value(sp[i] + KW, 1ull, m[i], 0ull); n_new += won ? 1u : 0u; }
                    else slow |= 1u << i;
                }
            }
        } else {
            // claim protocol: val 0 -> LOCK, write key words, publish val with release order
            u64 kw0[UP_PER_LANE][KW];
#pragma unroll
            for (int i = 0; i < UP_PER_LANE; ++i) {
                if (lane + 32u * i >= n) continue;
                const u64 v = pr[i].w[0];
                if (v == 0) { claim |= 1u << i; old0[i] = atomicCAS(sp[i] + KW, 0ull, VAL_LOCK); }
                else if (v == VAL_LOCK) slow |= 1u << i;
                else {
#pragma unroll
                    for (int j = 0; j < KW; ++j) kw0[i][j] = ld_relaxed(sp[i] + j);   // ordered after the acquire load of val
                }
            }
#pragma unroll
            for (int i = 0; i < UP_PER_LANE; ++i) {
                if (lane + 32u * i >= n || ((slow >> i) & 1u)) continue;
                if ((claim >> i) & 1u) {
                    if (old0[i] == 0ull) {
#pragma unroll
                        for (int j = 0; j < KW; ++j) st_relaxed(sp[i] + j, key[i][j]);
                        st_release(sp[i] + KW, 1ull | ((u64)m[i] << MASK_SHIFT));
                        ++n_new;
                    } else slow |= 1u << i;   // somebody else is claiming this slot: look again on the slow path
                } else {
                    bool eq = true;
#pragma unroll
                    for (int j = 0; j < KW; ++j) eq = eq && kw0[i][j] == key[i][j];
                    if (eq) fold_value(sp[i] + KW, 1ull, m[i], pr[i].w[0]);
                    else slow |= 1u << i;
                }
            }
        }
        // ---- slow path: compact the leftovers into the low lanes, walk their probe sequences
        if (__any_sync(0xffffffffu, slow != 0)) {
            u32 qn = 0;
#pragma unroll
            for (int i = 0; i < UP_PER_LANE; ++i) {
                const bool mine = (slow >> i) & 1u;
                const u32 bal = __ballot_sync(0xffffffffu, mine);
                if (mine) {
                    const u32 pos = qn + __popc(bal & lane_lt);
#pragma unroll
                    for (int j = 0; j < KW; ++j) qk[pos * KW + j] = key[i][j];
                    qm[pos] = (unsigned short)m[i];
                }
                qn += __popc(bal);
            }
            __syncwarp();
            for (u32 j = lane; j < qn; j += 32) {
                u64 k2[KW];
#pragma unroll
                for (int x = 0; x < KW; ++x) k2[x] = qk[j * KW + x];
                const u32 m2 = qm[j];
                bool is_new;
                u64 at;
#ifndef GX_NO_WIDE
                if constexpr (KW <= 2) at = table_upsert_wide<KW>(a.table, a.capacity, local_hash(hash_key<KW>(k2), a.hash_mul), k2, 1ull, m2, is_new);
                else
#endif
                at = table_upsert<KW>(a.table, a.capacity, local_hash(hash_key<KW>(k2), a.hash_mul), k2, 1ull, m2, is_new);
                if (at == a.capacity) spill_record<KW>(a.ctr, k2, m2);
                n_new += is_new ? 1u : 0u;
            }
            __syncwarp();
        }
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) n_new += __shfl_xor_sync(0xffffffffu, n_new, dlt);
        if (lane == 0 && n_new && (t & ((1u << a.sample_shift) - 1u)) == 0) atomicAdd(fill_p, n_new << a.sample_shift);
        last_new = n_new;
        unpublished += n_new;
        if (unpublished >= UP_PUBLISH) {
            if (lane == 0) atomicAdd(&a.ctr->distinct, (u64)unpublished);
            unpublished = 0;
        }
        if constexpr (!PREFETCH) fetch(w_next);
    }
    if (lane == 0 && unpublished) atomicAdd(&a.ctr->distinct, (u64)unpublished);
}

}  // namespace gx
