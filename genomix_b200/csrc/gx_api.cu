// gx_api.cu -- context, memory management and the C ABI of libgenomix_gb (include/genomix_gb.h).
//
// Host-side shape of one job (single GPU):
//   gx_push_lines*  : per chunk   line index -> parse -> split by table region -> region upserts [grow on deferral]  (K1+K2)
//   gx_push_records : serialised Nodes -> (key, mask, count) + heads -> upsert                                       (merge)
//   gx_finish       : read-head groups, one-pass size scan + dense node list, record writer                          (K3)
//   gx_next_*       : stream the record bytes / Hyracks frames back to the caller (serialised on demand when streaming)
// which replaces the six-operator Hyracks job of JobGenBuildBrujinGraph.assignJob
// (genomix-hyracks/.../graph/job/JobGenBuildBrujinGraph.java:79-90).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/genomix_gb.h"
#include "gx_engine.cuh"

using namespace gx;

namespace gx {
const EngineOps* engine_ops(int kw) {
    switch (kw) {
        case 1: return engine_ops_kw1();
        case 2: return engine_ops_kw2();
        case 3: return engine_ops_kw3();
        case 4: return engine_ops_kw4();
    }
    return nullptr;
}
}  // namespace gx

namespace {

constexpr double MAX_LOAD = 0.80;     // load limit of a table region while a chunk is upserted (per-region deferral)
constexpr double HARD_LOAD = 0.90;    // the table as a whole never holds more keys than this fraction of its capacity
constexpr double TARGET_LOAD = 0.60;  // capacity chosen for this load when the number of keys is known or estimated
constexpr u64 MIN_CAPACITY = 1ull << 20;
constexpr size_t DEFAULT_CHUNK = 256ull << 20;
constexpr size_t REGION_BYTES = 32ull << 20;   // table region kept L2-resident while it is being upserted

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// One chunk's k-mer records, sorted by bucket = (owner rank, table region); see gx_split.cuh.
struct Arena {
    DevBuf keys, meta;          // records: KW key words, 16-bit edge mask; bucket b's records sit at the start of its room
    DevBuf tab;                 // u64 [n_ranks][2 R + 1]: per owner the R region starts, the end of its block, the R region counts
    DevBuf owner_dev;           // u64 [n_ranks + 1] start of every owner's block (device copy of owner_off)
    DevBuf cursor, limit;       // u64 [n_buckets * CURSOR_PAD] placement cursors, [n_buckets] end of each bucket's room
    DevBuf bucket_count;        // u64 [n_buckets] (sampled) counts
    u64 occ = 0;                // records
    u64 cap = 0;                // record capacity of keys / meta
    u32 n_regions = 0;
    std::vector<u64> owner_off; // host copy of owner_dev (multi-GPU only)
};

enum Phase { PH_PARSE = 0, PH_INSERT = 1, PH_EXCHANGE = 2, PH_FINISH = 3, PH_H2D = 4, PH_XCOMM = 5, PH_XINSERT = 6, PH_SPLIT = 7, PH_COUNT = 8 };

struct PendingTimer {
    int phase;
    cudaEvent_t a, b;
};

}  // namespace

struct gx_ctx {
    gx_config cfg{};
    int k = 0, kw = 0;
    const EngineOps* ops = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int sticky = 0;

    Counters* d_ctr = nullptr;
    Counters* h_ctr = nullptr;  // pinned

    u64* table = nullptr;
    u64 capacity = 0;
    u64* kept_table = nullptr;     // the job table's allocation while the pilot stands in for it (upsert_sources)
    u64* pilot = nullptr;          // pilot table of a job without capacity hint (upsert_sources); kept for the next job
    u64 pilot_capacity = 0;
    bool table_live = false;  // false: allocation kept from before gx_reset, content stale
    u64 grows = 0;
    u64 min_capacity = MIN_CAPACITY;
    u64 hash_mul = 1;              // table mapping: home slot = slot_of(hash * hash_mul, capacity); n_ranks except in the pilot
    u64 chunk_distinct0 = 0;       // keys in the table when the current chunk's upserts began
    double new_key_rate = 1.0;     // new keys per record of the previous chunk
    bool test_start_small = false; // test hook: no hint, no pilot -> the table starts at min_capacity and grows by deferral

    DevBuf heads, store;
    DevBuf text, text2, nl_pos, nl_pos2, desc, tile_sums;
    DevBuf ht_key, ht_count, ht_start, hgroup, hentry, hperm, hoff, bkey, big_list;   // read-head groups (gx_finish)
    DevBuf tile_state, records, rec_offsets, parts, dense, dense_h, big_tiles;
    DevBuf sort_perm[2], sort_hist, sort_sizes, dense2, dense_h2;   // gx_config.sort_output
    // streaming delivery of the record stream (gx_config.reserved[2] bit 0)
    bool stream_records = false;
    size_t slice_bytes = 64ull << 20;
    DevBuf ring[2], slice_idx;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_written[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_text_ready[2] = {nullptr, nullptr}, ev_text_free[2] = {nullptr, nullptr};   // double-buffered gx_push_lines
    u64 stream_next_node = 0, stream_next_byte = 0;
    DevBuf gstats;
    EmitArgs last_emit{};      // arguments of the last emit (dense node list etc.), reused by gx_graph_statistics

    // build
    Arena arena;               // single GPU: the current chunk's records (reused chunk after chunk)
    u32 fixed_regions = 0;     // != 0: table regions per rank for the whole job (multi-GPU, or cfg.reserved[3])
    DevBuf tile_prefix, deferred[2], region_new;
    DevBuf mrec, moff, mbase, mkeys, mmeta, mcounts;   // gx_push_records staging
    u64 merged_records = 0;
    u64 split_redos = 0;           // chunks whose estimated bucket room overflowed and were split again with exact counts
    u64 upserted_records = 0;

    u64 global_lines = 0;
    u64 n_nodes = 0, record_bytes = 0;
    bool finished = false;
    size_t chunk_bytes = DEFAULT_CHUNK;

    // host slab cache for gx_next_frame
    std::vector<uint8_t> slab;
    u64 slab_off = 0, slab_len = 0;
    u64 frame_byte_cursor = 0, frame_rec_cursor = 0;

    void* mg = nullptr;  // MgState (gx_mg.inl) when n_ranks > 1

    // spill area (two buffers, swapped while one is being re-inserted)
    DevBuf spill_keys[2], spill_meta[2], spill_counts[2];
    int spill_cur = 0;
    u64 spill_cap = 0;

    float phase_ms[PH_COUNT] = {0};
    std::vector<PendingTimer> timers;
    std::vector<cudaEvent_t> event_pool;
    u64 launches = 0;
    u64 launches_at_sync = ~0ull;
};

namespace {

thread_local std::string g_create_error;

// multi-GPU hooks, defined in gx_mg.inl
int mg_stage_chunk(gx_ctx* c, const uint8_t* d_text, size_t n, u64 n_lines, u64 chunk_occ);
int mg_pending(gx_ctx* c, u64* pending);
int mg_complete(gx_ctx* c);
int mg_reset(gx_ctx* c);
void mg_destroy(gx_ctx* c);
u64 mg_exchanged(gx_ctx* c);

int fail(gx_ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CUDA_TRY(c, expr)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(c, e__ == cudaErrorMemoryAllocation ? GX_ERR_NOMEM : GX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                              \
    } while (0)

#define GX_TRY(expr)            \
    do {                        \
        int r__ = (expr);       \
        if (r__ != GX_OK) return r__; \
    } while (0)

// grow-only device buffer; keep_bytes of old content survive a reallocation
int ensure(gx_ctx* c, DevBuf& b, size_t bytes, size_t keep_bytes = 0, bool zero_new = false) {
    if (bytes <= b.cap) return GX_OK;
    size_t ncap = std::max(bytes, b.cap + b.cap / 2);
    ncap = (ncap + 255) & ~(size_t)255;
    void* np = nullptr;
    CUDA_TRY(c, cudaMalloc(&np, ncap + 64));  // +64: kernels may read whole aligned words past the end
    if (keep_bytes) CUDA_TRY(c, cudaMemcpyAsync(np, b.p, keep_bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (zero_new) CUDA_TRY(c, cudaMemsetAsync((char*)np + keep_bytes, 0, ncap + 64 - keep_bytes, c->stream));
    if (b.p) {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaFree(b.p));
    }
    b.p = np;
    b.cap = ncap;
    return GX_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

void release_arena(Arena& a) {
    release(a.keys); release(a.meta); release(a.tab); release(a.owner_dev); release(a.cursor); release(a.limit); release(a.bucket_count);
}

cudaEvent_t get_event(gx_ctx* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct ScopedPhase {
    gx_ctx* c;
    PendingTimer t;
    ScopedPhase(gx_ctx* ctx, int phase) : c(ctx) {
        t.phase = phase;
        t.a = get_event(c);
        t.b = get_event(c);
        cudaEventRecord(t.a, c->stream);
    }
    ~ScopedPhase() {
        cudaEventRecord(t.b, c->stream);
        c->timers.push_back(t);
    }
};

// resolve finished timers (call after a stream synchronise); timers of work still running on another stream stay pending
void drain_timers(gx_ctx* c) {
    size_t kept = 0;
    for (auto& t : c->timers) {
        if (cudaEventQuery(t.b) == cudaErrorNotReady) { c->timers[kept++] = t; continue; }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) c->phase_ms[t.phase] += ms;
        c->event_pool.push_back(t.a);
        c->event_pool.push_back(t.b);
    }
    c->timers.resize(kept);
    cudaGetLastError();   // cudaErrorNotReady is not an error here
}

int sync_counters(gx_ctx* c) {
    CUDA_TRY(c, cudaMemcpyAsync(c->h_ctr, c->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    c->launches_at_sync = c->launches;
    return GX_OK;
}

// the host copy of the counters is current if no kernel was launched since it was taken
int sync_counters_if_stale(gx_ctx* c) { return c->launches == c->launches_at_sync ? GX_OK : sync_counters(c); }

int check_launch(gx_ctx* c, const char* what) {
    ++c->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, GX_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return GX_OK;
}

int line_error_to_status(gx_ctx* c, u64 packed) {
    const u64 line = packed >> 8;
    const u32 code = (u32)(packed & 0xff);
    switch (code) {
        case LE_FORMAT:
            c->sticky = fail(c, GX_ERR_FORMAT,
                             "IllegalStateException: input format is not correct! only support id'\\t'readSeq'\\t'mateReadSeq or "
                             "id'\\t'readSeq' (input line %llu)", (unsigned long long)line);
            break;
        case LE_NUMBER:
            c->sticky = fail(c, GX_ERR_NUMBER, "NumberFormatException: read id is not a long (input line %llu)",
                             (unsigned long long)line);
            break;
        case LE_TOO_SHORT:
            c->sticky = fail(c, GX_ERR_READ_TOO_SHORT,
                             "IllegalArgumentException: kmersize (k=%d) is larger than the read length (input line %llu)", c->k,
                             (unsigned long long)line);
            break;
        case LE_READID:
            c->sticky = fail(c, GX_ERR_READID_RANGE,
                             "IllegalArgumentException: byte specified for readId will lose some of its bits when saved! "
                             "(input line %llu)", (unsigned long long)line);
            break;
        case LE_RECORD:
            c->sticky = fail(c, GX_ERR_FORMAT, "malformed or unsupported Node record (record %llu of the pushed stream): gx_push_records takes "
                             "graph-build records of this kmer length", (unsigned long long)line);
            break;
        default:
            c->sticky = fail(c, GX_ERR_INVALID, "unknown line error %u (input line %llu)", code, (unsigned long long)line);
    }
    return c->sticky;
}

// ---- table life cycle ---------------------------------------------------------------------------------------
// Capacity policy: see upsert_sources (sizing) and run_upsert_range (per-region load limit, deferral, growth).
int alloc_table(gx_ctx* c, u64 capacity, u64** out) {
    void* p = nullptr;
    const size_t bytes = (size_t)capacity * c->ops->slot_bytes;
    CUDA_TRY(c, cudaMalloc(&p, bytes));
    c->ops->init_table((u64*)p, capacity, c->stream);
    GX_TRY(check_launch(c, "init_table"));
    *out = (u64*)p;
    return GX_OK;
}

int ensure_table(gx_ctx* c, u64 min_capacity) {
    if (c->table && !c->table_live) {
        if (c->capacity >= min_capacity) {
            // allocation kept across gx_reset: re-initialise and reuse
            c->ops->init_table(c->table, c->capacity, c->stream);
            GX_TRY(check_launch(c, "init_table"));
            c->table_live = true;
            return GX_OK;
        }
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaFree(c->table));
        c->table = nullptr;
        c->capacity = 0;
    }
    if (!c->table) {
        GX_TRY(alloc_table(c, min_capacity, &c->table));
        c->capacity = min_capacity;
        c->table_live = true;
    }
    return GX_OK;
}

// rehash into a table of `ncap` slots (ncap > capacity)
int grow_table_to(gx_ctx* c, u64 ncap) {
    u64* nt = nullptr;
    GX_TRY(alloc_table(c, ncap, &nt));
    c->ops->rehash(c->table, c->capacity, nt, ncap, c->hash_mul, c->d_ctr, c->stream);
    GX_TRY(check_launch(c, "rehash"));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaFree(c->table));
    ++c->grows;
    c->table = nt;
    c->capacity = ncap;
    return GX_OK;
}

int set_spill_target(gx_ctx* c) {
    struct { u64 count, cap; u64* keys; unsigned short* meta; u32* counts; } t = {
        0, c->spill_cap, (u64*)c->spill_keys[c->spill_cur].p, (unsigned short*)c->spill_meta[c->spill_cur].p,
        (u32*)c->spill_counts[c->spill_cur].p};
    static_assert(offsetof(Counters, spill_counts) - offsetof(Counters, spill_count) == 32, "spill fields are contiguous");
    CUDA_TRY(c, cudaMemcpyAsync(&c->d_ctr->spill_count, &t, sizeof t, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

// Call after sync_counters(): re-insert records whose upsert ran out of probe budget (a last line of defence: with the
// load bounded by MAX_LOAD probe sequences are short, so this only triggers on adversarial clustering).
int handle_spills(gx_ctx* c) {
    while (c->h_ctr->spill_count > 0) {
        if (c->h_ctr->table_overflow)
            return c->sticky = fail(c, GX_ERR_NOMEM, "%llu k-mer records overflowed the table and its spill area",
                                    (unsigned long long)c->h_ctr->table_overflow);
        const u64 n = c->h_ctr->spill_count;
        const int old = c->spill_cur;
        c->spill_cur ^= 1;
        GX_TRY(set_spill_target(c));   // new spills (from the re-insertion itself) go to the other buffer
        GX_TRY(grow_table_to(c, c->capacity * 2));
        c->ops->insert_records((const u64*)c->spill_keys[old].p, (const unsigned short*)c->spill_meta[old].p,
                               (const u32*)c->spill_counts[old].p, n, c->table, c->capacity, c->hash_mul, c->d_ctr, c->stream);
        GX_TRY(check_launch(c, "insert_records"));
        GX_TRY(sync_counters(c));
    }
    return GX_OK;
}

// ---- build: split + region upsert -----------------------------------------------------------------------------
// Table regions per rank for a chunk of `occ` occurrences: regions of about REGION_BYTES of the table the job is
// expected to end up with.
u32 choose_regions(gx_ctx* c, u64 occ) {
    const u32 max_regions = (u32)(SP_MAX_BUCKETS / std::max(1, c->cfg.n_ranks));
    if (c->fixed_regions) return std::min(c->fixed_regions, max_regions);
    const u64 distinct = c->table_live ? c->h_ctr->distinct : 0;
    // without a hint: sequencing data repeats itself (coverage), so about half of a chunk's occurrences being new keys is
    // already the pessimistic case; regions that turn out larger than planned cost a little L2 locality, nothing else
    const u64 keys = c->cfg.expected_kmers ? std::max<u64>(c->cfg.expected_kmers, distinct) : distinct + occ / 2;
    size_t table_bytes = (size_t)((double)keys / TARGET_LOAD) * c->ops->slot_bytes;
    if (c->table_live) table_bytes = std::max(table_bytes, (size_t)c->capacity * c->ops->slot_bytes);
    return (u32)std::min<size_t>(max_regions, std::max<size_t>(1, (table_bytes + REGION_BYTES - 1) / REGION_BYTES));
}

// K1: the chunk's k-mer records into `ar`, sorted by (owner, region). The room of every bucket comes from a count pass over
// every 16th line (exact = false) -- an estimate with slack; if a bucket still overflows, the chunk is split again with the
// exact counts of a full count pass, which also is what small chunks get. Leaves the stream synchronised.
int split_chunk_once(gx_ctx* c, const uint8_t* d_text, size_t n, u64 n_lines, u64 chunk_occ, u32 n_regions, Arena& ar, bool exact) {
    const u32 n_ranks = (u32)c->cfg.n_ranks, n_buckets = n_ranks * n_regions;
    const u32 sample = (exact || n_lines < 16384) ? 1u : 16u;
    ar.occ = chunk_occ;
    ar.n_regions = n_regions;
    ar.cap = sample == 1 ? chunk_occ : chunk_occ + chunk_occ / 8 + (u64)4096 * n_buckets;
    GX_TRY(ensure(c, ar.keys, (size_t)ar.cap * c->kw * sizeof(u64)));
    GX_TRY(ensure(c, ar.meta, (size_t)ar.cap * sizeof(unsigned short)));
    GX_TRY(ensure(c, ar.tab, (size_t)(2 * SP_MAX_BUCKETS + SP_MAX_RANKS) * sizeof(u64)));   // n_ranks * (2 R + 1), n_ranks * R <= SP_MAX_BUCKETS
    GX_TRY(ensure(c, ar.owner_dev, (size_t)(SP_MAX_RANKS + 1) * sizeof(u64)));
    GX_TRY(ensure(c, ar.cursor, (size_t)SP_MAX_BUCKETS * CURSOR_PAD * sizeof(u64)));
    GX_TRY(ensure(c, ar.limit, (size_t)SP_MAX_BUCKETS * sizeof(u64)));
    GX_TRY(ensure(c, ar.bucket_count, (size_t)SP_MAX_BUCKETS * sizeof(u64)));
    {
        ScopedPhase ph(c, PH_SPLIT);
        CUDA_TRY(c, cudaMemsetAsync(ar.bucket_count.p, 0, (size_t)n_buckets * sizeof(u64), c->stream));
        SplitArgs a{};
        a.text = d_text; a.n_text = n;
        a.desc = (const LineDesc*)c->desc.p; a.n_lines = n_lines;
        a.first_line = ((u64)c->cfg.rank << 48) | (c->global_lines - n_lines);
        a.k = c->k;
        a.heads = c->heads.p;
        a.store = (uint8_t*)c->store.p;
        a.ctr = c->d_ctr;
        a.n_ranks = n_ranks; a.n_regions = n_regions; a.sample = sample;
        a.bucket_count = (u64*)ar.bucket_count.p;
        a.cursor = (u64*)ar.cursor.p;
        a.limit = (const u64*)ar.limit.p;
        for (u32 o = 0; o < n_ranks; ++o) { a.owner_keys[o] = (u64*)ar.keys.p; a.owner_meta[o] = (unsigned short*)ar.meta.p; }
        c->ops->split_count(a, c->stream);
        GX_TRY(check_launch(c, "split_count"));
        split_prefix_kernel<<<1, 1024, 0, c->stream>>>((const u64*)ar.bucket_count.p, n_buckets, n_regions, sample, ar.cap, (u64*)ar.tab.p,
                                                       (u64*)ar.owner_dev.p, (u64*)ar.cursor.p, (u64*)ar.limit.p, c->d_ctr);
        GX_TRY(check_launch(c, "split_prefix"));
        c->ops->split_place(a, c->stream);
        GX_TRY(check_launch(c, "split_place"));
        split_finish_kernel<<<1, 1024, 0, c->stream>>>((const u64*)ar.cursor.p, (const u64*)ar.limit.p, n_buckets, n_regions, (u64*)ar.tab.p);
        GX_TRY(check_launch(c, "split_finish"));
    }
    if (getenv("GENOMIX_GB_DEBUG")) {
        CUDA_TRY(c, cudaMemsetAsync(&c->d_ctr->scratch[0], 0, sizeof(u64), c->stream));
        c->ops->check_arena((const u64*)ar.keys.p, (const u64*)ar.tab.p, n_ranks, n_regions, &c->d_ctr->scratch[0], c->stream);
        GX_TRY(sync_counters(c));
        fprintf(stderr, "[genomix_gb debug] split: %llu records, %u regions x %u ranks, counted from every %u-th line, overflow %llu, "
                "%llu records outside their bucket\n", (unsigned long long)chunk_occ, n_regions, n_ranks, sample,
                (unsigned long long)c->h_ctr->split_overflow, (unsigned long long)c->h_ctr->scratch[0]);
    }
    if (n_ranks > 1) {   // the exchange needs every owner's block boundaries on the host: they travel with this sync
        ar.owner_off.assign((size_t)n_ranks + 1, 0);
        CUDA_TRY(c, cudaMemcpyAsync(ar.owner_off.data(), ar.owner_dev.p, ((size_t)n_ranks + 1) * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    }
    return sync_counters(c);
}

int split_chunk(gx_ctx* c, const uint8_t* d_text, size_t n, u64 n_lines, u64 chunk_occ, u32 n_regions, Arena& ar) {
    const bool force_exact = getenv("GENOMIX_GB_EXACT_SPLIT") != nullptr;   // tuning / tests
    GX_TRY(split_chunk_once(c, d_text, n, n_lines, chunk_occ, n_regions, ar, force_exact));
    if (c->h_ctr->split_overflow) {
        ++c->split_redos;
        GX_TRY(split_chunk_once(c, d_text, n, n_lines, chunk_occ, n_regions, ar, true));
        if (c->h_ctr->split_overflow) return fail(c, GX_ERR_INVALID, "internal error: the exact split overflowed its arena");
    }
    return GX_OK;
}

// K2: upsert regions [r0, r1) of the given record areas; grows the table and re-launches whatever the kernel deferred.
// `map_regions` = table regions the current mapping (c->hash_mul) spreads over the table: n_regions for the job's table,
// 1 for the pilot table. Leaves the stream synchronised and c->h_ctr current.
int run_upsert_range(gx_ctx* c, const UpsertSrc* src, u32 n_src, u32 n_regions, u32 r0, u32 r1, u64 max_items, u32 map_regions);

int run_upsert(gx_ctx* c, const UpsertSrc* src, u32 n_src, u32 n_regions, u32 r0, u32 r1, u64 max_items, u32 map_regions) {
    // the kernel keeps the (region, source) pair descriptors of a launch in shared memory: at most UP_MAX_PAIRS of them
    u32 step = std::max<u32>(1, UP_MAX_PAIRS / n_src);
    if (const char* e = getenv("GENOMIX_GB_UPSERT_STEP")) step = std::max<u32>(1, std::min<u32>(step, (u32)atoi(e)));   // tuning
    for (u32 r = r0; r < r1; r += step) GX_TRY(run_upsert_range(c, src, n_src, n_regions, r, std::min(r1, r + step), max_items, map_regions));
    return GX_OK;
}

int run_upsert_range(gx_ctx* c, const UpsertSrc* src, u32 n_src, u32 n_regions, u32 r0, u32 r1, u64 max_items, u32 map_regions) {
    UpsertArgs a{};
    for (u32 s = 0; s < n_src; ++s) a.src[s] = src[s];
    a.n_src = n_src; a.r0 = r0; a.r1 = r1; a.n_regions = n_regions; a.n_ranks = (u32)c->cfg.n_ranks;
    a.ctr = c->d_ctr;
    a.region_new = (u32*)c->region_new.p;
    const size_t n_pairs = (size_t)(r1 - r0) * n_src;
    GX_TRY(ensure(c, c->tile_prefix, (n_pairs + 1) * sizeof(u32)));
    GX_TRY(ensure(c, c->deferred[0], (size_t)max_items * sizeof(u32)));
    GX_TRY(ensure(c, c->deferred[1], (size_t)max_items * sizeof(u32)));
    a.item_prefix = (const u32*)c->tile_prefix.p;
    upsert_prefix_kernel<<<1, 1024, 0, c->stream>>>(a, (u32*)c->tile_prefix.p);
    GX_TRY(check_launch(c, "upsert_prefix"));
    int cur = 0;
    u64 n_deferred = 0;
    const u64 cta_warps = (u64)c->ops->upsert_warps;
    for (;;) {
        a.table = c->table; a.capacity = c->capacity; a.hash_mul = c->hash_mul;
        // (1) Whole table: every warp in flight may add an item's worth of keys that ctr->distinct does not show yet,
        // plus what it has not published (UP_PUBLISH), plus the item it decides on with a count that is one item old:
        // that margin below HARD_LOAD makes "the table never fills up" hold by construction.
        // (2) Per table region (performance, not safety): the keys the region held when the chunk began (hash-uniform
        // share of the table's keys, plus 4 sigma) and the keys the chunk has added since (region_new, a sampled count)
        // should not exceed MAX_LOAD of the region's slots; see the kernel for how in-flight work is accounted.
        // Small tables get fewer warps so that both margins stay fractions of their limits.
        const u64 region_slots = c->capacity / map_regions;
        const u64 limit = (u64)(MAX_LOAD * (double)region_slots);
        const u64 hard = (u64)(HARD_LOAD * (double)c->capacity);
        const u64 per_warp = UP_PUBLISH + 2 * UP_ITEM;
        const double base_mean = (double)c->chunk_distinct0 / map_regions;
        const u64 base = (u64)(base_mean + 4.0 * std::sqrt(base_mean)) + (c->chunk_distinct0 ? 1 : 0);
        const u64 warps = std::max<u64>(1, std::min<u64>((u64)148 * c->ops->upsert_blocks * cta_warps,
                                                         std::min<u64>(limit / (2 * UP_ITEM), hard / (4 * per_warp))));
        const unsigned grid = (unsigned)((warps + cta_warps - 1) / cta_warps);
        a.active_warps = (u32)std::min<u64>(cta_warps, warps);
        a.hard_limit = hard - std::min<u64>(hard, (u64)grid * a.active_warps * per_warp);
        a.region_room = limit > base ? limit - base : 0;
        // regions with thousands of work items are counted through a sample of their items (1 in 16: +-3 %); the sample's
        // granularity stays below 1/16 of the limit
        a.sample_shift = 0;
        if (max_items / std::max<u32>(1, n_regions) >= 4096)
            while (a.sample_shift < 4 && ((u64)UP_ITEM << (a.sample_shift + 1)) <= limit / 16) ++a.sample_shift;
        a.deferred_out = (u32*)c->deferred[cur].p;
        a.deferred_in = n_deferred ? (const u32*)c->deferred[cur ^ 1].p : nullptr;
        a.n_deferred_in = (u32)n_deferred;
        CUDA_TRY(c, cudaMemsetAsync(&c->d_ctr->upsert_ticket, 0, sizeof(u64), c->stream));
        c->ops->upsert_regions(a, grid, c->stream);
        GX_TRY(check_launch(c, "upsert_regions"));
        GX_TRY(sync_counters(c));
        GX_TRY(handle_spills(c));
        n_deferred = c->h_ctr->deferred_count;
        if (!n_deferred) return GX_OK;
        if (getenv("GENOMIX_GB_DEBUG"))
            fprintf(stderr, "[genomix_gb debug] upsert: %llu items deferred (regions %u-%u of %u, capacity %llu, keys %llu, region room %llu, "
                    "hard limit %llu): growing\n", (unsigned long long)n_deferred, r0, r1, n_regions, (unsigned long long)c->capacity,
                    (unsigned long long)c->h_ctr->distinct, (unsigned long long)a.region_room, (unsigned long long)a.hard_limit);
        // a region reached its load limit: double the table and apply what was postponed
        GX_TRY(grow_table_to(c, c->capacity * 2));
        CUDA_TRY(c, cudaMemsetAsync(&c->d_ctr->deferred_count, 0, sizeof(u64), c->stream));
        cur ^= 1;
    }
}

// Upsert all regions of `n_src` record areas holding `total` records: capacity policy around run_upsert.
// Correctness never depends on the sizing below (per-region deferral + growth bound the load by construction); it decides
// how often the table has to be rehashed:
//   * cfg.expected_kmers, if given, sizes the first allocation for TARGET_LOAD;
//   * else the first chunk's region 0 is upserted into a small *pilot table* whose mapping spreads that one region over
//     all of its slots (hash_mul = n_ranks * n_regions). Regions are uniform hash ranges, so its key count times the
//     number of regions estimates the chunk's keys; the job's table is allocated once for that and the pilot table is
//     rehashed into it;
//   * later chunks grow the table ahead of time when the previous chunk's new-key rate says they will not fit.
int upsert_sources(gx_ctx* c, const UpsertSrc* src, u32 n_src, u32 n_regions, u64 total, int phase = PH_INSERT) {
    if (total == 0) return GX_OK;
    const u64 max_items = total / UP_ITEM + (u64)n_regions * n_src + 1;
    if (max_items >= 0xf0000000ull) return fail(c, GX_ERR_INVALID, "chunk of %llu k-mer records is too large", (unsigned long long)total);
    const u64 distinct0 = c->table_live ? c->h_ctr->distinct : 0;
    const u64 hint = c->test_start_small ? 0 : c->cfg.expected_kmers;
    const u64 n_ranks = (u64)c->cfg.n_ranks;
    c->chunk_distinct0 = distinct0;
    c->hash_mul = n_ranks;
    GX_TRY(ensure(c, c->region_new, (size_t)SP_MAX_BUCKETS * sizeof(u32)));
    CUDA_TRY(c, cudaMemsetAsync(c->region_new.p, 0, (size_t)SP_MAX_BUCKETS * sizeof(u32), c->stream));
    u32 r = 0;
    if (!c->table_live) {
        u64 cap = c->min_capacity;
        if (hint) cap = std::max<u64>(cap, (u64)((double)hint / TARGET_LOAD) + 1);
        else if (!c->test_start_small && n_regions < 8) cap = std::max<u64>(cap, (u64)((double)total / TARGET_LOAD) + 1);
        if (!hint && !c->test_start_small && n_regions >= 8) {
            // pilot: region 0 alone, in a table of its own, tells how many of this chunk's records are distinct keys
            ScopedPhase ph(c, phase);
            const u64 share = total / n_regions;
            const u64 pilot_cap = std::max<u64>(65536, (u64)((double)(share + share / 8 + 4096) / MAX_LOAD) + 1);
            u64* const kept = c->table;             // allocation kept across gx_reset (content stale), if any
            const u64 kept_cap = c->capacity;
            c->kept_table = kept;                   // owned by the ctx while c->table points at the pilot (freed by gx_destroy on a failure)
            if (c->pilot && (c->pilot_capacity < pilot_cap || c->pilot_capacity > 4 * pilot_cap)) {
                CUDA_TRY(c, cudaStreamSynchronize(c->stream));
                CUDA_TRY(c, cudaFree(c->pilot));
                c->pilot = nullptr;
            }
            if (c->pilot) {
                c->ops->init_table(c->pilot, c->pilot_capacity, c->stream);
                GX_TRY(check_launch(c, "init_table"));
            } else {
                GX_TRY(alloc_table(c, pilot_cap, &c->pilot));
                c->pilot_capacity = pilot_cap;
            }
            c->table = c->pilot; c->capacity = c->pilot_capacity; c->table_live = true;
            c->pilot = nullptr;                     // run_upsert may replace (grow) the table it works on
            c->hash_mul = n_ranks * n_regions;
            c->chunk_distinct0 = 0;
            const u64 grows0 = c->grows;
            GX_TRY(run_upsert(c, src, n_src, n_regions, 0, 1, max_items, 1));
            c->grows = grows0;
            c->pilot = c->table; c->pilot_capacity = c->capacity;
            c->table = kept; c->capacity = kept_cap; c->table_live = false;
            c->kept_table = nullptr;
            c->hash_mul = n_ranks;
            const u64 expect = (u64)((double)c->h_ctr->distinct * n_regions * 1.03) + 65536;
            const u64 want = std::max<u64>(cap, (u64)((double)expect / TARGET_LOAD) + 1);
            if (c->table && (c->capacity < want || c->capacity > 2 * want)) {   // a kept allocation of the wrong size
                CUDA_TRY(c, cudaFree(c->table));
                c->table = nullptr; c->capacity = 0;
            }
            GX_TRY(ensure_table(c, want));
            c->ops->rehash(c->pilot, c->pilot_capacity, c->table, c->capacity, c->hash_mul, c->d_ctr, c->stream);   // pilot -> job table
            GX_TRY(check_launch(c, "rehash"));
            r = 1;
        } else {
            GX_TRY(ensure_table(c, cap));
        }
    } else if ((double)(distinct0 + total) > MAX_LOAD * (double)c->capacity && !(hint && distinct0 < hint)) {
        // the chunk's worst case does not fit: make room for what the previous chunk's new-key rate predicts
        const u64 expect = distinct0 + (u64)(c->new_key_rate * 1.25 * (double)total) + 65536;
        const u64 want = (u64)((double)expect / TARGET_LOAD) + 1;
        if (want > c->capacity) GX_TRY(grow_table_to(c, want));
    }
    ScopedPhase ph(c, phase);
    GX_TRY(run_upsert(c, src, n_src, n_regions, r, n_regions, max_items, n_regions));
    c->upserted_records += total;
    c->new_key_rate = (double)(c->h_ctr->distinct - distinct0) / (double)total;
    return GX_OK;
}

// After a parse kernel filled c->desc for `n_lines` lines of the text at d_text: check errors, make room for heads and
// packed reads, then split the chunk's k-mers by table region and upsert them (single GPU) or stage them for the
// exchange (multi GPU).
int insert_parsed_chunk(gx_ctx* c, const uint8_t* d_text, size_t n) {
    GX_TRY(sync_counters(c));
    GX_TRY(handle_spills(c));
    const Counters& h = *c->h_ctr;
    const u64 n_lines = h.chunk_lines;
    c->global_lines += n_lines;
    if (h.error != ~0ull) return line_error_to_status(c, h.error);
    if (h.chunk_reads == 0) return GX_OK;
    // room for this chunk's heads and packed reads
    GX_TRY(ensure(c, c->heads, (size_t)h.head_cursor * c->ops->head_bytes,
                  (size_t)(h.head_cursor - h.chunk_reads) * c->ops->head_bytes, true));
    GX_TRY(ensure(c, c->store, (size_t)h.store_cursor, (size_t)(h.store_cursor - h.chunk_store)));
    const u64 chunk_occ = h.chunk_occ;
    if (c->cfg.n_ranks > 1) return mg_stage_chunk(c, d_text, n, n_lines, chunk_occ);
    const u32 n_regions = choose_regions(c, chunk_occ);
    GX_TRY(split_chunk(c, d_text, n, n_lines, chunk_occ, n_regions, c->arena));
    UpsertSrc src{(const u64*)c->arena.keys.p, (const unsigned short*)c->arena.meta.p, (const u64*)c->arena.tab.p,
                  (const u64*)c->arena.tab.p + n_regions + 1, 0};
    return upsert_sources(c, &src, 1, n_regions, chunk_occ);
}

// Line index of text[0, n): fills *nl_buf with the offsets of the line terminators (virtual one for an unterminated
// last line) and returns an upper bound of the line count; the exact count is left in Counters::chunk_lines.
int index_lines(gx_ctx* c, const uint8_t* d_text, size_t n, DevBuf& nl_buf, u64* max_lines) {
    const u64 n_tiles = (n + LI_TILE - 1) / LI_TILE;
    {
        ScopedPhase ph(c, PH_PARSE);
        GX_TRY(ensure(c, c->tile_sums, (n_tiles + 1) * sizeof(u64)));
        count_newlines_kernel<<<(unsigned)n_tiles, LI_THREADS, 0, c->stream>>>(d_text, n, (u64*)c->tile_sums.p);
        GX_TRY(check_launch(c, "count_newlines"));
        scan_tile_sums_kernel<<<1, 1024, 0, c->stream>>>((u64*)c->tile_sums.p, n_tiles, &c->d_ctr->scratch[0]);
        GX_TRY(check_launch(c, "scan_tile_sums"));
    }
    GX_TRY(sync_counters(c));
    *max_lines = c->h_ctr->scratch[0] + 1;
    {
        ScopedPhase ph(c, PH_PARSE);
        GX_TRY(ensure(c, nl_buf, (*max_lines + 1) * sizeof(u32)));
        write_newlines_kernel<<<(unsigned)n_tiles, LI_THREADS, 0, c->stream>>>(d_text, n, (const u64*)c->tile_sums.p, (u32*)nl_buf.p);
        GX_TRY(check_launch(c, "write_newlines"));
        finish_line_index_kernel<<<1, 1, 0, c->stream>>>(d_text, n, &c->d_ctr->scratch[0], (u32*)nl_buf.p, c->d_ctr);
        GX_TRY(check_launch(c, "finish_line_index"));
    }
    return GX_OK;
}

// One chunk of text resident in device memory: line index -> parse -> reserve -> extract+insert.
int push_chunk_device(gx_ctx* c, const uint8_t* d_text, size_t n) {
    if (n == 0) return GX_OK;
    if (n >= (1ull << 32) - 64) return fail(c, GX_ERR_INVALID, "chunk of %zu bytes exceeds the 4 GiB chunk limit", n);
    u64 max_lines = 0;
    GX_TRY(index_lines(c, d_text, n, c->nl_pos, &max_lines));
    {
        ScopedPhase ph(c, PH_PARSE);
        GX_TRY(ensure(c, c->desc, max_lines * sizeof(LineDesc)));
        parse_lines_kernel<<<(unsigned)((max_lines + PL_THREADS - 1) / PL_THREADS), PL_THREADS, 0, c->stream>>>(
            d_text, (const u32*)c->nl_pos.p, c->global_lines, c->k, (LineDesc*)c->desc.p, c->d_ctr);
        GX_TRY(check_launch(c, "parse_lines"));
    }
    return insert_parsed_chunk(c, d_text, n);
}

// One chunk of fastq (r1 at [0, n1), r2 at [base2, base2 + n2) of the device text buffer; n2 == 0: single-end).
int push_fastq_chunk_device(gx_ctx* c, const uint8_t* d_text, size_t n_total, size_t n1, size_t base2, size_t n2, u64 first_record) {
    u64 max1 = 0, max2 = 0, lines1 = 0, lines2 = 0;
    GX_TRY(index_lines(c, d_text, n1, c->nl_pos, &max1));
    GX_TRY(sync_counters(c));
    lines1 = c->h_ctr->chunk_lines;
    if (n2) {
        GX_TRY(index_lines(c, d_text + base2, n2, c->nl_pos2, &max2));
        GX_TRY(sync_counters(c));
        lines2 = c->h_ctr->chunk_lines;
        if (lines1 != lines2)  // GenomixDriver.java:684-687
            return c->sticky = fail(c, GX_ERR_FORMAT, "IOException: Fastq files didn't have the same number of lines! (%llu vs %llu)",
                                    (unsigned long long)lines1, (unsigned long long)lines2);
    }
    const u64 n_records = (lines1 + 2) / 4;
    if (n_records == 0) return GX_OK;
    {
        ScopedPhase ph(c, PH_PARSE);
        GX_TRY(ensure(c, c->desc, n_records * sizeof(LineDesc)));
        set_chunk_lines_kernel<<<1, 1, 0, c->stream>>>(c->d_ctr, n_records);
        GX_TRY(check_launch(c, "set_chunk_lines"));
        parse_fastq_kernel<<<(unsigned)((n_records + PL_THREADS - 1) / PL_THREADS), PL_THREADS, 0, c->stream>>>(
            d_text, (const u32*)c->nl_pos.p, lines1, (u32)base2, (const u32*)c->nl_pos2.p, n2 ? 1 : 0, first_record,
            c->global_lines, c->k, (LineDesc*)c->desc.p, c->d_ctr);
        GX_TRY(check_launch(c, "parse_fastq"));
    }
    return insert_parsed_chunk(c, d_text, n_total);
}

__global__ void gather_u64_kernel(const u64* __restrict__ src, const u64* __restrict__ idx, u64 n, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

int make_copy_stream(gx_ctx* c) {
    CUDA_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_written[i], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_text_ready[i], cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_text_free[i], cudaEventDisableTiming));
    }
    return GX_OK;
}

// End (exclusive) of the last whole line inside text[lo, hi): the position behind the last '\n', or behind a '\r' that is not
// followed by '\n' (a "\r\n" pair is never cut in two); ~0 if the window holds no line end. One CTA, walking backwards.
__global__ void __launch_bounds__(1024) last_line_end_kernel(const uint8_t* __restrict__ text, u64 lo, u64 hi, u64 n, u64* __restrict__ out) {
    __shared__ unsigned long long best;
    constexpr u64 PER_THREAD = 64, BLOCK = 1024 * PER_THREAD;
    for (u64 end = hi; end > lo;) {
        const u64 begin = end - lo > BLOCK ? end - BLOCK : lo;
        if (threadIdx.x == 0) best = 0;
        __syncthreads();
        const u64 t0 = begin + (u64)threadIdx.x * PER_THREAD;
        u64 mine = 0;
        for (u64 p = t0; p < min(t0 + PER_THREAD, end); ++p) {
            const uint8_t ch = text[p];
            if (ch == '\n' || (ch == '\r' && !(p + 1 < n && text[p + 1] == '\n'))) mine = p + 1;
        }
        if (mine) atomicMax(&best, (unsigned long long)mine);
        __syncthreads();
        const u64 b = best;
        __syncthreads();
        if (b) { if (threadIdx.x == 0) out[0] = b; return; }
        end = begin;
    }
    if (threadIdx.x == 0) out[0] = ~0ull;
}

int require_live(gx_ctx* c) {
    if (!c) return GX_ERR_INVALID;
    if (c->sticky) return c->sticky;
    return GX_OK;
}

// single-thread binary search: largest index i in [lo, n] with offsets[i] <= limit
__global__ void upper_bound_kernel(const u64* __restrict__ offsets, u64 lo, u64 n, u64 limit, u64* __restrict__ out) {
    u64 a = lo, b = n;  // invariant offsets[a] <= limit
    while (a < b) {
        const u64 mid = a + (b - a + 1) / 2;
        if (offsets[mid] <= limit) a = mid; else b = mid - 1;
    }
    out[0] = a;
    out[1] = offsets[a];
}

}  // namespace

// ================================================================================================
extern "C" {

int gx_abi_version(void) { return GX_ABI_VERSION; }

const char* gx_last_error(const gx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gx_create(const gx_config* cfg, gx_ctx** out) {
    if (!cfg || !out) return fail(nullptr, GX_ERR_INVALID, "gx_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != GX_ABI_VERSION)
        return fail(nullptr, GX_ERR_INVALID, "gx_create: abi_version %d != %d", cfg->abi_version, GX_ABI_VERSION);
    if (cfg->kmer_length < 1 || cfg->kmer_length > 32 * GX_MAX_KW)
        return fail(nullptr, GX_ERR_INVALID, "gx_create: kmer_length %d outside [1, %d]", cfg->kmer_length, 32 * GX_MAX_KW);
    if (cfg->sort_output != 0 && cfg->sort_output != 1)
        return fail(nullptr, GX_ERR_INVALID, "gx_create: sort_output must be 0 (table-slot order) or 1 (KmerPointable order)");
    if (cfg->n_ranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->n_ranks || cfg->n_ranks > SP_MAX_RANKS)
        return fail(nullptr, GX_ERR_INVALID, "gx_create: bad rank %d of %d (at most %d ranks)", cfg->rank, cfg->n_ranks, SP_MAX_RANKS);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, GX_ERR_CUDA, "gx_create: no usable CUDA device (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= n_dev)
        return fail(nullptr, GX_ERR_INVALID, "gx_create: device %d of %d", cfg->device, n_dev);
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess)
        return fail(nullptr, GX_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10)
        return fail(nullptr, GX_ERR_CUDA, "gx_create: device %d is sm_%d%d; this library ships sm_100a code only", cfg->device,
                    prop.major, prop.minor);
    gx_ctx* c = new gx_ctx();
    c->cfg = *cfg;
    c->k = cfg->kmer_length;
    c->kw = (c->k + 31) / 32;
    c->ops = engine_ops(c->kw);
    if (cfg->reserved[0]) c->chunk_bytes = (size_t)cfg->reserved[0];
    if (cfg->reserved[1]) c->min_capacity = std::max<u64>(8192, cfg->reserved[1]);
    c->stream_records = cfg->reserved[2] & 1;           // records are serialised on demand by gx_next_records
    if (const char* e = getenv("GENOMIX_GB_SLICE")) c->slice_bytes = std::max<size_t>(4096, (size_t)atoll(e));   // tuning / tests
    c->test_start_small = (cfg->reserved[2] >> 8) & 1;  // test hook: start at min_capacity, no pilot -> exercises deferral + growth
    c->fixed_regions = (u32)std::min<u64>(cfg->reserved[3], SP_MAX_BUCKETS);
    auto bail = [&](int code) { g_create_error = c->err; gx_destroy(c); return code; };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(c, GX_ERR_CUDA, "stream"));
    c->own_stream = true;
    if (cudaMalloc((void**)&c->d_ctr, sizeof(Counters)) != cudaSuccess) return bail(fail(c, GX_ERR_NOMEM, "counters"));
    if (cudaMallocHost((void**)&c->h_ctr, sizeof(Counters)) != cudaSuccess) return bail(fail(c, GX_ERR_NOMEM, "pinned counters"));
    if (c->ops->prepare() != 0) return bail(fail(c, GX_ERR_CUDA, "cudaFuncSetAttribute failed"));
    c->spill_cap = 1ull << 20;
    for (int i = 0; i < 2; ++i) {
        if (ensure(c, c->spill_keys[i], (size_t)c->spill_cap * c->kw * sizeof(u64)) != GX_OK ||
            ensure(c, c->spill_meta[i], (size_t)c->spill_cap * sizeof(unsigned short)) != GX_OK ||
            ensure(c, c->spill_counts[i], (size_t)c->spill_cap * sizeof(u32)) != GX_OK)
            return bail(GX_ERR_NOMEM);
    }
    int r = gx_reset(c);
    if (r != GX_OK) return bail(r);
    *out = c;
    return GX_OK;
}

int gx_reset(gx_ctx* c) {
    if (!c) return GX_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    Counters z;
    memset(&z, 0, sizeof z);
    z.error = ~0ull;
    *c->h_ctr = z;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_ctr, c->h_ctr, sizeof z, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->table_live = false;  // keep the allocation; it is re-initialised on first use
    c->hash_mul = (u64)c->cfg.n_ranks;
    c->new_key_rate = 1.0;
    c->upserted_records = 0;
    c->merged_records = 0;
    c->split_redos = 0;
    c->spill_cur = 0;
    GX_TRY(set_spill_target(c));
    GX_TRY(mg_reset(c));
    if (c->heads.p) CUDA_TRY(c, cudaMemsetAsync(c->heads.p, 0, c->heads.cap, c->stream));
    c->grows = 0;
    c->global_lines = 0;
    c->n_nodes = c->record_bytes = 0;
    c->finished = false;
    c->sticky = 0;
    c->err.clear();
    c->slab_len = c->slab_off = 0;
    c->frame_byte_cursor = c->frame_rec_cursor = 0;
    memset(c->phase_ms, 0, sizeof c->phase_ms);
    return GX_OK;
}

void gx_destroy(gx_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    drain_timers(c);
    mg_destroy(c);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    DevBuf* bufs[] = {&c->heads, &c->store, &c->text, &c->text2, &c->nl_pos, &c->nl_pos2, &c->desc, &c->tile_sums, &c->ht_key, &c->ht_count, &c->ht_start,
                      &c->hgroup, &c->hentry, &c->hperm, &c->hoff, &c->bkey, &c->big_list, &c->tile_state, &c->records, &c->rec_offsets,
                      &c->parts, &c->dense, &c->dense_h, &c->big_tiles, &c->sort_perm[0], &c->sort_perm[1], &c->sort_hist, &c->sort_sizes,
                      &c->dense2, &c->dense_h2, &c->ring[0], &c->ring[1], &c->slice_idx,
                      &c->tile_prefix, &c->deferred[0], &c->deferred[1], &c->region_new, &c->gstats,
                      &c->mrec, &c->moff, &c->mbase, &c->mkeys, &c->mmeta, &c->mcounts};
    for (auto* b : bufs) release(*b);
    release_arena(c->arena);
    for (int i = 0; i < 2; ++i) { release(c->spill_keys[i]); release(c->spill_meta[i]); release(c->spill_counts[i]); }
    if (c->table) cudaFree(c->table);
    if (c->pilot) cudaFree(c->pilot);
    if (c->kept_table && c->kept_table != c->table) cudaFree(c->kept_table);
    if (c->d_ctr) cudaFree(c->d_ctr);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_written[i]) cudaEventDestroy(c->ev_written[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_text_ready[i]) cudaEventDestroy(c->ev_text_ready[i]);
        if (c->ev_text_free[i]) cudaEventDestroy(c->ev_text_free[i]);
    }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int gx_set_stream(gx_ctx* c, void* cuda_stream) {
    if (!c) return GX_ERR_INVALID;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return GX_OK;
}

int gx_push_lines_device(gx_ctx* c, const uint8_t* dev_text, size_t n_bytes) {
    GX_TRY(require_live(c));
    if (c->finished) return fail(c, GX_ERR_STATE, "gx_push_lines_device after gx_finish (call gx_reset first)");
    if (!dev_text && n_bytes) return fail(c, GX_ERR_INVALID, "null text");
    cudaSetDevice(c->cfg.device);
    // whole lines per chunk: the split points come from the device (last '\n' before the limit)
    size_t pos = 0;
    while (pos < n_bytes) {
        size_t len = std::min(c->chunk_bytes, n_bytes - pos);
        if (pos + len < n_bytes) {
            // back up to the end of the last whole line inside the window (found on the device)
            last_line_end_kernel<<<1, 1024, 0, c->stream>>>(dev_text, pos, pos + len, n_bytes, &c->d_ctr->scratch[0]);
            GX_TRY(check_launch(c, "last_line_end"));
            GX_TRY(sync_counters(c));
            const bool found = c->h_ctr->scratch[0] != ~0ull;
            if (found) len = (size_t)(c->h_ctr->scratch[0] - pos);
            if (!found) {  // one line longer than the chunk: extend to its end
                len = std::min(c->chunk_bytes * 4, n_bytes - pos);
                if (pos + len < n_bytes) return fail(c, GX_ERR_INVALID, "a single input line exceeds %zu bytes", c->chunk_bytes * 4);
            }
        }
        GX_TRY(push_chunk_device(c, dev_text + pos, len));
        pos += len;
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    return GX_OK;
}

int gx_push_lines(gx_ctx* c, const uint8_t* host_text, size_t n_bytes) {
    GX_TRY(require_live(c));
    if (c->finished) return fail(c, GX_ERR_STATE, "gx_push_lines after gx_finish (call gx_reset first)");
    if (!host_text && n_bytes) return fail(c, GX_ERR_INVALID, "null text");
    cudaSetDevice(c->cfg.device);
    // whole lines per chunk
    std::vector<std::pair<size_t, size_t>> chunks;
    for (size_t pos = 0; pos < n_bytes;) {
        size_t len = std::min(c->chunk_bytes, n_bytes - pos);
        if (pos + len < n_bytes) {
            const void* nl = memrchr(host_text + pos, '\n', len);
            if (nl) len = (const uint8_t*)nl - (host_text + pos) + 1;
            else {
                const void* fwd = memchr(host_text + pos + len, '\n', n_bytes - pos - len);
                len = fwd ? (size_t)((const uint8_t*)fwd - (host_text + pos) + 1) : n_bytes - pos;
            }
        }
        chunks.emplace_back(pos, len);
        pos += len;
    }
    if (chunks.empty()) return GX_OK;
    // Two text buffers: chunk i+1 travels to the device on the copy stream while chunk i is parsed, split and upserted.
    if (!c->copy_stream) GX_TRY(make_copy_stream(c));
    DevBuf* buf[2] = {&c->text, &c->text2};
    auto start_copy = [&](size_t i) -> int {
        const int b = (int)(i & 1);
        GX_TRY(ensure(c, *buf[b], chunks[i].second));
        if (i >= 2) CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_text_free[b], 0));   // chunk i-2 is done with the buffer
        PendingTimer t{PH_H2D, get_event(c), get_event(c)};
        cudaEventRecord(t.a, c->copy_stream);
        CUDA_TRY(c, cudaMemcpyAsync(buf[b]->p, host_text + chunks[i].first, chunks[i].second, cudaMemcpyHostToDevice, c->copy_stream));
        cudaEventRecord(t.b, c->copy_stream);
        c->timers.push_back(t);
        CUDA_TRY(c, cudaEventRecord(c->ev_text_ready[b], c->copy_stream));
        return GX_OK;
    };
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // kernels of an earlier push may still read the text buffers
    GX_TRY(start_copy(0));
    for (size_t i = 0; i < chunks.size(); ++i) {
        const int b = (int)(i & 1);
        if (i + 1 < chunks.size()) GX_TRY(start_copy(i + 1));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_text_ready[b], 0));
        GX_TRY(push_chunk_device(c, (const uint8_t*)buf[b]->p, chunks[i].second));
        CUDA_TRY(c, cudaEventRecord(c->ev_text_free[b], c->stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    return GX_OK;
}

int gx_push_fastq(gx_ctx* c, const uint8_t* host_r1, size_t n1, const uint8_t* host_r2, size_t n2, uint64_t first_record) {
    GX_TRY(require_live(c));
    if (c->finished) return fail(c, GX_ERR_STATE, "gx_push_fastq after gx_finish (call gx_reset first)");
    if ((!host_r1 && n1) || (!host_r2 && n2)) return fail(c, GX_ERR_INVALID, "null fastq buffer");
    if (n1 + n2 + 64 >= (1ull << 32)) return fail(c, GX_ERR_INVALID, "fastq chunk pair exceeds the 4 GiB chunk limit; push smaller chunks");
    if (n1 == 0 && n2 == 0) return GX_OK;
    cudaSetDevice(c->cfg.device);
    const size_t base2 = (n1 + 15) & ~(size_t)15;
    const size_t total = host_r2 ? base2 + n2 : n1;
    GX_TRY(ensure(c, c->text, total));
    {
        ScopedPhase ph(c, PH_H2D);
        CUDA_TRY(c, cudaMemcpyAsync(c->text.p, host_r1, n1, cudaMemcpyHostToDevice, c->stream));
        if (host_r2 && n2) CUDA_TRY(c, cudaMemcpyAsync((uint8_t*)c->text.p + base2, host_r2, n2, cudaMemcpyHostToDevice, c->stream));
    }
    GX_TRY(push_fastq_chunk_device(c, (const uint8_t*)c->text.p, total, n1, base2, host_r2 ? n2 : 0, first_record));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    return GX_OK;
}

// R2, merge half: AggregateKmerAggregateFactory.aggregate (:128-144) of serialised Nodes into the job's table.
int gx_push_records(gx_ctx* c, const uint8_t* host_records, size_t n_bytes) {
    GX_TRY(require_live(c));
    if (c->finished) return fail(c, GX_ERR_STATE, "gx_push_records after gx_finish (call gx_reset first)");
    if (!host_records && n_bytes) return fail(c, GX_ERR_INVALID, "null records");
    if (c->cfg.n_ranks > 1) return fail(c, GX_ERR_INVALID, "gx_push_records on a multi-rank ctx: fold the records in on one rank per key range");
    if (n_bytes == 0) return GX_OK;
    cudaSetDevice(c->cfg.device);
    // record boundaries: a chain of length fields (host walk: 4 bytes per record)
    std::vector<u64> off;
    off.reserve(n_bytes / 48 + 2);
    for (size_t pos = 0; pos < n_bytes;) {
        if (pos + 8 > n_bytes) return c->sticky = fail(c, GX_ERR_FORMAT, "truncated record header at byte %zu of the pushed stream", pos);
        const uint8_t* p = host_records + pos;
        const u64 len = 8ull + (((u64)p[0] << 24) | ((u64)p[1] << 16) | ((u64)p[2] << 8) | (u64)p[3]);
        if (pos + len > n_bytes) return c->sticky = fail(c, GX_ERR_FORMAT, "record at byte %zu runs past the end of the pushed stream", pos);
        off.push_back(pos);
        pos += len;
    }
    const u64 n = off.size();
    off.push_back(n_bytes);
    GX_TRY(ensure(c, c->mrec, n_bytes));
    GX_TRY(ensure(c, c->moff, (n + 1) * sizeof(u64)));
    GX_TRY(ensure(c, c->mbase, 2 * n * sizeof(u64)));
    GX_TRY(ensure(c, c->mkeys, n * c->kw * sizeof(u64)));
    GX_TRY(ensure(c, c->mmeta, n * sizeof(unsigned short)));
    GX_TRY(ensure(c, c->mcounts, n * sizeof(u32)));
    CUDA_TRY(c, cudaMemcpyAsync(c->mrec.p, host_records, n_bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->moff.p, off.data(), (n + 1) * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    MergeArgs a{};
    a.rec = (const uint8_t*)c->mrec.p; a.rec_off = (const u64*)c->moff.p; a.n_rec = n; a.k = c->k;
    a.head_base = (u64*)c->mbase.p; a.store_base = (u64*)c->mbase.p + n;
    a.ctr = c->d_ctr;
    a.keys = (u64*)c->mkeys.p; a.meta = (unsigned short*)c->mmeta.p; a.counts = (u32*)c->mcounts.p;
    a.order_base = (1ull << 62) + c->merged_records;   // behind every parsed line
    GX_TRY(sync_counters(c));
    const u64 heads0 = c->h_ctr->head_cursor, store0 = c->h_ctr->store_cursor;
    {
        ScopedPhase ph(c, PH_PARSE);
        c->ops->merge_scan(a, c->stream);
        GX_TRY(check_launch(c, "merge_scan"));
    }
    GX_TRY(sync_counters(c));   // also keeps `off` alive until the copy is done
    if (c->h_ctr->error != ~0ull) return line_error_to_status(c, c->h_ctr->error);
    GX_TRY(ensure(c, c->heads, (size_t)c->h_ctr->head_cursor * c->ops->head_bytes, (size_t)heads0 * c->ops->head_bytes, true));
    GX_TRY(ensure(c, c->store, (size_t)c->h_ctr->store_cursor, (size_t)store0));
    a.heads = c->heads.p; a.store = (uint8_t*)c->store.p;
    {
        ScopedPhase ph(c, PH_PARSE);
        c->ops->merge_apply(a, c->stream);
        GX_TRY(check_launch(c, "merge_apply"));
    }
    // room for the worst case (every record a new key), then the plain upsert of (key, mask, count) records
    const u64 distinct0 = c->table_live ? c->h_ctr->distinct : 0;
    const u64 want = std::max<u64>(c->min_capacity, (u64)((double)(distinct0 + n) / TARGET_LOAD) + 1);
    c->hash_mul = (u64)c->cfg.n_ranks;
    if (!c->table_live) GX_TRY(ensure_table(c, want));
    else if ((double)(distinct0 + n) > MAX_LOAD * (double)c->capacity) GX_TRY(grow_table_to(c, want));
    {
        ScopedPhase ph(c, PH_INSERT);
        c->ops->insert_records(a.keys, a.meta, a.counts, n, c->table, c->capacity, c->hash_mul, c->d_ctr, c->stream);
        GX_TRY(check_launch(c, "insert_records"));
    }
    GX_TRY(sync_counters(c));
    if (c->h_ctr->error != ~0ull) return line_error_to_status(c, c->h_ctr->error);
    GX_TRY(handle_spills(c));
    c->merged_records += n;
    return GX_OK;
}

int gx_finish(gx_ctx* c) {
    GX_TRY(require_live(c));
    if (c->finished) return GX_OK;
    cudaSetDevice(c->cfg.device);
    GX_TRY(sync_counters(c));
    if (c->h_ctr->error != ~0ull) return line_error_to_status(c, c->h_ctr->error);
    GX_TRY(handle_spills(c));
    if (c->cfg.n_ranks > 1) {
        GX_TRY(mg_complete(c));   // the round that is still on the wire (local: no collective here)
        u64 pending = 0;
        GX_TRY(mg_pending(c, &pending));
        if (pending) return fail(c, GX_ERR_STATE, "gx_finish: %llu routed records not exchanged yet (call gx_mg_exchange on every rank first)",
                                 (unsigned long long)pending);
    }
    if (!c->table || !c->table_live) GX_TRY(ensure_table(c, c->min_capacity));  // empty job: empty table, zero records
    const u64 cap = c->capacity;
    const u64 n_heads = c->h_ctr->head_cursor;
    const u64 distinct = c->h_ctr->distinct;
    const u64 n_tiles = (cap + ES_TILE - 1) / ES_TILE;
    EmitArgs a{};
    a.table = c->table; a.capacity = cap; a.k = c->k;
    {
        ScopedPhase ph(c, PH_FINISH);
        if (n_heads) {
            // read heads -> groups per node (gx_emit.cuh K3a); everything is sized by the number of heads
            if (n_heads >= 0x7fffffffull) return fail(c, GX_ERR_INVALID, "more than 2^31-1 read heads on one rank");
            u64 ht_size = 1024;
            while (ht_size < 2 * n_heads) ht_size <<= 1;
            const u64 s_tiles = (ht_size + TS_TILE - 1) / TS_TILE;
            GX_TRY(ensure(c, c->ht_key, ht_size * sizeof(u64)));
            GX_TRY(ensure(c, c->ht_count, ht_size * sizeof(u32)));
            GX_TRY(ensure(c, c->ht_start, (ht_size + 1) * sizeof(u32)));
            GX_TRY(ensure(c, c->hgroup, ht_size * sizeof(HeadGroup)));
            GX_TRY(ensure(c, c->hentry, n_heads * sizeof(u32)));
            GX_TRY(ensure(c, c->hperm, n_heads * sizeof(u32)));
            GX_TRY(ensure(c, c->hoff, n_heads * sizeof(u32)));
            GX_TRY(ensure(c, c->bkey, n_heads * sizeof(u64)));
            GX_TRY(ensure(c, c->big_list, (n_heads / HG_SMALL + 1) * sizeof(u32)));
            GX_TRY(ensure(c, c->tile_sums, (s_tiles + 1) * sizeof(u64)));
            CUDA_TRY(c, cudaMemsetAsync(c->ht_key.p, 0xff, ht_size * sizeof(u64), c->stream));
            CUDA_TRY(c, cudaMemsetAsync(c->ht_count.p, 0, ht_size * sizeof(u32), c->stream));
            CUDA_TRY(c, cudaMemsetAsync(&c->d_ctr->big_groups, 0, sizeof(u64), c->stream));
            c->ops->heads_lookup(c->heads.p, n_heads, c->table, cap, (u32)c->cfg.n_ranks, (u64*)c->ht_key.p, (u32*)c->ht_count.p,
                                 (u32)(ht_size - 1), (u32*)c->hentry.p, c->d_ctr, c->stream);
            GX_TRY(check_launch(c, "heads_lookup"));
            tile_sum_u32_kernel<<<(unsigned)s_tiles, TS_THREADS, 0, c->stream>>>((const u32*)c->ht_count.p, ht_size, (u64*)c->tile_sums.p);
            GX_TRY(check_launch(c, "tile_sum_u32"));
            scan_tile_sums_kernel<<<1, 1024, 0, c->stream>>>((u64*)c->tile_sums.p, s_tiles, &c->d_ctr->scratch[0]);
            GX_TRY(check_launch(c, "scan_tile_sums"));
            tile_scan_u32_kernel<<<(unsigned)s_tiles, TS_THREADS, 0, c->stream>>>((const u32*)c->ht_count.p, ht_size,
                                                                                  (const u64*)c->tile_sums.p, (u32*)c->ht_start.p);
            GX_TRY(check_launch(c, "tile_scan_u32"));
            CUDA_TRY(c, cudaMemsetAsync(c->ht_count.p, 0, ht_size * sizeof(u32), c->stream));
            heads_scatter_kernel<<<(unsigned)((n_heads + 255) / 256), 256, 0, c->stream>>>(
                (const u32*)c->hentry.p, n_heads, (const u32*)c->ht_start.p, (u32*)c->ht_count.p, (u32*)c->hperm.p);
            GX_TRY(check_launch(c, "heads_scatter"));
            c->ops->heads_group(c->heads.p, (const u64*)c->ht_key.p, (u32)ht_size, (const u32*)c->ht_start.p, (const u32*)c->ht_count.p,
                                (u32*)c->hperm.p, (u32*)c->hoff.p, (u64*)c->bkey.p, (HeadGroup*)c->hgroup.p, (u32*)c->big_list.p,
                                c->d_ctr, c->stream);
            GX_TRY(check_launch(c, "heads_group"));
            ++c->launches;   // heads_group_big
            a.heads = c->heads.p; a.hperm = (const u32*)c->hperm.p; a.hoff = (const u32*)c->hoff.p;
            a.store = (const uint8_t*)c->store.p;
            a.ht_key = (const u64*)c->ht_key.p; a.ht_mask = (u32)(ht_size - 1); a.group = (const HeadGroup*)c->hgroup.p;
        }
        // one pass over the table: record sizes, offsets, dense node list
        GX_TRY(ensure(c, c->tile_state, (n_tiles + 1) * 2 * sizeof(u64)));
        GX_TRY(ensure(c, c->dense, (size_t)std::max<u64>(distinct, 1) * (c->kw + 1) * sizeof(u64)));
        GX_TRY(ensure(c, c->dense_h, (size_t)std::max<u64>(distinct, 1) * sizeof(u32)));
        GX_TRY(ensure(c, c->rec_offsets, (size_t)(distinct + 1) * sizeof(u64)));
        CUDA_TRY(c, cudaMemsetAsync(c->tile_state.p, 0, (n_tiles + 1) * 2 * sizeof(u64), c->stream));
        a.tile_state = (u64*)c->tile_state.p;
        a.tile_counter = (u32*)((u64*)c->tile_state.p + 2 * n_tiles);
        a.totals = &c->d_ctr->scratch[0];
        a.dense = (u64*)c->dense.p; a.dense_h = (u32*)c->dense_h.p; a.rec_offsets = (u64*)c->rec_offsets.p;
        c->ops->emit_scan(a, c->stream);
        GX_TRY(check_launch(c, "emit_scan"));
    }
    GX_TRY(sync_counters(c));
    if (c->h_ctr->table_overflow)
        return c->sticky = fail(c, GX_ERR_NOMEM, "internal error: %llu k-mer inserts ran out of probe budget (table full?)",
                                (unsigned long long)c->h_ctr->table_overflow);
    if (c->h_ctr->heads_missing)
        return c->sticky = fail(c, GX_ERR_INVALID, "internal error: %llu read heads without a node",
                                (unsigned long long)c->h_ctr->heads_missing);
    c->record_bytes = c->h_ctr->scratch[0];
    c->n_nodes = c->h_ctr->scratch[1];
    if (c->n_nodes != distinct)
        return c->sticky = fail(c, GX_ERR_INVALID, "internal error: %llu occupied slots but %llu keys counted",
                                (unsigned long long)c->n_nodes, (unsigned long long)distinct);
    a.n_nodes = c->n_nodes;
    if (c->cfg.sort_output && c->n_nodes > 1) {
        // KmerPointable order (gx_sort.cuh): LSD radix sort of the node indices over the key bytes, then the node list, the
        // head-group references and the record offsets in the new order
        if (c->n_nodes >= 0xffffffffull) return fail(c, GX_ERR_INVALID, "sort_output: more than 2^32-1 nodes on one rank");
        const u64 n = c->n_nodes;
        const u32 nb = (u32)(c->k + 3) / 4, n_tiles_s = (u32)((n + RS_TILE - 1) / RS_TILE);
        const u64 s_tiles = (n + TS_TILE - 1) / TS_TILE;
        GX_TRY(ensure(c, c->sort_perm[0], n * sizeof(u32)));
        GX_TRY(ensure(c, c->sort_perm[1], n * sizeof(u32)));
        GX_TRY(ensure(c, c->sort_hist, (size_t)256 * n_tiles_s * sizeof(u32)));
        GX_TRY(ensure(c, c->dense2, (size_t)n * (c->kw + 1) * sizeof(u64)));
        GX_TRY(ensure(c, c->dense_h2, (size_t)n * sizeof(u32)));
        GX_TRY(ensure(c, c->sort_sizes, (size_t)n * sizeof(u32)));
        GX_TRY(ensure(c, c->tile_sums, (s_tiles + 1) * sizeof(u64)));
        ScopedPhase ph(c, PH_FINISH);
        SortArgs sa{};
        sa.dense = a.dense; sa.k = c->k; sa.n = n; sa.hist = (u32*)c->sort_hist.p; sa.n_tiles = n_tiles_s;
        int cur = 0;
        for (u32 byte = 0; byte < nb; ++byte) {
            sa.perm_in = byte ? (const u32*)c->sort_perm[cur].p : nullptr;
            sa.perm_out = (u32*)c->sort_perm[byte ? cur ^ 1 : 0].p;
            sa.byte = byte;
            c->ops->sort_pass(sa, c->stream);
            GX_TRY(check_launch(c, "sort_pass"));
            c->launches += 2;
            if (byte) cur ^= 1;
        }
        c->ops->sort_gather(a.dense, a.dense_h, a.rec_offsets, (const u32*)c->sort_perm[cur].p, n, (u64*)c->dense2.p, (u32*)c->dense_h2.p,
                            (u32*)c->sort_sizes.p, c->stream);
        GX_TRY(check_launch(c, "sort_gather"));
        tile_sum_u32_kernel<<<(unsigned)s_tiles, TS_THREADS, 0, c->stream>>>((const u32*)c->sort_sizes.p, n, (u64*)c->tile_sums.p);
        GX_TRY(check_launch(c, "tile_sum_u32"));
        scan_tile_sums_kernel<<<1, 1024, 0, c->stream>>>((u64*)c->tile_sums.p, s_tiles, &c->d_ctr->scratch[0]);
        GX_TRY(check_launch(c, "scan_tile_sums"));
        tile_scan_u32_to_u64_kernel<<<(unsigned)s_tiles, TS_THREADS, 0, c->stream>>>((const u32*)c->sort_sizes.p, n, (const u64*)c->tile_sums.p,
                                                                                     a.rec_offsets);
        GX_TRY(check_launch(c, "tile_scan_u32_to_u64"));
        std::swap(c->dense, c->dense2);
        std::swap(c->dense_h, c->dense_h2);
        a.dense = (u64*)c->dense.p; a.dense_h = (u32*)c->dense_h.p;
    }
    {
        // staging window per warp: 1.5x the average bytes of a tile of EW_NODES nodes; the rare tile that needs more is
        // written window by window
        const u64 avg = c->n_nodes ? c->record_bytes / c->n_nodes + 1 : 64;
        u64 stage = (avg * EW_NODES * 3 / 2 + 256 + 15) & ~15ull;
        a.stage_bytes = (u32)std::min<u64>(std::max<u64>(stage, 1024), EM_MAX_STAGE_BYTES / EW_WARPS);
    }
    // tiles whose span exceeds EW_MAX_WINDOWS windows: at most record_bytes / (EW_MAX_WINDOWS * stage_bytes) of them
    GX_TRY(ensure(c, c->big_tiles, (size_t)(c->record_bytes / ((u64)EW_MAX_WINDOWS * a.stage_bytes) + 2) * sizeof(u64)));
    a.big_tiles = (u64*)c->big_tiles.p;
    a.big_tile_count = &c->d_ctr->big_tiles;
    c->last_emit = a;
    c->stream_next_node = c->stream_next_byte = 0;
    if (!c->stream_records) {
        // the whole record stream, resident in device memory
        GX_TRY(ensure(c, c->records, (size_t)c->record_bytes + 16));
        ScopedPhase ph(c, PH_FINISH);
        a.out = (uint8_t*)c->records.p; a.out_base = 0; a.n_first = 0; a.n_last = c->n_nodes;
        CUDA_TRY(c, cudaMemsetAsync(a.big_tile_count, 0, sizeof(u64), c->stream));
        c->ops->emit_write(a, c->stream);
        ++c->launches;   // emit_write_big
        if (c->n_nodes) GX_TRY(check_launch(c, "emit_write"));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    c->finished = true;
    return GX_OK;
}

int64_t gx_num_nodes(gx_ctx* c) { return (c && c->finished) ? (int64_t)c->n_nodes : -1; }
int64_t gx_record_bytes(gx_ctx* c) { return (c && c->finished) ? (int64_t)c->record_bytes : -1; }

int gx_records_device(gx_ctx* c, const uint8_t** dev_records, const uint64_t** dev_offsets) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_records_device before gx_finish");
    if (c->stream_records && dev_records)
        return fail(c, GX_ERR_STATE, "gx_records_device: this job streams its records (gx_config.reserved[2] bit 0); use gx_next_records");
    if (dev_records) *dev_records = (const uint8_t*)c->records.p;
    if (dev_offsets) *dev_offsets = (const uint64_t*)c->rec_offsets.p;
    return GX_OK;
}

namespace {

// (node index, byte offset) of the last record boundary at or below `limit`; synchronises the stream
int record_boundary(gx_ctx* c, u64 limit, u64* node, u64* byte) {
    upper_bound_kernel<<<1, 1, 0, c->stream>>>((const u64*)c->rec_offsets.p, 0, c->n_nodes, limit, &c->d_ctr->scratch[0]);
    GX_TRY(check_launch(c, "upper_bound"));
    GX_TRY(sync_counters(c));
    *node = c->h_ctr->scratch[0];
    *byte = c->h_ctr->scratch[1];
    return GX_OK;
}

// Streaming delivery: records [cursor, end) are serialised slice by slice into two device buffers and copied to the
// host while the next slice is being written, so the record stream never has to exist in device memory as a whole.
int stream_records_to_host(gx_ctx* c, u64 node_lo, u64 cursor, u64 node_hi, u64 end, uint8_t* host_buf) {
    if (!c->copy_stream) GX_TRY(make_copy_stream(c));
    const u64 avg = c->n_nodes ? c->record_bytes / c->n_nodes + 1 : 64;
    const u64 slice_nodes = std::max<u64>(EW_NODES, (c->slice_bytes / avg) / EW_NODES * EW_NODES);
    // slice boundaries by node count (cheap: no search); the few slices that outgrow the buffers (large read-head sets)
    // make them grow
    std::vector<u64> bounds;   // byte offsets of the slice starts, from one small gather
    std::vector<u64> nodes;
    for (u64 n = node_lo; n < node_hi; n += slice_nodes) nodes.push_back(n);
    nodes.push_back(node_hi);
    bounds.resize(nodes.size());
    bounds.front() = cursor; bounds.back() = end;
    if (nodes.size() > 2) {
        GX_TRY(ensure(c, c->slice_idx, nodes.size() * 2 * sizeof(u64)));
        CUDA_TRY(c, cudaMemcpyAsync(c->slice_idx.p, nodes.data(), nodes.size() * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        gather_u64_kernel<<<(unsigned)((nodes.size() + 255) / 256), 256, 0, c->stream>>>(
            (const u64*)c->rec_offsets.p, (const u64*)c->slice_idx.p, nodes.size(), (u64*)c->slice_idx.p + nodes.size());
        GX_TRY(check_launch(c, "gather_u64"));
        CUDA_TRY(c, cudaMemcpyAsync(bounds.data(), (u64*)c->slice_idx.p + nodes.size(), nodes.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    EmitArgs a = c->last_emit;
    for (size_t s = 0; s + 1 < nodes.size(); ++s) {
        const int r = (int)(s & 1);
        const u64 bytes = bounds[s + 1] - bounds[s];
        if (s >= 2) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_copied[r], 0));   // buffer r is free again
        if (bytes + 16 > c->ring[r].cap) {
            if (s >= 2) CUDA_TRY(c, cudaEventSynchronize(c->ev_copied[r]));
            GX_TRY(ensure(c, c->ring[r], (size_t)bytes + 16));
        }
        a.out = (uint8_t*)c->ring[r].p; a.out_base = bounds[s]; a.n_first = nodes[s]; a.n_last = nodes[s + 1];
        {
            ScopedPhase ph(c, PH_FINISH);
            CUDA_TRY(c, cudaMemsetAsync(a.big_tile_count, 0, sizeof(u64), c->stream));
            c->ops->emit_write(a, c->stream);
            GX_TRY(check_launch(c, "emit_write"));
            ++c->launches;   // emit_write_big
        }
        CUDA_TRY(c, cudaEventRecord(c->ev_written[r], c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_written[r], 0));
        CUDA_TRY(c, cudaMemcpyAsync(host_buf + (bounds[s] - cursor), c->ring[r].p, (size_t)bytes, cudaMemcpyDeviceToHost, c->copy_stream));
        CUDA_TRY(c, cudaEventRecord(c->ev_copied[r], c->copy_stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    return GX_OK;
}

}  // namespace

int gx_next_records(gx_ctx* c, uint64_t* cursor, uint8_t* host_buf, size_t cap, size_t* used) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_next_records before gx_finish");
    if (!cursor || !used || (!host_buf && cap)) return fail(c, GX_ERR_INVALID, "null argument");
    cudaSetDevice(c->cfg.device);
    *used = 0;
    if (*cursor >= c->record_bytes) return GX_OK;
    u64 end = c->record_bytes, node_end = c->n_nodes;
    if (end - *cursor > cap) {
        // largest record boundary <= cursor + cap
        GX_TRY(record_boundary(c, *cursor + cap, &node_end, &end));
        if (end <= *cursor) return fail(c, GX_ERR_BUFFER, "buffer of %zu bytes cannot hold the next record", cap);
    }
    if (c->stream_records) {
        u64 node_lo = c->stream_next_node;
        if (*cursor != c->stream_next_byte) {   // not the continuation of the previous call: find the record
            u64 at = 0;
            GX_TRY(record_boundary(c, *cursor, &node_lo, &at));
            if (at != *cursor) return fail(c, GX_ERR_INVALID, "gx_next_records: cursor %llu is not a record boundary", (unsigned long long)*cursor);
        }
        GX_TRY(stream_records_to_host(c, node_lo, *cursor, node_end, end, host_buf));
        c->stream_next_node = node_end;
        c->stream_next_byte = end;
    } else {
        CUDA_TRY(c, cudaMemcpyAsync(host_buf, (const uint8_t*)c->records.p + *cursor, end - *cursor, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    *used = (size_t)(end - *cursor);
    *cursor = end;
    return GX_OK;
}

int gx_next_frame(gx_ctx* c, uint64_t* cursor, uint8_t* frame, int32_t frame_size, int32_t* n_tuples) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_next_frame before gx_finish");
    if (!cursor || !frame || !n_tuples || frame_size < 16) return fail(c, GX_ERR_INVALID, "bad argument");
    cudaSetDevice(c->cfg.device);
    *n_tuples = 0;
    if (*cursor == 0) { c->frame_byte_cursor = 0; c->frame_rec_cursor = 0; c->slab_len = 0; }
    if (*cursor != c->frame_rec_cursor) return fail(c, GX_ERR_INVALID, "gx_next_frame: frames must be pulled in order");
    const u32 nb = (u32)(c->k + 3) / 4;
    int32_t data_end = 0, count = 0;
    auto be32 = [](const uint8_t* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3]; };
    auto put32 = [](uint8_t* p, u32 v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; };
    while (c->frame_byte_cursor < c->record_bytes) {
        // make sure the next record is in the host slab
        auto in_slab = [&](u64 off, u64 len) { return off >= c->slab_off && off + len <= c->slab_off + c->slab_len; };
        if (!in_slab(c->frame_byte_cursor, 8) ||
            !in_slab(c->frame_byte_cursor, 8 + be32(&c->slab[c->frame_byte_cursor - c->slab_off]))) {
            u64 want = std::max<u64>(8ull << 20, 2 * (u64)frame_size);
            if (in_slab(c->frame_byte_cursor, 8))
                want = std::max<u64>(want, 8 + (u64)be32(&c->slab[c->frame_byte_cursor - c->slab_off]));
            want = std::min<u64>(want, c->record_bytes - c->frame_byte_cursor);
            if (c->stream_records) {
                // whole records through the streaming path; a record larger than the slab makes the slab grow
                for (;;) {
                    c->slab.resize(want);
                    uint64_t cur = c->frame_byte_cursor;
                    size_t got = 0;
                    const int rc = gx_next_records(c, &cur, c->slab.data(), (size_t)want, &got);
                    if (rc == GX_OK) { want = got; break; }
                    if (rc != GX_ERR_BUFFER || want >= c->record_bytes - c->frame_byte_cursor) return rc;
                    want = std::min<u64>(2 * want, c->record_bytes - c->frame_byte_cursor);
                }
            } else {
                c->slab.resize(want);
                CUDA_TRY(c, cudaMemcpyAsync(c->slab.data(), (const uint8_t*)c->records.p + c->frame_byte_cursor, want,
                                            cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            }
            c->slab_off = c->frame_byte_cursor;
            c->slab_len = want;
        }
        const uint8_t* r = &c->slab[c->frame_byte_cursor - c->slab_off];
        const u32 rec_len = be32(r);
        const u32 node_len = rec_len - (4 + nb);
        const int32_t tuple_len = 8 + (int32_t)nb + (int32_t)node_len;
        // FrameTupleAppender.append (FrameTupleAppender.java:57-70)
        if ((int64_t)data_end + tuple_len + 4 + (int64_t)(count + 1) * 4 > frame_size) {
            if (count == 0)
                return fail(c, GX_ERR_BUFFER, "Failed to copy an record into a frame: the record kmerByteSize is too large.");
            break;
        }
        uint8_t* t = frame + data_end;
        put32(t, nb);
        put32(t + 4, nb + node_len);
        memcpy(t + 8, r + 12, nb);                   // Kmer field: bytes without the VKmer length header
        memcpy(t + 8 + nb, r + 12 + nb, node_len);   // Node field
        data_end += tuple_len;
        ++count;
        put32(frame + frame_size - 4 - 4 * count, (u32)data_end);
        c->frame_byte_cursor += 8 + rec_len;
        ++c->frame_rec_cursor;
    }
    put32(frame + frame_size - 4, (u32)count);
    *n_tuples = count;
    *cursor = c->frame_rec_cursor;
    return GX_OK;
}

int gx_partition_records(gx_ctx* c, int32_t n_parts, int32_t* host_parts) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_partition_records before gx_finish");
    if (n_parts < 1 || !host_parts) return fail(c, GX_ERR_INVALID, "bad argument");
    if (c->stream_records) return fail(c, GX_ERR_STATE, "gx_partition_records needs the device-resident record stream (job streams its records)");
    if (c->n_nodes == 0) return GX_OK;
    cudaSetDevice(c->cfg.device);
    GX_TRY(ensure(c, c->parts, c->n_nodes * sizeof(int)));
    partition_records_kernel<<<(unsigned)((c->n_nodes + 255) / 256), 256, 0, c->stream>>>(
        (const uint8_t*)c->records.p, (const u64*)c->rec_offsets.p, c->n_nodes, n_parts, (int*)c->parts.p);
    GX_TRY(check_launch(c, "partition_records"));
    CUDA_TRY(c, cudaMemcpyAsync(host_parts, c->parts.p, c->n_nodes * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

// R4: SequenceFile v6 writer (hadoop-core 0.20.2 SequenceFile.Writer restated: header, append(), checkAndWriteSync()).
int gx_write_sequence_file(gx_ctx* c, const char* path, const uint8_t* sync16, int32_t n_parts, int32_t part,
                           uint64_t* bytes_written) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_write_sequence_file before gx_finish");
    if (!path || n_parts < 0 || (n_parts > 0 && (part < 0 || part >= n_parts))) return fail(c, GX_ERR_INVALID, "bad argument");
    cudaSetDevice(c->cfg.device);
    // stream the records through a pinned slab (allocated first so that no early return can leak the file handle)
    const size_t SLAB = 64ull << 20;
    uint8_t* slab = nullptr;
    CUDA_TRY(c, cudaMallocHost((void**)&slab, SLAB));
    FILE* f = fopen(path, "wb");
    if (!f) { cudaFreeHost(slab); return fail(c, GX_ERR_INVALID, "cannot open %s for writing", path); }
    uint8_t sync[16];
    if (sync16) memcpy(sync, sync16, 16);
    else {  // hadoop uses an MD5 of a fresh UID + time; any 16 bytes are valid -- derive them from the path
        u64 h0 = 0x9e3779b97f4a7c15ull, h1 = 0xd6e8feb86659fd93ull;
        for (const char* q = path; *q; ++q) { h0 = mix64(h0 ^ (u64)(unsigned char)*q); h1 = mix64(h1 + h0); }
        memcpy(sync, &h0, 8); memcpy(sync + 8, &h1, 8);
    }
    u64 pos = 0, last_sync = 0;
    bool io_ok = true;
    auto put = [&](const void* p, size_t nbytes) { if (nbytes && fwrite(p, 1, nbytes, f) != nbytes) io_ok = false; pos += nbytes; };
    auto put_string = [&](const char* str) {  // Text.writeString: vint length (one byte below 128) + UTF-8
        const uint8_t len = (uint8_t)strlen(str);
        put(&len, 1);
        put(str, len);
    };
    const uint8_t magic[4] = {'S', 'E', 'Q', 6};
    put(magic, 4);
    put_string("edu.uci.ics.genomix.data.types.VKmer");
    put_string("edu.uci.ics.genomix.data.types.Node");
    const uint8_t flags[2] = {0, 0};  // compressed, blockCompressed
    put(flags, 2);
    const uint8_t zero4[4] = {0, 0, 0, 0};
    put(zero4, 4);                    // metadata: 0 entries
    put(sync, 16);
    const uint8_t escape[4] = {0xff, 0xff, 0xff, 0xff};
    auto be32 = [](const uint8_t* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3]; };
    int rc = GX_OK;
    u64 cursor = 0;
    while (cursor < c->record_bytes && io_ok) {
        size_t used = 0;
        uint64_t cur = cursor;
        rc = gx_next_records(c, &cur, slab, SLAB, &used);
        if (rc != GX_OK || used == 0) break;
        for (size_t off = 0; off < used;) {
            const u32 rec_len = be32(slab + off);
            const u32 key_len = be32(slab + off + 4);
            bool mine = true;
            if (n_parts > 0) {  // KmerPartitionComputerFactory.partition on the Kmer bytes (VKmer minus its 4-byte header)
                int h = 1;
                for (u32 j = 4; j < key_len; ++j) h = 31 * h + (int)(signed char)slab[off + 8 + j];
                if (h < 0) h = -(h + 1);
                mine = (h % n_parts) == part;
            }
            if (mine) {
                if (pos >= last_sync + 2000) { put(escape, 4); put(sync, 16); last_sync = pos; }  // checkAndWriteSync
                put(slab + off, 8 + (size_t)rec_len);
            }
            off += 8 + (size_t)rec_len;
        }
        cursor = cur;
    }
    cudaFreeHost(slab);
    if (fclose(f) != 0) io_ok = false;
    if (rc != GX_OK) return rc;
    if (!io_ok) return fail(c, GX_ERR_INVALID, "I/O error while writing %s", path);
    if (bytes_written) *bytes_written = pos;
    return GX_OK;
}

int gx_graph_statistics(gx_ctx* c, gx_graph_stats* out) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_graph_statistics before gx_finish");
    if (!out) return fail(c, GX_ERR_INVALID, "null argument");
    static_assert(sizeof(GraphStatsDev) == sizeof(gx_graph_stats), "device and ABI statistics structs match");
    cudaSetDevice(c->cfg.device);
    GX_TRY(ensure(c, c->gstats, sizeof(GraphStatsDev)));
    CUDA_TRY(c, cudaMemsetAsync(c->gstats.p, 0, sizeof(GraphStatsDev), c->stream));
    EmitArgs a = c->last_emit;
    c->ops->graph_stats(a, (GraphStatsDev*)c->gstats.p, c->stream);
    if (c->n_nodes) GX_TRY(check_launch(c, "graph_stats"));
    CUDA_TRY(c, cudaMemcpyAsync(out, c->gstats.p, sizeof(GraphStatsDev), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

int gx_coverage_histogram(gx_ctx* c, uint64_t* host_bins, uint64_t n_bins) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_coverage_histogram before gx_finish");
    if (!host_bins || n_bins == 0) return fail(c, GX_ERR_INVALID, "bad argument");
    cudaSetDevice(c->cfg.device);
    GX_TRY(ensure(c, c->gstats, std::max<size_t>(sizeof(GraphStatsDev), (size_t)n_bins * sizeof(u64))));
    CUDA_TRY(c, cudaMemsetAsync(c->gstats.p, 0, (size_t)n_bins * sizeof(u64), c->stream));
    c->ops->coverage_histogram(c->last_emit, (u64*)c->gstats.p, n_bins, c->stream);
    if (c->n_nodes) GX_TRY(check_launch(c, "coverage_histogram"));
    CUDA_TRY(c, cudaMemcpyAsync(host_bins, c->gstats.p, (size_t)n_bins * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

// FittingMixture.fittingMixture (genomix-driver/.../mixture/model/FittingMixture.java:92-216) on the coverage histogram:
// `data` there holds one entry per node; every node of a bin behaves identically, so the sums run over bins weighted by
// their node counts. Densities as in commons-math3: Exponential(mean).density(x) = exp(-x/mean)/mean for x >= 0,
// Normal(mean, sd).density(x) = exp(-(x-mean)^2 / (2 sd^2)) / (sd sqrt(2 pi)).
int gx_coverage_cutoff(gx_ctx* c, int32_t iterations, int64_t* cutoff, double* exp_mean, double* normal_mean, double* normal_std) {
    GX_TRY(require_live(c));
    if (!c->finished) return fail(c, GX_ERR_STATE, "gx_coverage_cutoff before gx_finish");
    if (!cutoff || iterations < 0) return fail(c, GX_ERR_INVALID, "bad argument");
    gx_graph_stats gs;
    GX_TRY(gx_graph_statistics(c, &gs));
    const u64 max_cov = gs.coverage_max;
    if (max_cov == 0 || gs.nodes == 0) return fail(c, GX_ERR_STATE, "IllegalStateException: No information for coverage!");
    if (max_cov >= (1ull << 26)) return fail(c, GX_ERR_INVALID, "coverage histogram of %llu bins is too large", (unsigned long long)max_cov);
    std::vector<uint64_t> bins(max_cov + 1);
    GX_TRY(gx_coverage_histogram(c, bins.data(), max_cov + 1));
    std::vector<double> xs, ws;   // distinct coverage values and their node counts
    for (u64 v = 0; v <= max_cov; ++v)
        if (bins[v]) { xs.push_back((double)v); ws.push_back((double)bins[v]); }
    const size_t n = xs.size();
    auto exp_density = [](double mean, double x) { return x < 0 ? 0.0 : std::exp(-x / mean) / mean; };
    auto normal_density = [](double mean, double sd, double x) {
        const double z = (x - mean) / sd;
        return std::exp(-0.5 * z * z) / (sd * std::sqrt(2.0 * M_PI));
    };
    double e_mean = 5, n_mean = 20, n_sd = 5, p_exp = 0.5, p_norm = 0.5;   // initial guess (:97-109)
    std::vector<double> m_exp(n), m_norm(n);
    double exp_sum = 0, norm_sum = 0;
    auto expectation = [&]() {   // :118-140 / :166-189
        for (size_t i = 0; i < n; ++i) {
            double a = exp_density(e_mean, xs[i]) * p_exp, b = normal_density(n_mean, n_sd, xs[i]) * p_norm;
            if (a == 0) a = 0.000000001;
            if (b == 0) b = 0.000000001;
            m_exp[i] = a / (a + b);
            m_norm[i] = b / (a + b);
        }
        exp_sum = norm_sum = 0;
        for (size_t i = 0; i < n; ++i) { exp_sum += ws[i] * m_exp[i]; norm_sum += ws[i] * m_norm[i]; }
        p_exp = exp_sum / (exp_sum + norm_sum);
        p_norm = norm_sum / (exp_sum + norm_sum);
    };
    expectation();
    for (int it = 0; it < iterations; ++it) {   // maximisation (:144-164)
        double em = 0, nm = 0, nv = 0;
        for (size_t i = 0; i < n; ++i) { em += ws[i] * m_exp[i] * xs[i]; nm += ws[i] * m_norm[i] * xs[i]; }
        e_mean = em / exp_sum;
        n_mean = nm / norm_sum;
        for (size_t i = 0; i < n; ++i) nv += ws[i] * m_norm[i] * (xs[i] - n_mean) * (xs[i] - n_mean);
        n_sd = std::sqrt(nv / norm_sum);
        if (n_sd == 0) n_sd = 0.000000001f;
        expectation();
    }
    *cutoff = 0;   // :208-215
    for (double cov = 1; cov < (double)max_cov; cov += 1) {
        if (p_exp * exp_density(e_mean, cov) < p_norm * normal_density(n_mean, n_sd, cov)) { *cutoff = (int64_t)cov; break; }
    }
    if (exp_mean) *exp_mean = e_mean;
    if (normal_mean) *normal_mean = n_mean;
    if (normal_std) *normal_std = n_sd;
    return GX_OK;
}

int gx_get_stats(gx_ctx* c, gx_stats* out) {
    if (!c || !out) return GX_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    GX_TRY(sync_counters(c));
    memset(out, 0, sizeof *out);
    out->lines = c->global_lines;
    out->reads = c->h_ctr->reads;
    out->bases = c->h_ctr->bases;
    out->kmer_occurrences = c->h_ctr->occurrences;
    out->distinct_kmers = c->h_ctr->distinct;
    out->read_heads = c->h_ctr->read_heads;
    out->record_bytes = c->record_bytes;
    out->table_capacity = c->capacity;
    out->table_grows = c->grows;
    out->exchanged_records = mg_exchanged(c);
    out->split_redos = c->split_redos;
    return GX_OK;
}

int gx_phase_ms(gx_ctx* c, float out_ms[8]) {
    if (!c || !out_ms) return GX_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    drain_timers(c);
    for (int i = 0; i < 8; ++i) out_ms[i] = c->phase_ms[i];
    return GX_OK;
}

uint64_t gx_kernel_launches(gx_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"

#include "gx_mg.inl"
