// gx_engine.cuh -- the KW-templated kernels of the graph-build path behind one host-callable launch table.
//
//   gx_parse.cuh   line index, line / fastq parser                                    (k-mer-width independent)
//   gx_split.cuh   K1 split_count / split_place: read -> (key, edge mask) records sorted by (owner GPU, table region)
//                  K2 upsert_regions: region-by-region, L2-resident hash-table upserts
//   gx_build.cuh   extraction helpers, insert_records (spills), init_table, rehash
//   gx_emit.cuh    K3 read-head grouping, sizing, Node serialisation, graph statistics, Java partition hash
//   gx_merge.cuh   serialised Node records back into (key, mask, count) + read heads (gx_push_records)
//   gx_sort.cuh    optional KmerPointable order of the output (gx_config.sort_output)
#pragma once
#include "gx_emit.cuh"
#include "gx_merge.cuh"
#include "gx_sort.cuh"
#include "gx_split.cuh"

namespace gx {

// Host-callable launch table, one instance per KW (instantiated in gx_kw<N>.cu).
struct EngineOps {
    int kw;
    int upsert_warps, upsert_blocks;   // warps per CTA and CTAs per SM of the region upsert kernel
    size_t slot_bytes;
    size_t head_bytes;
    void (*init_table)(u64* table, u64 capacity, cudaStream_t st);
    void (*split_count)(const SplitArgs& a, cudaStream_t st);
    void (*split_place)(const SplitArgs& a, cudaStream_t st);
    void (*upsert_regions)(const UpsertArgs& a, unsigned grid, cudaStream_t st);
    void (*check_arena)(const u64* keys, const u64* tab, u32 n_ranks, u32 n_regions, u64* bad, cudaStream_t st);
    void (*insert_records)(const u64* keys, const unsigned short* meta, const u32* counts, u64 n, u64* table,
                           u64 capacity, u64 hash_mul, Counters* ctr, cudaStream_t st);
    void (*rehash)(const u64* old_table, u64 old_capacity, u64* table, u64 capacity, u64 hash_mul, Counters* ctr, cudaStream_t st);
    void (*heads_lookup)(const void* heads, u64 n_heads, u64* table, u64 capacity, u32 n_ranks, u64* ht_key, u32* ht_count,
                         u32 ht_mask, u32* hentry, Counters* ctr, cudaStream_t st);
    void (*heads_group)(const void* heads, const u64* ht_key, u32 ht_size, const u32* ht_start, const u32* ht_count, u32* hperm,
                        u32* hoff, u64* bkey, HeadGroup* group, u32* big_list, Counters* ctr, cudaStream_t st);
    void (*emit_scan)(const EmitArgs& a, cudaStream_t st);
    void (*emit_write)(const EmitArgs& a, cudaStream_t st);
    void (*graph_stats)(const EmitArgs& a, GraphStatsDev* out, cudaStream_t st);
    void (*coverage_histogram)(const EmitArgs& a, u64* bins, u64 n_bins, cudaStream_t st);
    void (*sort_pass)(const SortArgs& a, cudaStream_t st);   // histogram, scan, stable scatter of one key byte
    void (*sort_gather)(const u64* dense, const u32* dense_h, const u64* rec_offsets, const u32* perm, u64 n, u64* dense_out,
                        u32* dense_h_out, u32* sizes, cudaStream_t st);
    void (*merge_scan)(const MergeArgs& a, cudaStream_t st);
    void (*merge_apply)(const MergeArgs& a, cudaStream_t st);
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
