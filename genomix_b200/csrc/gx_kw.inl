// gx_kw.inl -- instantiates the engine for one key width; included by gx_kw<N>.cu with GX_KW defined.
#include <algorithm>
#include "gx_engine.cuh"

namespace gx {
namespace {
constexpr int KW = GX_KW;

void l_init_table(u64* table, u64 capacity, cudaStream_t st) {
    init_table_kernel<KW><<<grid_for(capacity * SlotTraits<KW>::WORDS, 256, 148 * 32), 256, 0, st>>>(table, capacity);
}
void l_split_count(const SplitArgs& a, cudaStream_t st) {
    split_count_kernel<KW><<<grid_for(a.n_lines, SP_WARPS, 148 * SplitBlocks<KW>::MIN), SP_THREADS, 0, st>>>(a);
}
void l_split_place(const SplitArgs& a, cudaStream_t st) {
    // one resident wave: the multisplit's global reservations then walk every bucket front to back exactly once per CTA
    split_place_kernel<KW><<<grid_for(a.n_lines, SP_WARPS, 148 * SplitBlocks<KW>::MIN), SP_THREADS, sizeof(SplitSmem<KW>), st>>>(a);
}
void l_upsert_regions(const UpsertArgs& a, unsigned grid, cudaStream_t st) {
    upsert_regions_kernel<KW><<<grid, UpsertCfg<KW>::THREADS, sizeof(UpsertSmem<KW>), st>>>(a);
}
void l_check_arena(const u64* keys, const u64* tab, u32 n_ranks, u32 n_regions, u64* bad, cudaStream_t st) {
    check_arena_kernel<KW><<<148 * 8, 256, 0, st>>>(keys, tab, n_ranks, n_regions, bad);
}
void l_insert_records(const u64* keys, const unsigned short* meta, const u32* counts, u64 n, u64* table, u64 capacity,
                      u64 hash_mul, Counters* ctr, cudaStream_t st) {
    if (n == 0) return;
    insert_records_kernel<KW><<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(keys, meta, counts, n, table, capacity, hash_mul, ctr);
}
void l_rehash(const u64* old_table, u64 old_capacity, u64* table, u64 capacity, u64 hash_mul, Counters* ctr, cudaStream_t st) {
    rehash_kernel<KW><<<grid_for(old_capacity, 256, 148 * 8), 256, 0, st>>>(old_table, old_capacity, table, capacity, hash_mul, ctr);
}
void l_heads_lookup(const void* heads, u64 n_heads, u64* table, u64 capacity, u32 n_ranks, u64* ht_key, u32* ht_count, u32 ht_mask,
                    u32* hentry, Counters* ctr, cudaStream_t st) {
    if (n_heads == 0) return;
    heads_lookup_kernel<KW><<<(unsigned)((n_heads + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const Head<KW>*>(heads), n_heads, table, capacity, n_ranks, ht_key, ht_count, ht_mask, hentry, ctr);
}
void l_heads_group(const void* heads, const u64* ht_key, u32 ht_size, const u32* ht_start, const u32* ht_count, u32* hperm, u32* hoff,
                   u64* bkey, HeadGroup* group, u32* big_list, Counters* ctr, cudaStream_t st) {
    const Head<KW>* h = reinterpret_cast<const Head<KW>*>(heads);
    heads_group_kernel<KW><<<(ht_size + 255) / 256, 256, 0, st>>>(h, ht_key, ht_size, ht_start, ht_count, hperm, hoff, group, big_list, ctr);
    // the (normally empty) list of large groups: a fixed grid, every CTA walks the list
    heads_group_big_kernel<KW><<<148, HB_THREADS, 0, st>>>(h, ht_start, ht_count, hperm, hoff, bkey, group, big_list, ctr);
}
void l_emit_scan(const EmitArgs& a, cudaStream_t st) {
    const u64 tiles = (a.capacity + ES_TILE - 1) / ES_TILE;   // persistent warps: tickets for tiles of ES_TILE slots
    emit_scan_kernel<KW><<<grid_for(tiles, EW_WARPS, 148 * 8), EM_THREADS, 0, st>>>(a);
}
void l_emit_write(const EmitArgs& a, cudaStream_t st) {
    if (a.n_last <= a.n_first) return;
    // persistent warps: every warp walks tiles of 32 nodes
    const u64 tiles = (a.n_last - a.n_first + EW_NODES - 1) / EW_NODES;
    emit_write_kernel<KW><<<grid_for(tiles, EW_WARPS, 148 * 8), EM_THREADS, (size_t)a.stage_bytes * EW_WARPS, st>>>(a);
    // the (normally empty) list of tiles with a huge record: a fixed grid, every warp walks the list
    emit_write_big_kernel<KW><<<148 * 2, EM_THREADS, (size_t)a.stage_bytes * EW_WARPS, st>>>(a);
}
void l_graph_stats(const EmitArgs& a, GraphStatsDev* out, cudaStream_t st) {
    if (a.n_nodes == 0) return;
    graph_stats_kernel<KW><<<grid_for(a.n_nodes, 256, 148 * 8), 256, 0, st>>>(a, out);
}
void l_coverage_histogram(const EmitArgs& a, u64* bins, u64 n_bins, cudaStream_t st) {
    if (a.n_nodes == 0) return;
    coverage_histogram_kernel<KW><<<grid_for(a.n_nodes, 256, 148 * 4), 256, 0, st>>>(a, bins, n_bins);
}
void l_sort_pass(const SortArgs& a, cudaStream_t st) {
    const unsigned grid = (a.n_tiles + RS_WARPS - 1) / RS_WARPS;
    sort_hist_kernel<KW><<<grid, RS_WARPS * 32, 0, st>>>(a);
    sort_scan_kernel<<<1, 1024, 0, st>>>(a.hist, (u64)256 * a.n_tiles);
    sort_scatter_kernel<KW><<<grid, RS_WARPS * 32, 0, st>>>(a);
}
void l_sort_gather(const u64* dense, const u32* dense_h, const u64* rec_offsets, const u32* perm, u64 n, u64* dense_out, u32* dense_h_out,
                   u32* sizes, cudaStream_t st) {
    sort_gather_kernel<KW><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dense, dense_h, rec_offsets, perm, n, dense_out, dense_h_out, sizes);
}
void l_merge_scan(const MergeArgs& a, cudaStream_t st) {
    if (a.n_rec) merge_scan_kernel<KW><<<(unsigned)((a.n_rec + 255) / 256), 256, 0, st>>>(a);
}
void l_merge_apply(const MergeArgs& a, cudaStream_t st) {
    if (a.n_rec) merge_apply_kernel<KW><<<(unsigned)((a.n_rec + 255) / 256), 256, 0, st>>>(a);
}
void l_route_heads(const HeadRouteArgs& a, cudaStream_t st) {
    if (a.n == 0) return;
    route_heads_kernel<KW><<<(unsigned)((a.n + 255) / 256), 256, 0, st>>>(a);
}
void l_rebase_heads(void* heads, u64 first, u64 n, u64 store_base, cudaStream_t st) {
    if (n == 0) return;
    rebase_heads_kernel<KW><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(heads, first, n, store_base);
}
int l_prepare() {
    int r = (int)cudaFuncSetAttribute(emit_write_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_MAX_STAGE_BYTES);
    if (r == 0)
        r = (int)cudaFuncSetAttribute(emit_write_big_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_MAX_STAGE_BYTES);
    if (r == 0)
        r = (int)cudaFuncSetAttribute(split_place_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem<KW>));
    if (r == 0)
        r = (int)cudaFuncSetAttribute(upsert_regions_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UpsertSmem<KW>));
    return r;
}

const EngineOps OPS = {KW,
                       UpsertCfg<KW>::WARPS,
                       UpsertCfg<KW>::MIN_BLOCKS,
                       sizeof(u64) * SlotTraits<KW>::WORDS,
                       sizeof(Head<KW>),
                       l_init_table,
                       l_split_count,
                       l_split_place,
                       l_upsert_regions,
                       l_check_arena,
                       l_insert_records,
                       l_rehash,
                       l_heads_lookup,
                       l_heads_group,
                       l_emit_scan,
                       l_emit_write,
                       l_graph_stats,
                       l_coverage_histogram,
                       l_sort_pass,
                       l_sort_gather,
                       l_merge_scan,
                       l_merge_apply,
                       l_route_heads,
                       l_rebase_heads,
                       l_prepare};
}  // namespace

#define GX_CAT2(a, b) a##b
#define GX_CAT(a, b) GX_CAT2(a, b)
const EngineOps* GX_CAT(engine_ops_kw, GX_KW)() { return &OPS; }
}  // namespace gx
