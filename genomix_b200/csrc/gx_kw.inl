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
void l_check_arena(const u64* keys, const u64* seg_start, u32 n_ranks, u32 n_regions, u64* bad, cudaStream_t st) {
    check_arena_kernel<KW><<<148 * 8, 256, 0, st>>>(keys, seg_start, n_ranks, n_regions, bad);
}
void l_insert_records(const u64* keys, const unsigned short* meta, const u32* counts, u64 n, u64* table, u64 capacity,
                      u64 hash_mul, Counters* ctr, cudaStream_t st) {
    if (n == 0) return;
    insert_records_kernel<KW><<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(keys, meta, counts, n, table, capacity, hash_mul, ctr);
}
void l_rehash(const u64* old_table, u64 old_capacity, u64* table, u64 capacity, u64 hash_mul, Counters* ctr, cudaStream_t st) {
    rehash_kernel<KW><<<grid_for(old_capacity, 256, 148 * 8), 256, 0, st>>>(old_table, old_capacity, table, capacity, hash_mul, ctr);
}
void l_heads_count(const void* heads, u64 n_heads, const u64* table, u64 capacity, u32 n_ranks, u64* hslot, u32* hcount,
                   Counters* ctr, cudaStream_t st) {
    if (n_heads == 0) return;
    heads_count_kernel<KW><<<(unsigned)((n_heads + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const Head<KW>*>(heads), n_heads, table, capacity, n_ranks, hslot, hcount, ctr);
}
void l_heads_sort(const void* heads, const u64* hslot, u64 n_heads, u64 capacity, const u32* hstart, u32* hcount,
                  u32* hperm, Counters* ctr, cudaStream_t st) {
    if (n_heads == 0) return;
    heads_sort_kernel<KW><<<(unsigned)((n_heads + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const Head<KW>*>(heads), hslot, n_heads, capacity, hstart, hcount, hperm, ctr);
}
void l_emit_size(const EmitArgs& a, cudaStream_t st) {
    emit_size_kernel<KW><<<(unsigned)((a.capacity + EM_TILE - 1) / EM_TILE), EM_THREADS, 0, st>>>(a);
}
void l_emit_compact(const EmitArgs& a, cudaStream_t st) {
    emit_compact_kernel<KW><<<(unsigned)((a.capacity + EM_TILE - 1) / EM_TILE), EM_THREADS, 0, st>>>(a);
}
void l_emit_serialise(const EmitArgs& a, cudaStream_t st) {
    if (a.n_nodes == 0) return;
    emit_serialise_kernel<KW><<<(unsigned)((a.n_nodes + EM_THREADS - 1) / EM_THREADS), EM_THREADS, a.stage_bytes, st>>>(a);
}
void l_graph_stats(const EmitArgs& a, GraphStatsDev* out, cudaStream_t st) {
    if (a.n_nodes == 0) return;
    graph_stats_kernel<KW><<<grid_for(a.n_nodes, 256, 148 * 8), 256, 0, st>>>(a, out);
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
    int r = (int)cudaFuncSetAttribute(emit_serialise_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_MAX_STAGE_BYTES);
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
                       l_heads_count,
                       l_heads_sort,
                       l_emit_size,
                       l_emit_compact,
                       l_emit_serialise,
                       l_graph_stats,
                       l_route_heads,
                       l_rebase_heads,
                       l_prepare};
}  // namespace

#define GX_CAT2(a, b) a##b
#define GX_CAT(a, b) GX_CAT2(a, b)
const EngineOps* GX_CAT(engine_ops_kw, GX_KW)() { return &OPS; }
}  // namespace gx
