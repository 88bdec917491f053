// gx_kw.inl -- instantiates the engine for one key width; included by gx_kw<N>.cu with GX_KW defined.
#include <algorithm>
#include "gx_engine.cuh"

namespace gx {
namespace {
constexpr int KW = GX_KW;

void l_init_table(u64* table, u64 capacity, cudaStream_t st) {
    init_table_kernel<KW><<<grid_for(capacity * SlotTraits<KW>::WORDS, 256, 148 * 32), 256, 0, st>>>(table, capacity);
}
void l_extract_insert(const ExtractArgs& a, cudaStream_t st) {
    extract_kernel<KW, EX_UPSERT><<<grid_for(a.n_lines, EX_WARPS, 148 * 8), EX_THREADS, 0, st>>>(a);
}
void l_extract_route(const ExtractArgs& a, cudaStream_t st) {
    extract_kernel<KW, EX_ROUTE><<<grid_for(a.n_lines, EX_WARPS, 148 * 8), EX_THREADS, 0, st>>>(a);
}
void l_extract_flat(const ExtractArgs& a, cudaStream_t st) {
    extract_kernel<KW, EX_FLAT><<<grid_for(a.n_lines, EX_WARPS, 148 * 8), EX_THREADS, 0, st>>>(a);
}
void l_partition_flat(const u64* flat_keys, const unsigned short* flat_meta, u64 n, u32 n_buckets, u64* bucket_cursor,
                      u64* out_keys, unsigned short* out_meta, cudaStream_t st) {
    if (n == 0) return;
    partition_flat_kernel<KW><<<(unsigned)((n + PT_THREADS * PT_ITEMS - 1) / (PT_THREADS * PT_ITEMS)), PT_THREADS, 0, st>>>(
        flat_keys, flat_meta, n, n_buckets, bucket_cursor, out_keys, out_meta);
}
void l_insert_records(const u64* keys, const unsigned short* meta, const u32* counts, u64 n, u64* table, u64 capacity,
                      Counters* ctr, cudaStream_t st) {
    if (n == 0) return;
    // short-lived CTAs (4 records per thread) so that concurrent kernels of other streams (NCCL) get SM slots promptly
    insert_records_kernel<KW><<<(unsigned)std::min<u64>((n + 1023) / 1024, 1u << 30), 256, 0, st>>>(keys, meta, counts, n, table, capacity, ctr);
}
void l_rehash(const u64* old_table, u64 old_capacity, u64* table, u64 capacity, cudaStream_t st) {
    rehash_kernel<KW><<<grid_for(old_capacity, 256, 148 * 8), 256, 0, st>>>(old_table, old_capacity, table, capacity);
}
void l_heads_count(const void* heads, u64 n_heads, const u64* table, u64 capacity, u64* hslot, u32* hcount,
                   Counters* ctr, cudaStream_t st) {
    if (n_heads == 0) return;
    heads_count_kernel<KW><<<(unsigned)((n_heads + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const Head<KW>*>(heads), n_heads, table, capacity, hslot, hcount, ctr);
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
    return (int)cudaFuncSetAttribute(emit_serialise_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_MAX_STAGE_BYTES);
}

const EngineOps OPS = {KW,
                       sizeof(u64) * SlotTraits<KW>::WORDS,
                       sizeof(Head<KW>),
                       l_init_table,
                       l_extract_insert,
                       l_extract_route,
                       l_extract_flat,
                       l_partition_flat,
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
