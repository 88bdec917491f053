// gx_mg.inl -- multi-GPU hash-partitioned exchange (included at the end of gx_api.cu).
//
// Replaces the reference's M:N hash-partitioning connector between the local and the global aggregator
// (MToNPartitioningMergingConnectorDescriptor with KmerPartitionComputerFactory, JobGenBuildBrujinGraph.java:
// 132-133; sender PartitionDataWriter.java:79-94 over Hyracks' TCP stack): every k-mer occurrence is owned by
// rank owner_of(hash(key)); the extract kernel inserts own keys directly and appends the others to per-owner
// send buckets; gx_mg_exchange() ships the buckets with one NCCL all-to-all-v (ncclSend/ncclRecv group over
// NVLink) and upserts what arrives. Read heads travel the same way with their packed sequences.
#include <nccl.h>

namespace {

#define NCCL_TRY(c, expr)                                                                        \
    do {                                                                                         \
        ncclResult_t r__ = (expr);                                                               \
        if (r__ != ncclSuccess)                                                                  \
            return fail(c, GX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, ncclGetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

struct MgState {
    ncclComm_t comm = nullptr;
    int n = 1, rank = 0;
    std::vector<DevBuf> route_keys, route_meta, send_heads, send_store;  // per destination
    DevBuf ptr_table;      // device copy of the 4 pointer arrays: [4][n]
    DevBuf counts;         // device u64 [3n]: kmer records, heads, store bytes per destination
    DevBuf all_counts;     // device u64 [n][3n] after the all-gather
    DevBuf recv_keys, recv_meta;
    std::vector<u64> h_counts;  // host copy of counts (valid after a sync)
    u64 routed_heads_upto = 0, routed_store_upto = 0;
    u64 exchanged = 0;
};

MgState* mg_of(gx_ctx* c) { return reinterpret_cast<MgState*>(c->mg); }

int mg_upload_ptrs(gx_ctx* c) {
    MgState* m = mg_of(c);
    std::vector<void*> h((size_t)4 * m->n);
    for (int d = 0; d < m->n; ++d) {
        h[0 * m->n + d] = m->route_keys[d].p;
        h[1 * m->n + d] = m->route_meta[d].p;
        h[2 * m->n + d] = m->send_heads[d].p;
        h[3 * m->n + d] = m->send_store[d].p;
    }
    CUDA_TRY(c, cudaMemcpyAsync(m->ptr_table.p, h.data(), h.size() * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // h is a stack vector
    return GX_OK;
}

int mg_fetch_counts(gx_ctx* c) {
    MgState* m = mg_of(c);
    m->h_counts.resize((size_t)3 * m->n);
    CUDA_TRY(c, cudaMemcpyAsync(m->h_counts.data(), m->counts.p, m->h_counts.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

// make sure every destination bucket can take `incoming` more k-mer records; fills the routing fields of `a`
int mg_prepare_route(gx_ctx* c, u64 incoming, ExtractArgs& a) {
    MgState* m = mg_of(c);
    if (!m || !m->comm) return fail(c, GX_ERR_STATE, "n_ranks > 1 but gx_mg_init has not been called");
    GX_TRY(mg_fetch_counts(c));
    for (int d = 0; d < m->n; ++d) {
        if (d == m->rank) continue;
        const u64 have = m->h_counts[d];
        GX_TRY(ensure(c, m->route_keys[d], (size_t)(have + incoming) * c->kw * sizeof(u64), (size_t)have * c->kw * sizeof(u64)));
        GX_TRY(ensure(c, m->route_meta[d], (size_t)(have + incoming) * sizeof(unsigned short), (size_t)have * sizeof(unsigned short)));
    }
    GX_TRY(mg_upload_ptrs(c));
    a.n_ranks = (u32)m->n;
    a.rank = (u32)m->rank;
    a.route_keys = reinterpret_cast<u64* const*>(m->ptr_table.p);
    a.route_meta = reinterpret_cast<unsigned short* const*>((void**)m->ptr_table.p + m->n);
    a.route_count = (u64*)m->counts.p;
    return GX_OK;
}

int mg_pending(gx_ctx* c, u64* pending) {
    MgState* m = mg_of(c);
    *pending = 0;
    if (!m) return fail(c, GX_ERR_STATE, "n_ranks > 1 but gx_mg_init has not been called");
    GX_TRY(mg_fetch_counts(c));
    for (int d = 0; d < m->n; ++d) *pending += m->h_counts[d];
    // heads created after the last exchange would also be lost
    if (c->h_ctr->head_cursor != m->routed_heads_upto) *pending += c->h_ctr->head_cursor - m->routed_heads_upto;
    return GX_OK;
}

u64 mg_exchanged(gx_ctx* c) { return mg_of(c) ? mg_of(c)->exchanged : 0; }

void mg_destroy(gx_ctx* c) {
    MgState* m = mg_of(c);
    if (!m) return;
    if (m->comm) ncclCommDestroy(m->comm);
    for (auto* v : {&m->route_keys, &m->route_meta, &m->send_heads, &m->send_store})
        for (auto& b : *v) release(b);
    release(m->ptr_table); release(m->counts); release(m->all_counts); release(m->recv_keys); release(m->recv_meta);
    delete m;
    c->mg = nullptr;
}

int mg_reset(gx_ctx* c) {
    MgState* m = mg_of(c);
    if (!m) return GX_OK;
    if (m->counts.p) CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)3 * m->n * sizeof(u64), c->stream));
    m->routed_heads_upto = m->routed_store_upto = 0;
    m->exchanged = 0;
    return GX_OK;
}

}  // namespace

extern "C" {

int gx_mg_unique_id(uint8_t out_id[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return GX_ERR_CUDA;
    memcpy(out_id, &id, 128);
    return GX_OK;
}

int gx_mg_init(gx_ctx* c, const uint8_t id_bytes[128]) {
    GX_TRY(require_live(c));
    if (c->cfg.n_ranks < 2) return fail(c, GX_ERR_INVALID, "gx_mg_init on a single-rank ctx");
    if (c->mg) return fail(c, GX_ERR_STATE, "gx_mg_init called twice");
    cudaSetDevice(c->cfg.device);
    MgState* m = new MgState();
    c->mg = m;
    m->n = c->cfg.n_ranks;
    m->rank = c->cfg.rank;
    m->route_keys.resize(m->n); m->route_meta.resize(m->n); m->send_heads.resize(m->n); m->send_store.resize(m->n);
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    NCCL_TRY(c, ncclCommInitRank(&m->comm, m->n, id, m->rank));
    GX_TRY(ensure(c, m->ptr_table, (size_t)4 * m->n * sizeof(void*)));
    GX_TRY(ensure(c, m->counts, (size_t)3 * m->n * sizeof(u64)));
    GX_TRY(ensure(c, m->all_counts, (size_t)3 * m->n * m->n * sizeof(u64)));
    CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)3 * m->n * sizeof(u64), c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

int gx_mg_exchange(gx_ctx* c) {
    GX_TRY(require_live(c));
    MgState* m = mg_of(c);
    if (!m || !m->comm) return fail(c, GX_ERR_STATE, "gx_mg_exchange before gx_mg_init");
    if (c->finished) return fail(c, GX_ERR_STATE, "gx_mg_exchange after gx_finish");
    cudaSetDevice(c->cfg.device);
    const int n = m->n, me = m->rank;
    GX_TRY(sync_counters(c));
    GX_TRY(handle_spills(c));
    note_sync(c, c->h_ctr->distinct);
    const u64 head_cursor = c->h_ctr->head_cursor, store_cursor = c->h_ctr->store_cursor;
    ScopedPhase ph(c, PH_EXCHANGE);
    // ---- 1. bucket the read heads created since the last exchange
    const u64 new_heads = head_cursor - m->routed_heads_upto;
    const u64 new_store = store_cursor - m->routed_store_upto;
    for (int d = 0; d < n; ++d) {
        if (d == me) continue;
        GX_TRY(ensure(c, m->send_heads[d], (size_t)std::max<u64>(new_heads, 1) * c->ops->head_bytes));
        // both mates of a pair reference the same two packed sequences and may go to the same owner: 2x
        GX_TRY(ensure(c, m->send_store[d], (size_t)std::max<u64>(2 * new_store, 1)));
    }
    GX_TRY(mg_upload_ptrs(c));
    if (new_heads) {
        HeadRouteArgs ha{};
        ha.heads = c->heads.p; ha.first = m->routed_heads_upto; ha.n = new_heads;
        ha.store = (const uint8_t*)c->store.p;
        ha.n_ranks = (u32)n; ha.rank = (u32)me;
        ha.send_heads = (void* const*)((void**)m->ptr_table.p + 2 * n);
        ha.send_store = (uint8_t* const*)((void**)m->ptr_table.p + 3 * n);
        ha.send_head_count = (u64*)m->counts.p + n;
        ha.send_store_bytes = (u64*)m->counts.p + 2 * n;
        c->ops->route_heads(ha, c->stream);
        GX_TRY(check_launch(c, "route_heads"));
    }
    // ---- 2. everybody learns everybody's counts
    NCCL_TRY(c, ncclAllGather(m->counts.p, m->all_counts.p, (size_t)3 * n, ncclUint64, m->comm, c->stream));
    std::vector<u64> all((size_t)3 * n * n);
    CUDA_TRY(c, cudaMemcpyAsync(all.data(), m->all_counts.p, all.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    auto cnt = [&](int src, int kind, int dst) { return all[(size_t)src * 3 * n + (size_t)kind * n + dst]; };
    u64 recv_kmers = 0, recv_heads = 0, recv_store = 0;
    for (int s = 0; s < n; ++s) {
        if (s == me) continue;
        recv_kmers += cnt(s, 0, me); recv_heads += cnt(s, 1, me); recv_store += cnt(s, 2, me);
    }
    // ---- 3. room for what arrives
    GX_TRY(ensure(c, m->recv_keys, (size_t)std::max<u64>(recv_kmers, 1) * c->kw * sizeof(u64)));
    GX_TRY(ensure(c, m->recv_meta, (size_t)std::max<u64>(recv_kmers, 1) * sizeof(unsigned short)));
    GX_TRY(ensure(c, c->heads, (size_t)(head_cursor + recv_heads) * c->ops->head_bytes, (size_t)head_cursor * c->ops->head_bytes, true));
    GX_TRY(ensure(c, c->store, (size_t)(store_cursor + recv_store), (size_t)store_cursor));
    // ---- 4. all-to-all-v over NVLink
    NCCL_TRY(c, ncclGroupStart());
    {
        u64 rk = 0, rh = 0, rs = 0;
        for (int p = 0; p < n; ++p) {
            if (p == me) continue;
            const u64 sk = cnt(me, 0, p), sh = cnt(me, 1, p), ss = cnt(me, 2, p);
            if (sk) {
                NCCL_TRY(c, ncclSend(m->route_keys[p].p, (size_t)sk * c->kw, ncclUint64, p, m->comm, c->stream));
                NCCL_TRY(c, ncclSend(m->route_meta[p].p, (size_t)sk * 2, ncclUint8, p, m->comm, c->stream));
            }
            if (sh) NCCL_TRY(c, ncclSend(m->send_heads[p].p, (size_t)sh * c->ops->head_bytes, ncclUint8, p, m->comm, c->stream));
            if (ss) NCCL_TRY(c, ncclSend(m->send_store[p].p, (size_t)ss, ncclUint8, p, m->comm, c->stream));
            const u64 gk = cnt(p, 0, me), gh = cnt(p, 1, me), gs = cnt(p, 2, me);
            if (gk) {
                NCCL_TRY(c, ncclRecv((u64*)m->recv_keys.p + rk * c->kw, (size_t)gk * c->kw, ncclUint64, p, m->comm, c->stream));
                NCCL_TRY(c, ncclRecv((unsigned short*)m->recv_meta.p + rk, (size_t)gk * 2, ncclUint8, p, m->comm, c->stream));
            }
            if (gh) NCCL_TRY(c, ncclRecv((uint8_t*)c->heads.p + (size_t)(head_cursor + rh) * c->ops->head_bytes,
                                         (size_t)gh * c->ops->head_bytes, ncclUint8, p, m->comm, c->stream));
            if (gs) NCCL_TRY(c, ncclRecv((uint8_t*)c->store.p + store_cursor + rs, (size_t)gs, ncclUint8, p, m->comm, c->stream));
            rk += gk; rh += gh; rs += gs;
            m->exchanged += sk;
        }
    }
    NCCL_TRY(c, ncclGroupEnd());
    // ---- 5. fold what arrived
    {
        u64 distinct = c->h_ctr->distinct;
        for (u64 done = 0; done < recv_kmers;) {
            u64 room = 0;
            GX_TRY(reserve_room(c, distinct, 1, recv_kmers, &room));
            u64 take = recv_kmers - done;
            const u64 predicted = predict_new_keys(c, take);
            if (predicted > room) take = std::max<u64>(std::min<u64>(take, room), (u64)((double)take * (double)room / (double)predicted));
            c->ratio_pending_occ += take;
            c->ops->insert_records((const u64*)m->recv_keys.p + done * c->kw, (const unsigned short*)m->recv_meta.p + done, nullptr,
                                   take, c->table, c->capacity, c->d_ctr, c->stream);
            GX_TRY(check_launch(c, "insert_records"));
            done += take;
            if (done < recv_kmers) {
                GX_TRY(sync_counters(c));
                GX_TRY(handle_spills(c));
                distinct = c->h_ctr->distinct;
                note_sync(c, distinct);
            }
        }
    }
    {
        u64 rh = 0, rs = 0;
        for (int p = 0; p < n; ++p) {
            if (p == me) continue;
            const u64 gh = cnt(p, 1, me), gs = cnt(p, 2, me);
            if (gh) {
                c->ops->rebase_heads(c->heads.p, head_cursor + rh, gh, store_cursor + rs, c->stream);
                GX_TRY(check_launch(c, "rebase_heads"));
            }
            rh += gh; rs += gs;
        }
    }
    if (recv_heads || recv_store) {
        bump_cursors_kernel<<<1, 1, 0, c->stream>>>(c->d_ctr, recv_heads, recv_store);
        GX_TRY(check_launch(c, "bump_cursors"));
    }
    CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)3 * n * sizeof(u64), c->stream));
    m->routed_heads_upto = head_cursor + recv_heads;
    m->routed_store_upto = store_cursor + recv_store;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

}  // extern "C"
