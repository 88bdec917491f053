// gx_mg.inl -- multi-GPU hash-partitioned exchange (included at the end of gx_api.cu).
//
// Replaces the reference's M:N hash-partitioning connector between the local and the global aggregator
// (MToNPartitioningMergingConnectorDescriptor with KmerPartitionComputerFactory, JobGenBuildBrujinGraph.java:
// 132-133; sender PartitionDataWriter.java:79-94 over Hyracks' TCP stack): every k-mer occurrence is owned by
// rank owner_of(hash(key)); the extract kernel inserts own keys directly and appends the others to per-owner
// send buckets; gx_mg_exchange() ships the buckets with one NCCL all-to-all-v (ncclSend/ncclRecv group over
// NVLink) and upserts what arrives. Read heads travel the same way with their packed sequences.
#include <dlfcn.h>
#include <nccl.h>   // types and constants only: the library is bound at run time (see NcclApi)

namespace {

// NCCL is resolved lazily with dlopen instead of being a link-time dependency: a single-GPU job never loads it, and a
// multi-GPU job binds to the NCCL already in the process (e.g. the one PyTorch ships, which must not be shadowed by a
// different libnccl.so.2 loaded earlier) or else to the system library.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl_api() {
    static NcclApi api;
    if (api.handle || !api.error.empty()) return api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (api.error.empty()) api.handle = h;
    return api;
}

#define NCCL_TRY(c, expr)                                                                        \
    do {                                                                                         \
        ncclResult_t r__ = (expr);                                                               \
        if (r__ != ncclSuccess)                                                                  \
            return fail(c, GX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(r__), __FILE__, __LINE__); \
    } while (0)

struct MgState {
    ncclComm_t comm = nullptr;
    int n = 1, rank = 0;
    std::vector<DevBuf> route_keys, route_meta, send_heads, send_store;  // per destination
    DevBuf ptr_table;      // device copy of the 4 pointer arrays: [4][n]
    DevBuf counts;         // device u64 [3n]: kmer records, heads, store bytes per destination
    DevBuf all_counts;     // device u64 [n][3n] after the all-gather
    DevBuf inbox[4];       // receive areas: key words, masks, Head<KW> records, packed sequences
    bool use_ipc = true;   // deliver with copy engines into CUDA-IPC mapped peer inboxes (else ncclSend/ncclRecv)
    std::vector<size_t> pub_cap;   // [n][4] inbox capacities every rank has published
    std::vector<void*> peer_ptr;   // [n][4] peers' inboxes mapped into this process
    DevBuf pub_dev, token;
    std::vector<u64> h_counts;  // host copy of counts (valid after a sync)
    u64 routed_heads_upto = 0, routed_store_upto = 0;
    u64 exchanged = 0;
    cudaStream_t comm_stream = nullptr;   // NCCL traffic runs here, overlapped with the upserts on the ctx stream
    cudaEvent_t ev_ready = nullptr;
    std::vector<cudaEvent_t> ev_step;
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
    if (m->comm_stream) cudaStreamSynchronize(m->comm_stream);
    if (m->comm) nccl_api().CommDestroy(m->comm);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    if (m->ev_ready) cudaEventDestroy(m->ev_ready);
    for (auto e : m->ev_step) if (e) cudaEventDestroy(e);
    for (auto* v : {&m->route_keys, &m->route_meta, &m->send_heads, &m->send_store})
        for (auto& b : *v) release(b);
    for (void* mp : m->peer_ptr) if (mp) cudaIpcCloseMemHandle(mp);
    release(m->ptr_table); release(m->counts); release(m->all_counts); release(m->pub_dev); release(m->token);
    for (auto& b : m->inbox) release(b);
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
    if (!nccl_api().handle || nccl_api().GetUniqueId(&id) != ncclSuccess) return GX_ERR_CUDA;
    memcpy(out_id, &id, 128);
    return GX_OK;
}

int gx_mg_init(gx_ctx* c, const uint8_t id_bytes[128]) {
    GX_TRY(require_live(c));
    if (c->cfg.n_ranks < 2) return fail(c, GX_ERR_INVALID, "gx_mg_init on a single-rank ctx");
    if (c->mg) return fail(c, GX_ERR_STATE, "gx_mg_init called twice");
    if (!nccl_api().handle) return fail(c, GX_ERR_CUDA, "NCCL unavailable: %s", nccl_api().error.c_str());
    cudaSetDevice(c->cfg.device);
    MgState* m = new MgState();
    c->mg = m;
    m->n = c->cfg.n_ranks;
    m->rank = c->cfg.rank;
    m->route_keys.resize(m->n); m->route_meta.resize(m->n); m->send_heads.resize(m->n); m->send_store.resize(m->n);
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    NCCL_TRY(c, nccl_api().CommInitRank(&m->comm, m->n, id, m->rank));
    {   // highest priority: NCCL's copy CTAs get the next free SM slots while the upsert kernel is running
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CUDA_TRY(c, cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(c, cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming));
    m->ev_step.resize(m->n);
    for (auto& e : m->ev_step) CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    GX_TRY(ensure(c, m->ptr_table, (size_t)4 * m->n * sizeof(void*)));
    GX_TRY(ensure(c, m->counts, (size_t)3 * m->n * sizeof(u64)));
    GX_TRY(ensure(c, m->all_counts, (size_t)3 * m->n * m->n * sizeof(u64)));
    GX_TRY(ensure(c, m->token, 256, 0, true));
    m->pub_cap.assign((size_t)m->n * 4, 0);
    m->peer_ptr.assign((size_t)m->n * 4, nullptr);
    m->use_ipc = getenv("GENOMIX_GB_NO_IPC") == nullptr;
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
    NCCL_TRY(c, nccl_api().AllGather(m->counts.p, m->all_counts.p, (size_t)3 * n, ncclUint64, m->comm, c->stream));
    std::vector<u64> all((size_t)3 * n * n);
    CUDA_TRY(c, cudaMemcpyAsync(all.data(), m->all_counts.p, all.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    auto cnt = [&](int src, int kind, int dst) { return all[(size_t)src * 3 * n + (size_t)kind * n + dst]; };
    u64 recv_kmers = 0, recv_heads = 0, recv_store = 0;
    for (int s = 0; s < n; ++s) {
        if (s == me) continue;
        recv_kmers += cnt(s, 0, me); recv_heads += cnt(s, 1, me); recv_store += cnt(s, 2, me);
    }
    // ---- 3. inboxes. Every rank knows every rank's needs (the count matrix is global), so all ranks take the same
    //         decision about who has to (re)allocate: mappings of a growing inbox are closed everywhere, a barrier lets
    //         its owner reallocate, new CUDA-IPC handles are published and opened. Steady state: nothing to do.
    auto unit_bytes = [&](int kind) -> size_t {
        return kind == 0 ? (size_t)c->kw * sizeof(u64) : kind == 1 ? sizeof(unsigned short) : kind == 2 ? c->ops->head_bytes : 1;
    };
    auto count_kind = [&](int src, int kind, int dst) { return cnt(src, kind == 0 || kind == 1 ? 0 : kind - 1, dst); };
    auto need_bytes = [&](int dst, int kind) {
        u64 tot = 0;
        for (int s2 = 0; s2 < n; ++s2) if (s2 != dst) tot += count_kind(s2, kind, dst);
        return (size_t)tot * unit_bytes(kind);
    };
    if (m->use_ipc) {
        bool any_grow = false;
        std::vector<char> grows(n, 0);
        for (int p = 0; p < n; ++p)
            for (int kind = 0; kind < 4; ++kind)
                if (need_bytes(p, kind) > m->pub_cap[(size_t)p * 4 + kind]) { grows[p] = 1; any_grow = true; }
        if (any_grow) {
            for (int p = 0; p < n; ++p) {
                if (!grows[p] || p == me) continue;
                for (int kind = 0; kind < 4; ++kind) {
                    void*& mp = m->peer_ptr[(size_t)p * 4 + kind];
                    if (mp) { cudaIpcCloseMemHandle(mp); mp = nullptr; }
                }
            }
            NCCL_TRY(c, nccl_api().AllReduce(m->token.p, m->token.p, 1, ncclFloat, ncclSum, m->comm, c->stream));  // everyone closed
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            struct Pub { cudaIpcMemHandle_t h[4]; u64 cap[4]; };
            Pub mine;
            memset(&mine, 0, sizeof mine);
            for (int kind = 0; kind < 4; ++kind) {
                if (grows[me]) {
                    const size_t want = std::max<size_t>(need_bytes(me, kind) + need_bytes(me, kind) / 4 + 4096, 2 * m->inbox[kind].cap);
                    if (need_bytes(me, kind) > m->inbox[kind].cap) {
                        release(m->inbox[kind]);
                        GX_TRY(ensure(c, m->inbox[kind], want));
                    }
                }
                mine.cap[kind] = m->inbox[kind].cap;
                if (m->inbox[kind].p && cudaIpcGetMemHandle(&mine.h[kind], m->inbox[kind].p) != cudaSuccess) {
                    cudaGetLastError();
                    return fail(c, GX_ERR_CUDA, "cudaIpcGetMemHandle failed (set GENOMIX_GB_NO_IPC=1 to use NCCL send/recv)");
                }
            }
            GX_TRY(ensure(c, m->pub_dev, sizeof(Pub) * (size_t)(n + 1)));
            CUDA_TRY(c, cudaMemcpyAsync(m->pub_dev.p, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
            NCCL_TRY(c, nccl_api().AllGather(m->pub_dev.p, (uint8_t*)m->pub_dev.p + sizeof(Pub), sizeof(Pub), ncclUint8, m->comm, c->stream));
            std::vector<Pub> pubs(n);
            CUDA_TRY(c, cudaMemcpyAsync(pubs.data(), (uint8_t*)m->pub_dev.p + sizeof(Pub), sizeof(Pub) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            for (int p = 0; p < n; ++p) {
                for (int kind = 0; kind < 4; ++kind) {
                    m->pub_cap[(size_t)p * 4 + kind] = pubs[p].cap[kind];
                    if (p == me || !grows[p] || pubs[p].cap[kind] == 0) continue;
                    void* mp = nullptr;
                    if (cudaIpcOpenMemHandle(&mp, pubs[p].h[kind], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                        cudaGetLastError();
                        return fail(c, GX_ERR_CUDA, "cudaIpcOpenMemHandle failed for rank %d (set GENOMIX_GB_NO_IPC=1 to use NCCL send/recv)", p);
                    }
                    m->peer_ptr[(size_t)p * 4 + kind] = mp;
                }
            }
        }
    } else {
        for (int kind = 0; kind < 4; ++kind) GX_TRY(ensure(c, m->inbox[kind], std::max<size_t>(need_bytes(me, kind), 1)));
    }
    GX_TRY(ensure(c, c->heads, (size_t)(head_cursor + recv_heads) * c->ops->head_bytes, (size_t)head_cursor * c->ops->head_bytes, true));
    GX_TRY(ensure(c, c->store, (size_t)(store_cursor + recv_store), (size_t)store_cursor));
    // offset (in units) of source s's segment inside rank dst's inbox of a kind: sources are laid out in rank order
    auto seg_off = [&](int s2, int kind, int dst) {
        u64 off = 0;
        for (int q = 0; q < s2; ++q) if (q != dst) off += count_kind(q, kind, dst);
        return off;
    };
    // ---- 4. all-to-all-v over NVLink in n-1 ring-shift steps on a communication stream: in step i every rank delivers
    //         to (me+i) and is delivered to by (me-i). With CUDA IPC the delivery is a copy-engine push straight into the
    //         peer's inbox (no SMs involved) followed by a tiny all-reduce as arrival barrier; without, ncclSend/ncclRecv.
    //         The segment that arrived in step i is upserted on the compute stream while step i+1 is on the wire.
    CUDA_TRY(c, cudaEventRecord(m->ev_ready, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(m->comm_stream, m->ev_ready, 0));
    PendingTimer comm_t{PH_XCOMM, get_event(c), get_event(c)};
    cudaEventRecord(comm_t.a, m->comm_stream);
    for (int i = 1; i < n; ++i) {
        const int to = (me + i) % n, from = (me - i + n) % n;
        const void* src[4] = {m->route_keys[to].p, m->route_meta[to].p, m->send_heads[to].p, m->send_store[to].p};
        if (m->use_ipc) {
            for (int kind = 0; kind < 4; ++kind) {
                const size_t bytes = (size_t)count_kind(me, kind, to) * unit_bytes(kind);
                if (!bytes) continue;
                uint8_t* dst = (uint8_t*)m->peer_ptr[(size_t)to * 4 + kind] + (size_t)seg_off(me, kind, to) * unit_bytes(kind);
                CUDA_TRY(c, cudaMemcpyAsync(dst, src[kind], bytes, cudaMemcpyDeviceToDevice, m->comm_stream));
            }
            NCCL_TRY(c, nccl_api().AllReduce(m->token.p, m->token.p, 1, ncclFloat, ncclSum, m->comm, m->comm_stream));
        } else {
            NCCL_TRY(c, nccl_api().GroupStart());
            for (int kind = 0; kind < 4; ++kind) {
                const size_t sb = (size_t)count_kind(me, kind, to) * unit_bytes(kind);
                if (sb) NCCL_TRY(c, nccl_api().Send(src[kind], sb, ncclUint8, to, m->comm, m->comm_stream));
                const size_t rb = (size_t)count_kind(from, kind, me) * unit_bytes(kind);
                if (rb) NCCL_TRY(c, nccl_api().Recv((uint8_t*)m->inbox[kind].p + (size_t)seg_off(from, kind, me) * unit_bytes(kind), rb, ncclUint8,
                                             from, m->comm, m->comm_stream));
            }
            NCCL_TRY(c, nccl_api().GroupEnd());
        }
        CUDA_TRY(c, cudaEventRecord(m->ev_step[i], m->comm_stream));
        m->exchanged += cnt(me, 0, to);
    }
    cudaEventRecord(comm_t.b, m->comm_stream);
    c->timers.push_back(comm_t);
    // ---- 5. fold what arrives, segment by segment
    {
        ScopedPhase phi(c, PH_XINSERT);
        u64 distinct = c->h_ctr->distinct;
        for (int i = 1; i < n; ++i) {
            const int from = (me - i + n) % n;
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream, m->ev_step[i], 0));
            const u64 gk = cnt(from, 0, me), base = seg_off(from, 0, me);
            for (u64 done = 0; done < gk;) {
                u64 room = 0;
                GX_TRY(reserve_room(c, distinct, 1, recv_kmers, &room));
                u64 take = gk - done;
                const u64 predicted = predict_new_keys(c, take);
                if (predicted > room) take = std::max<u64>(std::min<u64>(take, room), (u64)((double)take * (double)room / (double)predicted));
                c->ratio_pending_occ += take;
                c->ops->insert_records((const u64*)m->inbox[0].p + (base + done) * c->kw, (const unsigned short*)m->inbox[1].p + base + done,
                                       nullptr, take, c->table, c->capacity, c->d_ctr, c->stream);
                GX_TRY(check_launch(c, "insert_records"));
                done += take;
                if (done < gk) {
                    GX_TRY(sync_counters(c));
                    GX_TRY(handle_spills(c));
                    distinct = c->h_ctr->distinct;
                    note_sync(c, distinct);
                }
            }
        }
    }
    // ---- 6. received read heads and their sequences join the local arrays (all steps have been waited for above)
    if (recv_heads) {
        CUDA_TRY(c, cudaMemcpyAsync((uint8_t*)c->heads.p + (size_t)head_cursor * c->ops->head_bytes, m->inbox[2].p,
                                    (size_t)recv_heads * c->ops->head_bytes, cudaMemcpyDeviceToDevice, c->stream));
        if (recv_store)
            CUDA_TRY(c, cudaMemcpyAsync((uint8_t*)c->store.p + store_cursor, m->inbox[3].p, (size_t)recv_store, cudaMemcpyDeviceToDevice, c->stream));
        for (int p = 0; p < n; ++p) {
            if (p == me) continue;
            const u64 gh = cnt(p, 1, me);
            if (!gh) continue;
            c->ops->rebase_heads(c->heads.p, head_cursor + seg_off(p, 2, me), gh, store_cursor + seg_off(p, 3, me), c->stream);
            GX_TRY(check_launch(c, "rebase_heads"));
        }
        bump_cursors_kernel<<<1, 1, 0, c->stream>>>(c->d_ctr, recv_heads, recv_store);
        GX_TRY(check_launch(c, "bump_cursors"));
    }
    CUDA_TRY(c, cudaStreamSynchronize(m->comm_stream));  // send buckets are reusable from here on
    CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)3 * n * sizeof(u64), c->stream));
    m->routed_heads_upto = head_cursor + recv_heads;
    m->routed_store_upto = store_cursor + recv_store;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (m->use_ipc) {
        // nobody may reuse (overwrite) its send buckets or inbox before every rank has finished reading: arrival barriers
        // ordered the copies, this final one orders the end of the upserts that read the inboxes
        NCCL_TRY(c, nccl_api().AllReduce(m->token.p, m->token.p, 1, ncclFloat, ncclSum, m->comm, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return GX_OK;
}

}  // extern "C"
