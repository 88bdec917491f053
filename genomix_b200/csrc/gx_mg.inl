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
    u32 n_regions = 0;                            // table regions per rank, identical on every rank
    std::vector<Arena*> pending, free_arenas;     // chunks split but not yet exchanged / arenas to reuse
    std::vector<DevBuf> send_heads, send_store;   // per destination
    DevBuf ptr_table;      // device copy of the 2 pointer arrays: [2][n]
    DevBuf counts;         // device u64 [2n]: heads, store bytes per destination
    DevBuf round_vec;      // device u64 [V] + [n][V], V = 3n + 1: this round's counts, then everybody's
    DevBuf zero_seg;       // (n_regions + 1) zeros: what a rank without a chunk sends as its region index
    DevBuf inbox[5];       // receive areas: key words, masks, region index, Head<KW> records, packed sequences
    bool use_ipc = true;   // deliver with copy engines into CUDA-IPC mapped peer inboxes (else ncclSend/ncclRecv)
    std::vector<size_t> pub_cap;   // [n][5] inbox capacities every rank has published
    std::vector<void*> peer_ptr;   // [n][5] peers' inboxes mapped into this process
    DevBuf pub_dev, token;
    u64 routed_heads_upto = 0, routed_store_upto = 0;
    u64 exchanged = 0;
    cudaStream_t comm_stream = nullptr;   // NVLink traffic runs here
    cudaEvent_t ev_ready = nullptr, ev_arrived = nullptr;
    // The round whose records are on the wire: its upserts are launched by the NEXT gx_mg_exchange round (or gx_finish), so
    // that the transfer overlaps whatever the caller does in between -- normally the split of the next chunk.
    struct Rebase { u64 first, n, store_base; };
    struct InFlight {
        bool active = false;
        Arena* arena = nullptr;             // send blocks + own records of the round (kept until the upserts are done)
        std::vector<UpsertSrc> srcs;
        u64 total = 0;
        u64 recv_heads = 0, recv_store = 0, heads_at = 0, store_at = 0;
        std::vector<Rebase> rebase;
    } inflight;
    std::vector<PendingTimer> comm_timers;
    std::vector<void*> ptr_host;   // host copy of ptr_table
};

constexpr int MG_KINDS = 5;

MgState* mg_of(gx_ctx* c) { return reinterpret_cast<MgState*>(c->mg); }

// n_ranks > 1: the chunk's records are split by (owner, region) into a fresh arena and wait there for gx_mg_exchange
int mg_stage_chunk(gx_ctx* c, const uint8_t* d_text, size_t n, u64 n_lines, u64 chunk_occ) {
    MgState* m = mg_of(c);
    if (!m || !m->comm) return fail(c, GX_ERR_STATE, "n_ranks > 1 but gx_mg_init has not been called");
    Arena* ar = nullptr;
    if (!m->free_arenas.empty()) { ar = m->free_arenas.back(); m->free_arenas.pop_back(); }
    else ar = new Arena();
    m->pending.push_back(ar);
    return split_chunk(c, d_text, n, n_lines, chunk_occ, m->n_regions, *ar);
}

int mg_pending(gx_ctx* c, u64* pending) {
    MgState* m = mg_of(c);
    *pending = 0;
    if (!m) return fail(c, GX_ERR_STATE, "n_ranks > 1 but gx_mg_init has not been called");
    for (const Arena* ar : m->pending) if (ar) *pending += ar->occ;
    // heads created after the last exchange would also be lost
    if (c->h_ctr->head_cursor != m->routed_heads_upto) *pending += c->h_ctr->head_cursor - m->routed_heads_upto;
    return GX_OK;
}

u64 mg_exchanged(gx_ctx* c) { return mg_of(c) ? mg_of(c)->exchanged : 0; }

// Local (not collective): upsert the round that is on the wire -- own records and what the peers delivered -- and let the
// received read heads join the local arrays. Leaves both streams synchronised.
int mg_complete(gx_ctx* c) {
    MgState* m = mg_of(c);
    if (!m || !m->inflight.active) return GX_OK;
    MgState::InFlight& f = m->inflight;
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, m->ev_arrived, 0));
    GX_TRY(upsert_sources(c, f.srcs.data(), (u32)f.srcs.size(), m->n_regions, f.total, PH_XINSERT));
    if (f.recv_heads) {
        CUDA_TRY(c, cudaMemcpyAsync((uint8_t*)c->heads.p + (size_t)f.heads_at * c->ops->head_bytes, m->inbox[3].p,
                                    (size_t)f.recv_heads * c->ops->head_bytes, cudaMemcpyDeviceToDevice, c->stream));
        if (f.recv_store)
            CUDA_TRY(c, cudaMemcpyAsync((uint8_t*)c->store.p + f.store_at, m->inbox[4].p, (size_t)f.recv_store, cudaMemcpyDeviceToDevice, c->stream));
        for (const MgState::Rebase& r : f.rebase) {
            c->ops->rebase_heads(c->heads.p, r.first, r.n, r.store_base, c->stream);
            GX_TRY(check_launch(c, "rebase_heads"));
        }
    }
    CUDA_TRY(c, cudaStreamSynchronize(m->comm_stream));  // the round's send blocks are reusable from here on
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (auto& t : m->comm_timers) c->timers.push_back(t);
    m->comm_timers.clear();
    drain_timers(c);
    if (f.arena) m->free_arenas.push_back(f.arena);
    f = MgState::InFlight();
    return GX_OK;
}

void mg_destroy(gx_ctx* c) {
    MgState* m = mg_of(c);
    if (!m) return;
    if (m->comm_stream) cudaStreamSynchronize(m->comm_stream);
    if (m->inflight.arena) { release_arena(*m->inflight.arena); delete m->inflight.arena; }
    if (m->comm) nccl_api().CommDestroy(m->comm);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    if (m->ev_ready) cudaEventDestroy(m->ev_ready);
    if (m->ev_arrived) cudaEventDestroy(m->ev_arrived);
    for (auto* v : {&m->send_heads, &m->send_store})
        for (auto& b : *v) release(b);
    for (auto* v : {&m->pending, &m->free_arenas})
        for (Arena* ar : *v) if (ar) { release_arena(*ar); delete ar; }
    for (void* mp : m->peer_ptr) if (mp) cudaIpcCloseMemHandle(mp);
    release(m->ptr_table); release(m->counts); release(m->round_vec); release(m->zero_seg); release(m->pub_dev); release(m->token);
    for (auto& b : m->inbox) release(b);
    delete m;
    c->mg = nullptr;
}

int mg_reset(gx_ctx* c) {
    MgState* m = mg_of(c);
    if (!m) return GX_OK;
    if (m->comm_stream) CUDA_TRY(c, cudaStreamSynchronize(m->comm_stream));   // a round on the wire is dropped with the job
    for (auto& t : m->comm_timers) { c->event_pool.push_back(t.a); c->event_pool.push_back(t.b); }
    m->comm_timers.clear();
    if (m->inflight.arena) m->free_arenas.push_back(m->inflight.arena);
    m->inflight = MgState::InFlight();
    if (m->counts.p) CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)2 * m->n * sizeof(u64), c->stream));
    for (Arena* ar : m->pending) if (ar) m->free_arenas.push_back(ar);
    m->pending.clear();
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
    // table regions per rank: every rank must split with the same number, so it is fixed for the job from what all ranks
    // share -- the expected number of keys per rank if given, else a default that keeps regions L2-sized up to ~1 GB tables
    {
        const u32 max_regions = (u32)(SP_MAX_BUCKETS / m->n);
        u32 r = c->fixed_regions;
        if (!r) {
            const size_t table_bytes = (size_t)((double)c->cfg.expected_kmers / TARGET_LOAD) * c->ops->slot_bytes;
            r = c->cfg.expected_kmers ? (u32)std::max<size_t>(1, (table_bytes + REGION_BYTES - 1) / REGION_BYTES) : 32u;
        }
        m->n_regions = std::max(1u, std::min(r, max_regions));
    }
    m->send_heads.resize(m->n); m->send_store.resize(m->n);
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    NCCL_TRY(c, nccl_api().CommInitRank(&m->comm, m->n, id, m->rank));
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CUDA_TRY(c, cudaStreamCreateWithPriority(&m->comm_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(c, cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming));
    CUDA_TRY(c, cudaEventCreateWithFlags(&m->ev_arrived, cudaEventDisableTiming));
    const size_t V = (size_t)3 * m->n + 1;
    GX_TRY(ensure(c, m->ptr_table, (size_t)2 * m->n * sizeof(void*)));
    GX_TRY(ensure(c, m->counts, (size_t)2 * m->n * sizeof(u64)));
    GX_TRY(ensure(c, m->round_vec, V * (size_t)(m->n + 1) * sizeof(u64)));
    GX_TRY(ensure(c, m->zero_seg, (size_t)(2 * m->n_regions + 1) * sizeof(u64), 0, true));
    GX_TRY(ensure(c, m->token, 256, 0, true));
    m->pub_cap.assign((size_t)m->n * MG_KINDS, 0);
    m->peer_ptr.assign((size_t)m->n * MG_KINDS, nullptr);
    m->use_ipc = getenv("GENOMIX_GB_NO_IPC") == nullptr;
    CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)2 * m->n * sizeof(u64), c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return GX_OK;
}

// Collective. Ships every staged chunk's records to their owners, one round per staged chunk (ranks with fewer chunks take
// part with empty blocks); read heads travel in the first round. Pipelined: a round's transfer is started and the call goes
// on (or returns) without waiting for it; its upserts -- own and received records, region by region -- are launched by the
// next round, the next gx_mg_exchange or gx_finish. So with one gx_mg_exchange per pushed chunk the NVLink transfer of chunk
// i runs under the split of chunk i+1 (K1(i+1) || a2a(i) || K2(i-1)), and the send side never holds more
// than two chunks of records.
int gx_mg_exchange(gx_ctx* c) {
    GX_TRY(require_live(c));
    MgState* m = mg_of(c);
    if (!m || !m->comm) return fail(c, GX_ERR_STATE, "gx_mg_exchange before gx_mg_init");
    if (c->finished) return fail(c, GX_ERR_STATE, "gx_mg_exchange after gx_finish");
    cudaSetDevice(c->cfg.device);
    const int n = m->n, me = m->rank;
    const u32 R = m->n_regions;
    const size_t V = (size_t)3 * n + 1;
    GX_TRY(sync_counters_if_stale(c));   // normally current: the split of the last chunk ended with a counter sync
    GX_TRY(handle_spills(c));
    ScopedPhase ph(c, PH_EXCHANGE);
    auto unit_bytes = [&](int kind) -> size_t {
        switch (kind) {
            case 0: return (size_t)c->kw * sizeof(u64);
            case 1: return sizeof(unsigned short);
            case 2: return (size_t)(2 * R + 1) * sizeof(u64);   // the owner's table: R starts, block end, R counts
            case 3: return c->ops->head_bytes;
            default: return 1;
        }
    };
    const size_t n_local_rounds = m->pending.size();
    size_t n_rounds = 1;
    for (size_t round = 0; round < n_rounds; ++round) {
        // ---- 0. the round before this one (from this call or an earlier one): its records have arrived by now
        GX_TRY(mg_complete(c));
        Arena* ar = round < n_local_rounds ? m->pending[round] : nullptr;
        if (ar) m->pending[round] = nullptr;   // from here on the round owns it (mg_complete hands it back)
        // ---- 1. first round: bucket the read heads created since the last exchange
        std::vector<u64> head_counts((size_t)2 * n, 0);
        const u64 head_cursor = c->h_ctr->head_cursor, store_cursor = c->h_ctr->store_cursor;   // current after mg_complete / sync_counters
        if (round == 0) {
            const u64 new_heads = head_cursor - m->routed_heads_upto;
            const u64 new_store = store_cursor - m->routed_store_upto;
            for (int d = 0; d < n; ++d) {
                if (d == me) continue;
                GX_TRY(ensure(c, m->send_heads[d], (size_t)std::max<u64>(new_heads, 1) * c->ops->head_bytes));
                // both mates of a pair reference the same two packed sequences and may go to the same owner: 2x
                GX_TRY(ensure(c, m->send_store[d], (size_t)std::max<u64>(2 * new_store, 1)));
            }
            std::vector<void*>& h = m->ptr_host;   // lives in the state: the asynchronous copy below may read it after this scope
            h.assign((size_t)2 * n, nullptr);
            for (int d = 0; d < n; ++d) { h[d] = m->send_heads[d].p; h[(size_t)n + d] = m->send_store[d].p; }
            CUDA_TRY(c, cudaMemcpyAsync(m->ptr_table.p, h.data(), h.size() * sizeof(void*), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(c, cudaMemsetAsync(m->counts.p, 0, (size_t)2 * n * sizeof(u64), c->stream));
            if (new_heads) {
                HeadRouteArgs ha{};
                ha.heads = c->heads.p; ha.first = m->routed_heads_upto; ha.n = new_heads;
                ha.store = (const uint8_t*)c->store.p;
                ha.n_ranks = (u32)n; ha.rank = (u32)me;
                ha.send_heads = (void* const*)m->ptr_table.p;
                ha.send_store = (uint8_t* const*)((void**)m->ptr_table.p + n);
                ha.send_head_count = (u64*)m->counts.p;
                ha.send_store_bytes = (u64*)m->counts.p + n;
                c->ops->route_heads(ha, c->stream);
                GX_TRY(check_launch(c, "route_heads"));
            }
            CUDA_TRY(c, cudaMemcpyAsync(head_counts.data(), m->counts.p, head_counts.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        }
        // (the owner offsets of this round's arena came to the host with the split's sync)
        if (round == 0) CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // head_counts
        // ---- 2. everybody learns everybody's counts of this round. The collective also is the barrier that keeps a rank
        //         from pushing into an inbox its owner is still reading: every rank enters it after its own mg_complete.
        std::vector<u64> vec(V, 0), all(V * (size_t)n);
        vec[0] = n_local_rounds;
        for (int d = 0; d < n; ++d) {
            if (ar) vec[1 + d] = ar->owner_off[(size_t)d + 1] - ar->owner_off[d];
            if (round == 0) { vec[1 + (size_t)n + d] = head_counts[d]; vec[1 + (size_t)2 * n + d] = head_counts[(size_t)n + d]; }
        }
        CUDA_TRY(c, cudaMemcpyAsync(m->round_vec.p, vec.data(), V * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        NCCL_TRY(c, nccl_api().AllGather(m->round_vec.p, (u64*)m->round_vec.p + V, V, ncclUint64, m->comm, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(all.data(), (u64*)m->round_vec.p + V, all.size() * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (round == 0)
            for (int s2 = 0; s2 < n; ++s2) n_rounds = std::max<size_t>(n_rounds, (size_t)all[(size_t)s2 * V]);
        // count of `kind` units that src sends to dst this round
        auto count_kind = [&](int src, int kind, int dst) -> u64 {
            switch (kind) {
                case 0: case 1: return all[(size_t)src * V + 1 + dst];
                case 2: return 1;
                case 3: return all[(size_t)src * V + 1 + (size_t)n + dst];
                default: return all[(size_t)src * V + 1 + (size_t)2 * n + dst];
            }
        };
        auto need_bytes = [&](int dst, int kind) {
            u64 tot = 0;
            for (int s2 = 0; s2 < n; ++s2) if (s2 != dst) tot += count_kind(s2, kind, dst);
            return (size_t)tot * unit_bytes(kind);
        };
        // offset (in units) of source s's segment inside rank dst's inbox of a kind: sources are laid out in rank order
        auto seg_off = [&](int s2, int kind, int dst) {
            u64 off = 0;
            for (int q = 0; q < s2; ++q) if (q != dst) off += count_kind(q, kind, dst);
            return off;
        };
        u64 recv_heads = 0, recv_store = 0;
        for (int s2 = 0; s2 < n; ++s2) {
            if (s2 == me) continue;
            recv_heads += count_kind(s2, 3, me); recv_store += count_kind(s2, 4, me);
        }
        // ---- 3. inboxes. Every rank knows every rank's needs (the count matrix is global), so all ranks take the same
        //         decision about who has to (re)allocate: mappings of a growing inbox are closed everywhere, a barrier lets
        //         its owner reallocate, new CUDA-IPC handles are published and opened. Steady state: nothing to do.
        if (m->use_ipc) {
            bool any_grow = false;
            std::vector<char> grows(n, 0);
            for (int p = 0; p < n; ++p)
                for (int kind = 0; kind < MG_KINDS; ++kind)
                    if (need_bytes(p, kind) > m->pub_cap[(size_t)p * MG_KINDS + kind]) { grows[p] = 1; any_grow = true; }
            if (any_grow) {
                for (int p = 0; p < n; ++p) {
                    if (!grows[p] || p == me) continue;
                    for (int kind = 0; kind < MG_KINDS; ++kind) {
                        void*& mp = m->peer_ptr[(size_t)p * MG_KINDS + kind];
                        if (mp) { cudaIpcCloseMemHandle(mp); mp = nullptr; }
                    }
                }
                NCCL_TRY(c, nccl_api().AllReduce(m->token.p, m->token.p, 1, ncclFloat, ncclSum, m->comm, c->stream));  // everyone closed
                CUDA_TRY(c, cudaStreamSynchronize(c->stream));
                struct Pub { cudaIpcMemHandle_t h[MG_KINDS]; u64 cap[MG_KINDS]; };
                Pub mine;
                memset(&mine, 0, sizeof mine);
                for (int kind = 0; kind < MG_KINDS; ++kind) {
                    if (grows[me]) {
                        const size_t need = need_bytes(me, kind);
                        if (need > m->inbox[kind].cap) {
                            const size_t want = std::max<size_t>(need + need / 4 + 4096, 2 * m->inbox[kind].cap);
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
                    for (int kind = 0; kind < MG_KINDS; ++kind) {
                        m->pub_cap[(size_t)p * MG_KINDS + kind] = pubs[p].cap[kind];
                        if (p == me || !grows[p] || pubs[p].cap[kind] == 0) continue;
                        void* mp = nullptr;
                        if (cudaIpcOpenMemHandle(&mp, pubs[p].h[kind], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                            cudaGetLastError();
                            return fail(c, GX_ERR_CUDA, "cudaIpcOpenMemHandle failed for rank %d (set GENOMIX_GB_NO_IPC=1 to use NCCL send/recv)", p);
                        }
                        m->peer_ptr[(size_t)p * MG_KINDS + kind] = mp;
                    }
                }
            }
        } else {
            for (int kind = 0; kind < MG_KINDS; ++kind) GX_TRY(ensure(c, m->inbox[kind], std::max<size_t>(need_bytes(me, kind), 1)));
        }
        // ---- 4. room for the read heads that will arrive: reserved now, because the caller may push the next chunk (whose
        //         heads are appended behind them) before they are copied in
        const u64 heads_at = head_cursor, store_at = store_cursor;
        if (recv_heads) {
            GX_TRY(ensure(c, c->heads, (size_t)(heads_at + recv_heads) * c->ops->head_bytes, (size_t)heads_at * c->ops->head_bytes, true));
            GX_TRY(ensure(c, c->store, (size_t)(store_at + recv_store), (size_t)store_at));
            bump_cursors_kernel<<<1, 1, 0, c->stream>>>(c->d_ctr, recv_heads, recv_store);
            GX_TRY(check_launch(c, "bump_cursors"));
        }
        if (round == 0) {   // everything up to here (local + arriving) is routed or owned
            m->routed_heads_upto = heads_at + recv_heads;
            m->routed_store_upto = store_at + recv_store;
        }
        // ---- 5. all-to-all-v over NVLink on the communication stream. With CUDA IPC a delivery is a copy-engine push
        //         straight into the peer's inbox (no SMs involved), all pushes of the round are followed by one tiny
        //         all-reduce as arrival barrier; without, one grouped ncclSend/ncclRecv exchange.
        CUDA_TRY(c, cudaEventRecord(m->ev_ready, c->stream));
        CUDA_TRY(c, cudaStreamWaitEvent(m->comm_stream, m->ev_ready, 0));
        PendingTimer comm_t{PH_XCOMM, get_event(c), get_event(c)};
        cudaEventRecord(comm_t.a, m->comm_stream);
        auto src_ptr = [&](int to, int kind) -> const void* {
            switch (kind) {
                case 0: return ar ? (const void*)((const u64*)ar->keys.p + ar->owner_off[to] * c->kw) : nullptr;
                case 1: return ar ? (const void*)((const unsigned short*)ar->meta.p + ar->owner_off[to]) : nullptr;
                case 2: return ar ? (const void*)((const u64*)ar->tab.p + (size_t)to * (2 * R + 1)) : m->zero_seg.p;
                case 3: return m->send_heads[to].p;
                default: return m->send_store[to].p;
            }
        };
        if (!m->use_ipc) NCCL_TRY(c, nccl_api().GroupStart());
        for (int i = 1; i < n; ++i) {
            const int to = (me + i) % n, from = (me - i + n) % n;
            for (int kind = 0; kind < MG_KINDS; ++kind) {
                const size_t sb = (size_t)count_kind(me, kind, to) * unit_bytes(kind);
                if (m->use_ipc) {
                    if (!sb) continue;
                    uint8_t* dst = (uint8_t*)m->peer_ptr[(size_t)to * MG_KINDS + kind] + (size_t)seg_off(me, kind, to) * unit_bytes(kind);
                    CUDA_TRY(c, cudaMemcpyAsync(dst, src_ptr(to, kind), sb, cudaMemcpyDeviceToDevice, m->comm_stream));
                } else {
                    if (sb) NCCL_TRY(c, nccl_api().Send(src_ptr(to, kind), sb, ncclUint8, to, m->comm, m->comm_stream));
                    const size_t rb = (size_t)count_kind(from, kind, me) * unit_bytes(kind);
                    if (rb) NCCL_TRY(c, nccl_api().Recv((uint8_t*)m->inbox[kind].p + (size_t)seg_off(from, kind, me) * unit_bytes(kind), rb,
                                                        ncclUint8, from, m->comm, m->comm_stream));
                }
            }
            m->exchanged += count_kind(me, 0, to);
        }
        if (m->use_ipc) NCCL_TRY(c, nccl_api().AllReduce(m->token.p, m->token.p, 1, ncclFloat, ncclSum, m->comm, m->comm_stream));
        else NCCL_TRY(c, nccl_api().GroupEnd());
        CUDA_TRY(c, cudaEventRecord(m->ev_arrived, m->comm_stream));
        cudaEventRecord(comm_t.b, m->comm_stream);
        m->comm_timers.push_back(comm_t);   // resolved by mg_complete, once the communication stream has been synchronised
        // ---- 6. what mg_complete will upsert and append once the round has arrived
        MgState::InFlight& f = m->inflight;
        f.active = true;
        f.arena = ar;
        f.srcs.clear(); f.rebase.clear();
        f.total = 0;
        if (ar && vec[1 + me]) {
            const u64* tab = (const u64*)ar->tab.p + (size_t)me * (2 * R + 1);
            f.srcs.push_back(UpsertSrc{(const u64*)ar->keys.p, (const unsigned short*)ar->meta.p, tab, tab + R + 1, 0});
            f.total += vec[1 + me];
        }
        for (int s2 = 0; s2 < n; ++s2) {
            if (s2 == me || !count_kind(s2, 0, me)) continue;
            const u64 off = seg_off(s2, 0, me);
            const u64* tab = (const u64*)((const uint8_t*)m->inbox[2].p + (size_t)seg_off(s2, 2, me) * unit_bytes(2));
            f.srcs.push_back(UpsertSrc{(const u64*)m->inbox[0].p + off * c->kw, (const unsigned short*)m->inbox[1].p + off, tab, tab + R + 1, 1});
            f.total += count_kind(s2, 0, me);
        }
        f.recv_heads = recv_heads; f.recv_store = recv_store; f.heads_at = heads_at; f.store_at = store_at;
        for (int p = 0; p < n; ++p) {
            if (p == me) continue;
            const u64 gh = count_kind(p, 3, me);
            if (gh) f.rebase.push_back(MgState::Rebase{heads_at + seg_off(p, 3, me), gh, store_at + seg_off(p, 4, me)});
        }
        // the host copy of the cursors shows the reservation as well (what bump_cursors_kernel does on the device)
        if (recv_heads) { c->h_ctr->head_cursor += recv_heads; c->h_ctr->store_cursor += recv_store; }
    }
    m->pending.clear();   // every staged arena went through a round: the last one is in flight, the others are free again
    return GX_OK;
}

}  // extern "C"
