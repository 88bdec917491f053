/*
 * gx_oracle.c -- CPU restatement of the Genomix graph-build job. TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The reference is 100 % Java and no JVM exists here, so this C twin of oracle/oracle.py restates the same
 * pipeline SHAPE the reference runs (JobGenBuildBrujinGraph.java:79-90):
 *
 *   parse lines -> one (Kmer, Node) tuple per k-mer occurrence, Node serialised   [ReadsKeyValueParserFactory.java:95-254]
 *   -> sort tuples by KmerPointable order                                          [KmerPointable.java:94-107; ExternalSortOperatorDescriptor]
 *   -> streaming group-by + aggregate (local)                                      [AggregateKmerAggregateFactory.java:93-144; PreclusteredGroupWriter.java:76-136]
 *   -> hash repartition with the Java 31-polynomial hash                           [KmerPartitionComputerFactory.java:28-52]
 *   -> merge + streaming group-by + aggregate (global)                             [same aggregator, JobGenBuildBrujinGraph.java:105-108]
 *   -> records `VKmer key | Node`                                                  [KmerNodePairSequenceWriterFactory.java:79-94; Node.java:408-427]
 *
 * with one worker thread per "partition" (the reference runs nNC x threadsPerMachine partition threads,
 * JobGen.java:65-70). The k-mer arithmetic is the reference's byte-wise code, not the word-parallel
 * identities the CUDA path uses: setFromStringBytes / setReversedFromStringBytes recomputed per position /
 * shiftKmerWithNextCode (Kmer.java:225-303), unsigned byte compare (hadoop WritableComparator.compareBytes).
 *
 * Parity pin: tests/test_oracle_golden.py checks this library against the reference's 9 golden files and
 * byte-for-byte against oracle/oracle.py. For k > 4 parity is pinned by code reading only (see oracle.py).
 *
 * Paths cited are relative to /root/reference/genomix/{genomix-hyracks,genomix-data}/src/main/java/edu/uci/ics/genomix/...
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ utils */
typedef struct {
    uint8_t* p;
    size_t len, cap;
} Buf;

static void buf_reserve(Buf* b, size_t extra) {
    if (b->len + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 4096;
    while (nc < b->len + extra) nc *= 2;
    b->p = (uint8_t*)realloc(b->p, nc);
    if (!b->p) { fprintf(stderr, "gx_oracle: out of memory\n"); abort(); }
    b->cap = nc;
}
static void buf_put(Buf* b, const void* src, size_t n) {
    buf_reserve(b, n);
    memcpy(b->p + b->len, src, n);
    b->len += n;
}
static void buf_put8(Buf* b, uint8_t v) { buf_put(b, &v, 1); }
static void buf_put32(Buf* b, uint32_t v) { /* Marshal.putInt, utils/Marshal.java:40-45: big-endian */
    uint8_t t[4] = {(uint8_t)(v >> 24), (uint8_t)(v >> 16), (uint8_t)(v >> 8), (uint8_t)v};
    buf_put(b, t, 4);
}
static void buf_put64(Buf* b, uint64_t v) {
    buf_put32(b, (uint32_t)(v >> 32));
    buf_put32(b, (uint32_t)v);
}
static uint32_t get32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static uint64_t get64(const uint8_t* p) { return ((uint64_t)get32(p) << 32) | get32(p + 4); }

/* ------------------------------------------------------------------------------------------------ GeneCode / Kmer */
static int code_from_symbol(uint8_t ch) { /* utils/GeneCode.java:29-50 (no default: anything else is A) */
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return 0;
}
static int byte_num_from_k(int k) { return k / 4 + (k % 4 != 0); } /* utils/KmerUtil.java:21-27 */

static void kmer_set_from_string(int k, const uint8_t* s, size_t slen, size_t start, uint8_t* out) { /* types/Kmer.java:225-242 */
    int nb = byte_num_from_k(k), bcount = nb - 1, bits = 0;
    uint8_t l = 0;
    for (size_t i = start; i < start + (size_t)k && i < slen; i++) {
        l |= (uint8_t)(code_from_symbol(s[i]) << bits);
        bits += 2;
        if (bits == 8) { out[bcount--] = l; l = 0; bits = 0; }
    }
    if (bcount >= 0) out[0] = l;
}
static void kmer_set_reversed_from_string(int k, const uint8_t* s, size_t slen, size_t start, uint8_t* out) { /* types/Kmer.java:253-272 */
    int nb = byte_num_from_k(k), bcount = nb - 1, bits = 0;
    uint8_t l = 0;
    for (long i = (long)start + k - 1; i >= (long)start && i < (long)slen; i--) {
        l |= (uint8_t)((3 - code_from_symbol(s[i])) << bits); /* getPairedCodeFromSymbol, GeneCode.java:52-61 */
        bits += 2;
        if (bits == 8) { out[bcount--] = l; l = 0; bits = 0; }
    }
    if (bcount >= 0) out[0] = l;
}
static void kmer_shift_with_next_code(int k, uint8_t* b, int c) { /* types/Kmer.java:292-303 + clearLeadBit :332-336 */
    int nb = byte_num_from_k(k);
    for (int i = nb - 1; i > 0; i--) {
        uint8_t in = b[i - 1] & 0x03;
        b[i] = (uint8_t)(((b[i] >> 2) & 0x3f) | (in << 6));
    }
    int pos = ((k - 1) % 4) << 1;
    b[0] = (uint8_t)(((b[0] >> 2) & 0x3f) | (c << pos));
    if (k % 4 != 0) b[0] &= (uint8_t)((1 << ((k % 4) << 1)) - 1);
}
static int compare_bytes(const uint8_t* a, const uint8_t* b, int n) { /* hadoop WritableComparator.compareBytes */
    for (int i = 0; i < n; i++)
        if (a[i] != b[i]) return (int)a[i] - (int)b[i];
    return 0;
}
static int pointable_compare(const uint8_t* a, const uint8_t* b, int n) { /* data/primitive/KmerPointable.java:94-107 */
    for (int i = n - 1; i >= 0; i--) {
        int c = (int)a[i] - (int)b[i];
        if (c) return c;
    }
    return 0;
}
static int java_partition(const uint8_t* key, int n, int n_parts) { /* data/primitive/KmerPartitionComputerFactory.java:28-52 */
    int32_t h = 1;
    for (int i = 0; i < n; i++) h = (int32_t)((uint32_t)31 * (uint32_t)h + (uint32_t)(int32_t)(int8_t)key[i]);
    if (h < 0) h = -(h + 1);
    return h % n_parts;
}

/* ------------------------------------------------------------------------------------------------ tuples */
typedef struct {
    const uint8_t* key;  /* nb bytes */
    const uint8_t* node; /* serialised Node */
    uint32_t node_len;
    uint32_t sender;     /* tie-breaks keep equal keys in input order (the TreeSet keeps the first equal ReadHeadInfo) */
    uint64_t seq;
} Tuple;

typedef struct {
    Tuple* t;
    size_t n, cap;
} TupleVec;

static void tv_push(TupleVec* v, Tuple t) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->t = (Tuple*)realloc(v->t, v->cap * sizeof(Tuple));
        if (!v->t) abort();
    }
    v->t[v->n++] = t;
}

/* append-only arena made of big blocks so that pointers stay valid */
typedef struct ArenaBlock {
    struct ArenaBlock* next;
    size_t used, cap;
    uint8_t data[];
} ArenaBlock;
typedef struct { ArenaBlock* head; } Arena;
static uint8_t* arena_alloc(Arena* a, size_t n) {
    if (!a->head || a->head->used + n > a->head->cap) {
        size_t cap = n > (8u << 20) ? n : (8u << 20);
        ArenaBlock* b = (ArenaBlock*)malloc(sizeof(ArenaBlock) + cap);
        if (!b) abort();
        b->next = a->head; b->used = 0; b->cap = cap;
        a->head = b;
    }
    uint8_t* p = a->head->data + a->head->used;
    a->head->used += n;
    return p;
}
static void arena_free(Arena* a) {
    while (a->head) { ArenaBlock* n = a->head->next; free(a->head); a->head = n; }
}

static __thread int g_nb; /* key bytes, for the qsort comparator */
static int tuple_cmp(const void* x, const void* y) {
    const Tuple* a = (const Tuple*)x; const Tuple* b = (const Tuple*)y;
    int c = pointable_compare(a->key, b->key, g_nb);
    if (c) return c;
    if (a->sender != b->sender) return a->sender < b->sender ? -1 : 1;
    return a->seq < b->seq ? -1 : (a->seq > b->seq);
}

/* ------------------------------------------------------------------------------------------------ Node (aggregation state) */
typedef struct {
    uint64_t value;      /* uuid */
    const uint8_t* raw;  /* serialised ReadHeadInfo */
    uint32_t raw_len;
} HeadRef;
typedef struct { HeadRef* h; size_t n, cap; } HeadSet;
typedef struct { const uint8_t** e; size_t n, cap; } EdgeList; /* pointers to serialised VKmers (4 + nb bytes) */
typedef struct {
    EdgeList edges[4];
    HeadSet heads[2]; /* unflipped, flipped */
    float coverage;
    int has_coverage;
} AggNode;

static void agg_reset(AggNode* a) {
    for (int i = 0; i < 4; i++) a->edges[i].n = 0;
    a->heads[0].n = a->heads[1].n = 0;
    a->coverage = 0; a->has_coverage = 0;
}
static void agg_free(AggNode* a) {
    for (int i = 0; i < 4; i++) free(a->edges[i].e);
    free(a->heads[0].h); free(a->heads[1].h);
}

/* ReadHeadInfo.compareTo (types/ReadHeadInfo.java:247-264): offset, library, mate, readId */
static int head_compare(uint64_t a, uint64_t b) {
    long oa = (long)(a >> 40), ob = (long)(b >> 40);
    if (oa & (1 << 23)) oa = -(oa & ((1 << 23) - 1));
    if (ob & (1 << 23)) ob = -(ob & ((1 << 23) - 1));
    if (oa != ob) return oa < ob ? -1 : 1;
    int la = (int)((a >> 36) & 0xf), lb = (int)((b >> 36) & 0xf);
    if (la != lb) return la < lb ? -1 : 1;
    int ma = (int)((a >> 35) & 1), mb = (int)((b >> 35) & 1);
    if (ma != mb) return ma < mb ? -1 : 1;
    uint64_t ra = a & ((1ull << 35) - 1), rb = b & ((1ull << 35) - 1);
    return ra < rb ? -1 : (ra > rb);
}
static void headset_add(HeadSet* s, HeadRef r) { /* TreeSet.add: an equal element already present is kept */
    size_t lo = 0, hi = s->n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        int c = head_compare(s->h[mid].value, r.value);
        if (c == 0) return;
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 4; s->h = (HeadRef*)realloc(s->h, s->cap * sizeof(HeadRef)); }
    memmove(s->h + lo + 1, s->h + lo, (s->n - lo) * sizeof(HeadRef));
    s->h[lo] = r;
    s->n++;
}
static void edgelist_union_add(EdgeList* l, const uint8_t* vk, int vk_len) { /* VKmerList.unionUpdate: set by VKmer bytes, types/VKmerList.java:117-133 */
    for (size_t i = 0; i < l->n; i++)
        if (memcmp(l->e[i], vk, vk_len) == 0) return;
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 4; l->e = (const uint8_t**)realloc(l->e, l->cap * sizeof(uint8_t*)); }
    l->e[l->n++] = vk;
}

/* Node.setAsCopy(byte[], int) (types/Node.java:336-360) folded with the aggregator's init/aggregate
 * (graph/dataflow/AggregateKmerAggregateFactory.java:93-125,128-144) */
static void agg_accumulate(AggNode* a, const uint8_t* p, int first) {
    uint8_t active = *p++;
    for (int et = 0; et < 4; et++) {
        if (!(active & (1 << et))) continue;
        uint32_t cnt = get32(p); p += 4;
        for (uint32_t i = 0; i < cnt; i++) {
            int vk_len = 4 + byte_num_from_k((int)get32(p));
            edgelist_union_add(&a->edges[et], p, vk_len);
            p += vk_len;
        }
    }
    for (int s = 0; s < 2; s++) {
        if (!(active & (1 << (4 + s)))) continue;
        p += 1; /* wholeBodyInStream == true */
        uint32_t cnt = get32(p); p += 4;
        for (uint32_t i = 0; i < cnt; i++) {
            const uint8_t* start = p;
            uint8_t flags = *p++;
            uint64_t value = get64(p); p += 8;
            p += 4 + byte_num_from_k((int)get32(p));
            if (flags & 1) p += 4 + byte_num_from_k((int)get32(p));
            HeadRef r = {value, start, (uint32_t)(p - start)};
            headset_add(&a->heads[s], r);
        }
    }
    /* bit 6 (internal kmer) is never produced by graph build */
    if (active & (1 << 7)) {
        uint32_t u = get32(p);
        float f; memcpy(&f, &u, 4);
        if (first || !a->has_coverage) a->coverage = f; else a->coverage = a->coverage + f; /* float add, :124,144 */
        a->has_coverage = 1;
    }
}

/* Node.write (types/Node.java:408-427) + getActiveFields (:466-487) */
static void agg_write(const AggNode* a, Buf* out) {
    uint8_t active = 0;
    for (int et = 0; et < 4; et++) if (a->edges[et].n) active |= (uint8_t)(1 << et);
    if (a->heads[0].n) active |= 1 << 4;
    if (a->heads[1].n) active |= 1 << 5;
    if (a->has_coverage) active |= 1 << 7;
    buf_put8(out, active);
    for (int et = 0; et < 4; et++) {
        if (!a->edges[et].n) continue;
        buf_put32(out, (uint32_t)a->edges[et].n);
        for (size_t i = 0; i < a->edges[et].n; i++) {
            const uint8_t* vk = a->edges[et].e[i];
            buf_put(out, vk, 4 + byte_num_from_k((int)get32(vk)));
        }
    }
    for (int s = 0; s < 2; s++) {
        if (!a->heads[s].n) continue;
        buf_put8(out, 1); /* ExternalableTreeSet.write with forceWriteEntireBody(true), types/ExternalableTreeSet.java:236-253 */
        buf_put32(out, (uint32_t)a->heads[s].n);
        for (size_t i = 0; i < a->heads[s].n; i++) buf_put(out, a->heads[s].h[i].raw, a->heads[s].h[i].raw_len);
    }
    if (a->has_coverage) {
        uint32_t u; memcpy(&u, &a->coverage, 4);
        buf_put32(out, u);
    }
}

/* ------------------------------------------------------------------------------------------------ job */
enum { GXO_OK = 0, GXO_FORMAT = -4, GXO_NUMBER = -5, GXO_TOO_SHORT = -6, GXO_READID = -7 };

typedef struct Job Job;
typedef struct {
    Job* job;
    int tid;
    size_t begin, end;       /* byte range of whole lines */
    uint64_t first_line;
    Arena arena;
    TupleVec tuples;         /* phase 1: per-occurrence tuples */
    TupleVec* outbox;        /* [n_parts] locally aggregated tuples per destination */
    Buf records;             /* phase 2 output of partition tid */
    uint64_t n_nodes;
    int err; uint64_t err_line;
    uint64_t reads, occurrences, lines;
} Worker;

struct Job {
    const uint8_t* text; size_t n;
    int k, nb, n_threads;
    Worker* w;
    pthread_barrier_t barrier;
};

static void set_error(Worker* w, int code, uint64_t line) {
    if (!w->err || line < w->err_line) { w->err = code; w->err_line = line; }
}

/* serialise a per-occurrence Node the way Node.marshalToByteArray does (Node.java:329-334), then InsertToFrame
 * (ReadsKeyValueParserFactory.java:235-254): here the "frame" is the worker's arena */
typedef struct {
    int edge_type[2]; uint8_t edge_key[2][32 + 4]; int n_edges;
    int has_head, head_flipped;
    Buf head_raw;
} OccNode;

static void emit_tuple(Worker* w, const uint8_t* key, const OccNode* n) {
    const int nb = w->job->nb, k = w->job->k;
    Buf b = {0};
    uint8_t active = 1 << 7;
    for (int i = 0; i < n->n_edges; i++) active |= (uint8_t)(1 << n->edge_type[i]);
    if (n->has_head) active |= (uint8_t)(1 << (n->head_flipped ? 5 : 4));
    buf_put8(&b, active);
    for (int et = 0; et < 4; et++) {
        int cnt = 0;
        for (int i = 0; i < n->n_edges; i++) cnt += n->edge_type[i] == et;
        if (!cnt) continue;
        buf_put32(&b, (uint32_t)cnt);
        for (int i = 0; i < n->n_edges; i++) {
            if (n->edge_type[i] != et) continue;
            buf_put32(&b, (uint32_t)k);
            buf_put(&b, n->edge_key[i], nb);
        }
    }
    if (n->has_head) {
        buf_put8(&b, 1);
        buf_put32(&b, 1);
        buf_put(&b, n->head_raw.p, n->head_raw.len);
    }
    float one = 1.0f; uint32_t u; memcpy(&u, &one, 4);
    buf_put32(&b, u);
    uint8_t* dst = arena_alloc(&w->arena, (size_t)nb + b.len);
    memcpy(dst, key, nb);
    memcpy(dst + nb, b.p, b.len);
    Tuple t = {dst, dst + nb, (uint32_t)b.len, (uint32_t)w->tid, w->tuples.n};
    tv_push(&w->tuples, t);
    free(b.p);
}

static void put_vkmer_from_string(Buf* b, const uint8_t* s, size_t len) { /* VKmer.setAsCopy(String), types/VKmer.java:140-142 */
    int nb = byte_num_from_k((int)len);
    buf_put32(b, (uint32_t)len);
    buf_reserve(b, nb);
    memset(b->p + b->len, 0, nb);
    kmer_set_from_string((int)len, s, len, 0, b->p + b->len);
    b->len += nb;
}

/* ReadsKeyValueParserFactory.SplitReads (:150-196) with setEdgesForCurAndNext (:209-233) and writeToFrame (:198-207) */
static int split_reads(Worker* w, uint64_t line_no, int mate, uint64_t read_id, const uint8_t* letters, size_t len,
                       const uint8_t* mate_letters, size_t mate_len, int has_mate_field) {
    const int k = w->job->k, nb = w->job->nb;
    if ((size_t)k >= len) { set_error(w, GXO_TOO_SHORT, line_no); return -1; }
    uint8_t cur_f[32], cur_r[32], nxt_f[32], nxt_r[32];
    memset(cur_f, 0, sizeof cur_f); memset(cur_r, 0, sizeof cur_r);
    kmer_set_from_string(k, letters, len, 0, cur_f);
    kmer_set_reversed_from_string(k, letters, len, 0, cur_r);
    int cur_dir = compare_bytes(cur_f, cur_r, nb) <= 0 ? 0 : 1; /* 0 FORWARD, 1 REVERSE */
    OccNode cur; memset(&cur, 0, sizeof cur);
    OccNode nxt; memset(&nxt, 0, sizeof nxt);
    /* read head on the first k-mer: unflipped offset 0, flipped offset K-1 (:165-170); library always 0 (:98-106) */
    cur.has_head = 1; cur.head_flipped = cur_dir;
    {
        uint64_t off = cur_dir ? (uint64_t)(k - 1) : 0;
        uint64_t uuid = (off << 40) + ((uint64_t)mate << 35) + read_id; /* ReadHeadInfo.makeUUID :100-127 */
        int mate_flag = has_mate_field && mate_len > 0;
        buf_put8(&cur.head_raw, (uint8_t)mate_flag); /* ReadHeadInfo.write :205-212 */
        buf_put64(&cur.head_raw, uuid);
        put_vkmer_from_string(&cur.head_raw, letters, len);
        if (mate_flag) put_vkmer_from_string(&cur.head_raw, mate_letters, mate_len);
    }
    memcpy(nxt_f, cur_f, sizeof cur_f);
    for (size_t i = (size_t)k; i < len; i++) {
        kmer_shift_with_next_code(k, nxt_f, code_from_symbol(letters[i]));
        memset(nxt_r, 0, sizeof nxt_r);
        kmer_set_reversed_from_string(k, letters, len, i - k + 1, nxt_r);
        int nxt_dir = compare_bytes(nxt_f, nxt_r, nb) <= 0 ? 0 : 1;
        /* EDGETYPE FF=0 FR=1 RF=2 RR=3 (types/EDGETYPE.java:6-9) */
        if (!cur_dir && !nxt_dir) {
            cur.edge_type[cur.n_edges] = 0; memcpy(cur.edge_key[cur.n_edges++], nxt_f, nb);
            nxt.edge_type[nxt.n_edges] = 3; memcpy(nxt.edge_key[nxt.n_edges++], cur_f, nb);
        } else if (!cur_dir && nxt_dir) {
            cur.edge_type[cur.n_edges] = 1; memcpy(cur.edge_key[cur.n_edges++], nxt_r, nb);
            nxt.edge_type[nxt.n_edges] = 1; memcpy(nxt.edge_key[nxt.n_edges++], cur_f, nb);
        } else if (cur_dir && !nxt_dir) {
            cur.edge_type[cur.n_edges] = 2; memcpy(cur.edge_key[cur.n_edges++], nxt_f, nb);
            nxt.edge_type[nxt.n_edges] = 2; memcpy(nxt.edge_key[nxt.n_edges++], cur_r, nb);
        } else {
            cur.edge_type[cur.n_edges] = 3; memcpy(cur.edge_key[cur.n_edges++], nxt_r, nb);
            nxt.edge_type[nxt.n_edges] = 0; memcpy(nxt.edge_key[nxt.n_edges++], cur_r, nb);
        }
        emit_tuple(w, cur_dir ? cur_r : cur_f, &cur);
        w->occurrences++;
        free(cur.head_raw.p);
        memcpy(cur_f, nxt_f, sizeof cur_f); memcpy(cur_r, nxt_r, sizeof cur_r);
        cur = nxt; cur_dir = nxt_dir;
        memset(&nxt, 0, sizeof nxt);
    }
    emit_tuple(w, cur_dir ? cur_r : cur_f, &cur);
    w->occurrences++;
    free(cur.head_raw.p);
    w->reads++;
    return 0;
}

static int is_gene(const uint8_t* s, size_t n) { /* Pattern "[ACGTacgt]+" full match (:125-128) */
    if (!n) return 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = s[i] | 0x20;
        if (c != 'a' && c != 'c' && c != 'g' && c != 't') return 0;
    }
    return 1;
}

/* ReadsKeyValueParserFactory.parse (:95-148) */
static int parse_line(Worker* w, uint64_t line_no, const uint8_t* s, size_t n) {
    /* String.split("\t"): trailing empty strings removed */
    const uint8_t* fs[4]; size_t fl[4];
    int nf_all = 0, last_nonempty = -1;
    size_t st = 0;
    for (size_t i = 0; i <= n; i++) {
        if (i == n || s[i] == '\t') {
            if (i > st) last_nonempty = nf_all;
            if (nf_all < 4) { fs[nf_all] = s + st; fl[nf_all] = i - st; }
            nf_all++;
            st = i + 1;
        }
    }
    int nf = last_nonempty + 1;
    if (nf != 2 && nf != 3) { set_error(w, GXO_FORMAT, line_no); return -1; }
    /* Long.parseLong */
    const uint8_t* p = fs[0]; size_t l = fl[0];
    int neg = 0;
    if (l && (p[0] == '-' || p[0] == '+')) { neg = p[0] == '-'; p++; l--; }
    if (!l) { set_error(w, GXO_NUMBER, line_no); return -1; }
    uint64_t mag = 0;
    for (size_t i = 0; i < l; i++) {
        if (p[i] < '0' || p[i] > '9') { set_error(w, GXO_NUMBER, line_no); return -1; }
        uint64_t d = (uint64_t)(p[i] - '0');
        if (mag > (0x8000000000000000ull - d) / 10) { set_error(w, GXO_NUMBER, line_no); return -1; }
        mag = mag * 10 + d;
    }
    if (!neg && mag > 0x7fffffffffffffffull) { set_error(w, GXO_NUMBER, line_no); return -1; }
    const uint8_t* m0 = fs[1]; size_t l0 = fl[1];
    const uint8_t* m1 = nf == 3 ? fs[2] : NULL; size_t l1 = nf == 3 ? fl[2] : 0;
    int id_bad = neg || mag >= (1ull << 29); /* makeUUID guard, ReadHeadInfo.java:110-113 */
    if (is_gene(m0, l0)) {
        if (id_bad) { set_error(w, GXO_READID, line_no); return -1; }
        if (split_reads(w, line_no, 0, mag, m0, l0, m1, l1, nf == 3)) return -1;
    }
    if (nf == 3 && is_gene(m1, l1)) {
        if (id_bad) { set_error(w, GXO_READID, line_no); return -1; }
        /* :140-145 -- mate sequence = raw mate-0 text even when it failed the regex */
        if (split_reads(w, line_no, 1, mag, m1, l1, m0, l0, 1)) return -1;
    }
    return 0;
}

/* streaming group-by over sorted tuples (PreclusteredGroupWriter.java:76-136) */
static void group_and_emit(Worker* w, TupleVec* in, void (*sink)(Worker*, const uint8_t* key, Buf* node, void* arg), void* arg) {
    const int nb = w->job->nb;
    AggNode agg; memset(&agg, 0, sizeof agg);
    Buf node = {0};
    size_t i = 0;
    while (i < in->n) {
        size_t j = i;
        agg_reset(&agg);
        while (j < in->n && pointable_compare(in->t[i].key, in->t[j].key, nb) == 0) {
            agg_accumulate(&agg, in->t[j].node, j == i);
            j++;
        }
        node.len = 0;
        agg_write(&agg, &node);
        sink(w, in->t[i].key, &node, arg);
        i = j;
    }
    free(node.p);
    agg_free(&agg);
}

static void sink_outbox(Worker* w, const uint8_t* key, Buf* node, void* arg) {
    (void)arg;
    const int nb = w->job->nb;
    int part = java_partition(key, nb, w->job->n_threads);
    uint8_t* dst = arena_alloc(&w->arena, (size_t)nb + node->len);
    memcpy(dst, key, nb);
    memcpy(dst + nb, node->p, node->len);
    Tuple t = {dst, dst + nb, (uint32_t)node->len, (uint32_t)w->tid, w->outbox[part].n};
    tv_push(&w->outbox[part], t);
}

static void sink_records(Worker* w, const uint8_t* key, Buf* node, void* arg) {
    (void)arg;
    const int nb = w->job->nb, k = w->job->k;
    /* SequenceFile record: recordLength, keyLength, VKmer.write (VKmer.java:389-391), Node.write */
    buf_put32(&w->records, (uint32_t)(4 + nb + node->len));
    buf_put32(&w->records, (uint32_t)(4 + nb));
    buf_put32(&w->records, (uint32_t)k);
    buf_put(&w->records, key, nb);
    buf_put(&w->records, node->p, node->len);
    w->n_nodes++;
}

static void* worker_main(void* arg) {
    Worker* w = (Worker*)arg;
    Job* job = w->job;
    g_nb = job->nb;
    /* phase 1: parse my split, sort, local aggregate, route */
    size_t pos = w->begin;
    uint64_t line_no = w->first_line;
    while (pos < w->end) {
        /* hadoop LineReader.readLine: a record ends at \n, \r or \r\n */
        size_t e = pos;
        while (e < w->end && job->text[e] != '\n' && job->text[e] != '\r') e++;
        w->lines++;
        if (parse_line(w, line_no, job->text + pos, e - pos)) break;
        line_no++;
        pos = e + 1;
        if (e < w->end && job->text[e] == '\r' && pos < job->n && job->text[pos] == '\n') pos++;
    }
    if (!w->err) {
        qsort(w->tuples.t, w->tuples.n, sizeof(Tuple), tuple_cmp);
        group_and_emit(w, &w->tuples, sink_outbox, NULL);
    }
    free(w->tuples.t); w->tuples.t = NULL; w->tuples.n = w->tuples.cap = 0;
    pthread_barrier_wait(&job->barrier);
    /* phase 2: I am the receiving partition `tid`: merge what every sender routed to me, aggregate again */
    int any_err = 0;
    for (int s = 0; s < job->n_threads; s++) any_err |= job->w[s].err;
    if (!any_err) {
        TupleVec in = {0};
        for (int s = 0; s < job->n_threads; s++)
            for (size_t i = 0; i < job->w[s].outbox[w->tid].n; i++) tv_push(&in, job->w[s].outbox[w->tid].t[i]);
        qsort(in.t, in.n, sizeof(Tuple), tuple_cmp);
        group_and_emit(w, &in, sink_records, NULL);
        free(in.t);
    }
    return NULL;
}

/* ------------------------------------------------------------------------------------------------ C API (ctypes) */
typedef struct {
    uint64_t lines, reads, occurrences, nodes;
    int32_t err; uint64_t err_line;
} gxo_stats;

/* Build the graph of `text` (readid lines) with k-mer length k on n_threads partition threads.
 * On success *out (malloc'ed, release with gxo_free) holds the concatenated record streams of all partitions. */
int gxo_build(const uint8_t* text, size_t n, int k, int n_threads, uint8_t** out, size_t* out_len, gxo_stats* stats) {
    if (k < 1 || k > 128 || n_threads < 1) return -1;
    Job job; memset(&job, 0, sizeof job);
    job.text = text; job.n = n; job.k = k; job.nb = byte_num_from_k(k); job.n_threads = n_threads;
    job.w = (Worker*)calloc((size_t)n_threads, sizeof(Worker));
    pthread_barrier_init(&job.barrier, NULL, (unsigned)n_threads);
    /* splits: equal byte ranges moved forward to the next line start (like HDFS splits + LineRecordReader) */
    size_t prev = 0;
    uint64_t line_base = 0;
    for (int t = 0; t < n_threads; t++) {
        size_t end = (t == n_threads - 1) ? n : (n * (size_t)(t + 1)) / (size_t)n_threads;
        if (end < prev) end = prev;
        if (t != n_threads - 1 && end > 0 && end < n) {
            const uint8_t* nl = (const uint8_t*)memchr(text + end - 1, '\n', n - end + 1);
            end = nl ? (size_t)(nl - text) + 1 : n;
        }
        Worker* w = &job.w[t];
        w->job = &job; w->tid = t; w->begin = prev; w->end = end; w->first_line = line_base;
        w->outbox = (TupleVec*)calloc((size_t)n_threads, sizeof(TupleVec));
        for (size_t i = prev; i < end; i++)
            line_base += text[i] == '\n' || (text[i] == '\r' && !(i + 1 < n && text[i + 1] == '\n'));
        if (end > prev && text[end - 1] != '\n' && text[end - 1] != '\r') line_base++;
        prev = end;
    }
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, worker_main, &job.w[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
    gxo_stats st; memset(&st, 0, sizeof st);
    size_t total = 0;
    for (int t = 0; t < n_threads; t++) {
        Worker* w = &job.w[t];
        st.lines += w->lines; st.reads += w->reads; st.occurrences += w->occurrences; st.nodes += w->n_nodes;
        if (w->err && (!st.err || w->err_line < st.err_line)) { st.err = w->err; st.err_line = w->err_line; }
        total += w->records.len;
    }
    uint8_t* res = NULL;
    if (!st.err) {
        res = (uint8_t*)malloc(total ? total : 1);
        size_t off = 0;
        for (int t = 0; t < n_threads; t++) { memcpy(res + off, job.w[t].records.p, job.w[t].records.len); off += job.w[t].records.len; }
    }
    for (int t = 0; t < n_threads; t++) {
        Worker* w = &job.w[t];
        for (int p = 0; p < n_threads; p++) free(w->outbox[p].t);
        free(w->outbox); free(w->records.p); arena_free(&w->arena);
    }
    free(job.w);
    pthread_barrier_destroy(&job.barrier);
    if (stats) *stats = st;
    if (st.err) { *out = NULL; *out_len = 0; return st.err; }
    *out = res; *out_len = total;
    return GXO_OK;
}

void gxo_free(uint8_t* p) { free(p); }

int gxo_java_partition(const uint8_t* key, int n, int n_parts) { return java_partition(key, n, n_parts); }

/* ------------------------------------------------------------------------------------------------ canonical checksum
 * Order-independent fingerprint of a record stream (recordLength | keyLength | VKmer | Node ...): the sum over records
 * of a 64-bit hash of (key, per-type SET of neighbour VKmers, unflipped/flipped read-head lists, coverage). Two streams
 * have the same fingerprint iff (up to 2^-64 collisions) they hold the same records after canonical sorting -- the
 * comparison rule of the reference's tests (utils/TestUtils.java:67-181: record order and VKmerList order are free). */
static uint64_t fp_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
static uint64_t fp_bytes(const uint8_t* p, size_t n, uint64_t seed) {
    uint64_t h = seed ^ (0x9e3779b97f4a7c15ull * (n + 1));
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = fp_mix(h ^ w) + 0x9e3779b97f4a7c15ull; }
    uint64_t w = 0;
    if (i < n) memcpy(&w, p + i, n - i);
    return fp_mix(h ^ w ^ ((uint64_t)(n - i) << 56));
}

typedef struct { uint64_t sum, xor_, records, heads, edges; double coverage_total; int32_t bad; } gxo_fingerprint;

int gxo_canonical_fingerprint(const uint8_t* stream, size_t n, gxo_fingerprint* out) {
    gxo_fingerprint f; memset(&f, 0, sizeof f);
    size_t off = 0;
    while (off + 8 <= n) {
        uint32_t rec_len = get32(stream + off), key_len = get32(stream + off + 4);
        if (off + 8 + rec_len > n || key_len > rec_len) { f.bad = 1; break; }
        const uint8_t* key = stream + off + 8;
        const uint8_t* p = key + key_len;
        const uint8_t* end = stream + off + 8 + rec_len;
        uint64_t h = fp_bytes(key, key_len, 1);
        uint8_t active = *p++;
        h = fp_mix(h ^ active);
        for (int et = 0; et < 4; et++) {
            if (!(active & (1 << et))) continue;
            uint32_t cnt = get32(p); p += 4;
            uint64_t set_sum = 0;
            for (uint32_t i = 0; i < cnt; i++) {
                int len = 4 + byte_num_from_k((int)get32(p));
                set_sum += fp_bytes(p, (size_t)len, 100 + (uint64_t)et);   /* order-free inside the list */
                p += len;
            }
            h = fp_mix(h ^ set_sum ^ ((uint64_t)cnt << 40));
            f.edges += cnt;
        }
        for (int s = 0; s < 2; s++) {
            if (!(active & (1 << (4 + s)))) continue;
            const uint8_t* start = p;
            p += 1;
            uint32_t cnt = get32(p); p += 4;
            for (uint32_t i = 0; i < cnt; i++) {
                uint8_t flags = *p++;
                p += 8;
                p += 4 + byte_num_from_k((int)get32(p));
                if (flags & 1) p += 4 + byte_num_from_k((int)get32(p));
            }
            h = fp_mix(h ^ fp_bytes(start, (size_t)(p - start), 200 + (uint64_t)s));  /* TreeSet order is part of the format */
            f.heads += cnt;
        }
        if (active & (1 << 6)) { p += 4 + byte_num_from_k((int)get32(p)); }
        if (active & (1 << 7)) {
            uint32_t u = get32(p); float c; memcpy(&c, &u, 4);
            f.coverage_total += c;
            h = fp_mix(h ^ u);
            p += 4;
        }
        if (p != end) { f.bad = 2; break; }
        f.sum += h; f.xor_ ^= h; f.records++;
        off += 8 + rec_len;
    }
    if (off != n && !f.bad) f.bad = 3;
    *out = f;
    return f.bad;
}

/* Edge symmetry: every adjacency is recorded on both nodes (setEdgesForCurAndNext, ReadsKeyValueParserFactory.java:209-233:
 * cur.edges[t] += next, next.edges[mirror(t)] += cur), so the signed sum over all (node X, type t, neighbour Y) of
 * E(X,t,Y) - E(Y,mirror(t),X) must vanish. Returns that sum (0 = symmetric, w.h.p.) and the edge count. */
static uint64_t edge_hash(const uint8_t* x, const uint8_t* y, int vk_len, int t) {
    return fp_mix(fp_bytes(x, (size_t)vk_len, 7) * 0x9e3779b97f4a7c15ull + fp_bytes(y, (size_t)vk_len, 11) + (uint64_t)t * 0xd6e8feb86659fd93ull);
}
int gxo_edge_symmetry(const uint8_t* stream, size_t n, uint64_t* signed_sum, uint64_t* n_edges) {
    static const int mirror[4] = {3, 1, 2, 0}; /* EDGETYPE.mirror, types/EDGETYPE.java:44-58 */
    uint64_t acc = 0, edges = 0;
    size_t off = 0;
    while (off + 8 <= n) {
        uint32_t rec_len = get32(stream + off), key_len = get32(stream + off + 4);
        if (off + 8 + rec_len > n) return 1;
        const uint8_t* key = stream + off + 8;
        const uint8_t* p = key + key_len;
        uint8_t active = *p++;
        for (int et = 0; et < 4; et++) {
            if (!(active & (1 << et))) continue;
            uint32_t cnt = get32(p); p += 4;
            for (uint32_t i = 0; i < cnt; i++) {
                int len = 4 + byte_num_from_k((int)get32(p));
                if ((uint32_t)len != key_len) return 2; /* graph build only links k-mers of the same k */
                acc += edge_hash(key, p, len, et);
                acc -= edge_hash(p, key, len, mirror[et]);
                edges++;
                p += len;
            }
        }
        off += 8 + rec_len;
    }
    *signed_sum = acc; *n_edges = edges;
    return 0;
}
