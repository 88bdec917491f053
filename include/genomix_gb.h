/*
 * genomix_gb.h -- C ABI of the B200-native Genomix graph-build path (libgenomix_gb.so).
 *
 * This is the drop-in boundary: what a JNI / ctypes / cgo stub binds in place of the reference's
 * pure-Java operators for the graph-build job (reference: genomix/genomix-hyracks,
 * JobGenBuildBrujinGraph.java:79-90). Plain pointers and sizes only; no C++ or torch types; no
 * exceptions or callbacks cross the boundary. All functions return 0 on success or a negative
 * gx_status; gx_last_error() gives the message the reference would have thrown.
 *
 * Reference interfaces replaced (paths relative to /root/reference/):
 *   R1  IKeyValueParser.open/parse/close           hyracks/hyracks-hdfs/hyracks-hdfs-core/src/main/java/edu/uci/ics/hyracks/hdfs/api/IKeyValueParser.java:29-58
 *       impl ReadsKeyValueParserFactory.parse       genomix/genomix-hyracks/src/main/java/edu/uci/ics/genomix/hyracks/graph/dataflow/ReadsKeyValueParserFactory.java:95-254
 *   R2  IAggregatorDescriptor.init/aggregate/outputFinalResult
 *                                                   hyracks/hyracks-dataflow-std/src/main/java/edu/uci/ics/hyracks/dataflow/std/group/IAggregatorDescriptor.java:21-106
 *       impl AggregateKmerAggregateFactory          genomix/genomix-hyracks/.../graph/dataflow/AggregateKmerAggregateFactory.java:93-214
 *   R3  ITuplePartitionComputer.partition           genomix/genomix-hyracks/.../data/primitive/KmerPartitionComputerFactory.java:28-52
 *   R4  ITupleWriter.open/write/close               hyracks/hyracks-hdfs/hyracks-hdfs-core/.../api/ITupleWriter.java:26-57
 *       impl KmerNodePairSequenceWriterFactory      genomix/genomix-hyracks/.../graph/dataflow/KmerNodePairSequenceWriterFactory.java:66-94
 *   R5  frame layout FrameTupleAppender.append      hyracks/hyracks-dataflow-common/src/main/java/edu/uci/ics/hyracks/dataflow/common/comm/io/FrameTupleAppender.java:57-70
 *   R6  GenomixDriver.convertAndUploadFastqToHDFS   genomix/genomix-driver/src/main/java/edu/uci/ics/genomix/driver/GenomixDriver.java:665-714
 *
 * Threading: one gx_ctx is driven by one thread at a time (the reference runs one parser /
 * aggregator / writer instance per task thread, HDFSReadOperatorDescriptor.java:98-143).
 * Ownership: the caller owns every buffer it passes in; buffers returned by *_device calls are
 * owned by the ctx and stay valid until the next gx_reset()/gx_destroy().
 *
 * There is no CPU fallback: every entry point that computes runs CUDA kernels on an sm_100a
 * device and fails with GX_ERR_CUDA when none is usable.
 */
#ifndef GENOMIX_GB_H
#define GENOMIX_GB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_ABI_VERSION 2

typedef struct gx_ctx gx_ctx;

typedef enum gx_status {
    GX_OK = 0,
    GX_ERR_INVALID = -1,      /* bad argument / call sequence (IllegalArgumentException at the call site) */
    GX_ERR_CUDA = -2,         /* CUDA runtime / driver / NCCL failure, or no usable device */
    GX_ERR_NOMEM = -3,        /* device or host allocation failed */
    GX_ERR_FORMAT = -4,       /* IllegalStateException: line is not id\tseq[\tmate]   (ReadsKeyValueParserFactory.java:110-123) */
    GX_ERR_NUMBER = -5,       /* NumberFormatException: read id is not a Java long     (:112,115) */
    GX_ERR_READ_TOO_SHORT = -6, /* IllegalArgumentException: k >= read length           (:152-155) */
    GX_ERR_READID_RANGE = -7, /* IllegalArgumentException: readId loses bits (>= 2^29 or negative) (ReadHeadInfo.java:110-113) */
    GX_ERR_BUFFER = -8,       /* caller buffer too small / tuple larger than a frame   (:245-249) */
    GX_ERR_STATE = -9         /* call not valid in the ctx's current phase */
} gx_status;

typedef struct gx_config {
    int32_t abi_version;        /* GX_ABI_VERSION */
    int32_t kmer_length;        /* genomix.conf.kmerLength (GenomixJobConf.java:303); 1 <= k <= 128 */
    int32_t device;             /* CUDA device ordinal */
    int32_t rank;               /* this process's partition index, 0 <= rank < n_ranks */
    int32_t n_ranks;            /* number of GPUs sharing the key space (1 = single GPU) */
    int32_t sort_output;        /* 0: records leave in table-slot order (default: nothing downstream depends on the order, the
                                 * Pregelix loader hashes vertices by key). 1: in KmerPointable order (KmerPointable.java:94-107:
                                 * unsigned byte order of the Kmer bytes) like the reference's part files, which come out of a
                                 * pre-clustered group-by over ExternalSortOperatorDescriptor; costs a radix sort of the nodes. */
    uint64_t expected_kmers;    /* hint: distinct canonical k-mers this rank will own (0 = grow on demand) */
    uint64_t reserved[4];       /* tuning/test knobs, 0 = default: [0] internal chunk bytes (256 MiB), [1] smallest table
                                 * capacity in slots (2^20), [2] bit 0: stream the records -- gx_finish sizes them, gx_next_records
                                 * serialises them slice by slice while the previous slice travels to the host, so the stream never
                                 * exists in device memory as a whole (gx_records_device / gx_partition_records unavailable);
                                 * bit 8: test hook (no sizing heuristics: the table starts at [1] and grows by deferral),
                                 * [3] table regions per rank of the region-sorted build (0 = regions of about 32 MB) */
} gx_config;

/* Counters of the job so far (the reference only logs wall time: GenomixDriver.java:459,467). */
typedef struct gx_stats {
    uint64_t lines;             /* input records seen by parse() */
    uint64_t reads;             /* mates that passed [ACGTacgt]+ and were split */
    uint64_t bases;             /* letters in those mates */
    uint64_t kmer_occurrences;  /* (read, position) tuples = what the reference would have put in frames */
    uint64_t distinct_kmers;    /* nodes owned by this rank */
    uint64_t read_heads;        /* ReadHeadInfo entries emitted (after TreeSet de-duplication) */
    uint64_t record_bytes;      /* bytes of the serialised record stream after gx_finish() */
    uint64_t table_capacity;    /* hash-table slots */
    uint64_t table_grows;       /* number of rehashes */
    uint64_t exchanged_records; /* k-mer records (and slack) this rank sent to other ranks */
    uint64_t split_redos;       /* chunks whose sampled bucket sizes were too small and that were split again with exact counts */
    uint64_t reserved[5];
} gx_stats;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int gx_create(const gx_config* cfg, gx_ctx** out);
void gx_destroy(gx_ctx* ctx);
int gx_reset(gx_ctx* ctx);                       /* drop all state, keep allocations: ready for a new job */
const char* gx_last_error(const gx_ctx* ctx);    /* never NULL; ctx may be NULL for gx_create failures */
int gx_get_stats(gx_ctx* ctx, gx_stats* out);
int gx_abi_version(void);

/* ---- R1: parser input ------------------------------------------------------------------------
 * Exactly the bytes parse() sees: lines `id\tseq[\tmate]`, '\n'-terminated (a final unterminated
 * line is a record; "\r\n" accepted). May be called any number of times before gx_finish().
 * Line errors are those of the reference (see gx_status); the job is then dead (sticky error). */
int gx_push_lines(gx_ctx* ctx, const uint8_t* host_text, size_t n_bytes);
/* Same, text already resident in device memory of cfg.device (no copy). */
int gx_push_lines_device(gx_ctx* ctx, const uint8_t* dev_text, size_t n_bytes);

/* ---- R6 fused with R1: fastq front end ---------------------------------------------------------
 * r1/r2 are whole-record-aligned chunks of (uncompressed) fastq; r2 == NULL for single-end.
 * first_record is the 0-based index of the chunk's first record in the file, so read ids are the
 * reference's 4*i+2 (GenomixDriver.java:689-690,702-703). Sequence lines are trimmed like
 * String.trim(). Paired chunks must hold the same number of records (else GX_ERR_FORMAT, the
 * reference's IOException at :684-687). */
int gx_push_fastq(gx_ctx* ctx, const uint8_t* host_r1, size_t n1, const uint8_t* host_r2, size_t n2,
                  uint64_t first_record);

/* ---- R2, merge half: serialised Nodes back into the job -------------------------------------------
 * AggregateKmerAggregateFactory.aggregate (genomix-hyracks/.../graph/dataflow/AggregateKmerAggregateFactory.java:128-144)
 * merges the accumulated Node of a key with another serialised Node of the same key: per edge type
 * VKmerList.unionUpdate, ReadHeadSet.unionUpdate (TreeSet.addAll), coverage added. gx_push_records does that for a
 * whole stream of records (the framing of gx_next_records: recordLength | keyLength | VKmer | Node) against
 * everything the job holds so far -- earlier pushes of lines, fastq or records. Graph-build Nodes of this kmer length
 * only (every edge a one-letter shift of the key, no internal kmer, read-head sets stored by value); anything else is
 * GX_ERR_FORMAT. Single-rank jobs. */
int gx_push_records(gx_ctx* ctx, const uint8_t* host_records, size_t n_bytes);

/* ---- R2 + shuffle + emit ------------------------------------------------------------------------
 * Aggregate everything pushed so far and serialise this rank's nodes into the record stream.
 * With n_ranks > 1 the multi-GPU exchange must already have been driven (gx_mg_* below). */
int gx_finish(gx_ctx* ctx);

/* ---- R4: output -------------------------------------------------------------------------------
 * Record stream = concatenation, one per node, of
 *     int32be recordLength | int32be keyLength | VKmer key (int32be k, ceil(k/4) bytes) | Node bytes
 * i.e. the body of an uncompressed SequenceFile v6 <VKmer,Node> between syncs
 * (KmerNodePairSequenceWriterFactory.java:66-94; Node.write Node.java:408-427). */
int64_t gx_num_nodes(gx_ctx* ctx);
int64_t gx_record_bytes(gx_ctx* ctx);
/* Copy up to cap bytes of whole records starting at *cursor (0 at first) into host buf; advances
 * *cursor, sets *used. *used == 0 with GX_OK means end of stream. */
int gx_next_records(gx_ctx* ctx, uint64_t* cursor, uint8_t* host_buf, size_t cap, size_t* used);
/* Device-resident view of the whole stream and of each record's start offset (n_nodes + 1 entries). */
int gx_records_device(gx_ctx* ctx, const uint8_t** dev_records, const uint64_t** dev_offsets);
/* R5: fill one Hyracks frame (frame_size bytes) with as many (Kmer,Node) tuples as fit, in stream
 * order from *cursor (record index). *n_tuples == 0 means end. A tuple that cannot fit an empty
 * frame is GX_ERR_BUFFER (ReadsKeyValueParserFactory.java:245-249). */
int gx_next_frame(gx_ctx* ctx, uint64_t* cursor, uint8_t* host_frame, int32_t frame_size, int32_t* n_tuples);

/* R4 complete: write the record stream as an uncompressed SequenceFile v6 <VKmer,Node> part file, byte-compatible with
 * what SequenceFile.createWriter(conf, out, VKmer.class, Node.class, CompressionType.NONE, null) + append() produce
 * (KmerNodePairSequenceWriterFactory.java:66-94; hadoop-core 0.20.2 SequenceFile.Writer): header "SEQ\6", the two
 * class names, no compression, empty metadata, the 16-byte sync marker; then records, with a sync escape
 * (int -1 + marker) before a record whenever >= 2000 bytes were written since the last sync. `sync16` is the
 * marker (hadoop draws it at random; pass NULL for a marker derived from the file name). If n_parts > 0 only the
 * records whose Java partition hash % n_parts == part are written (part-<part> of an n_parts job). */
int gx_write_sequence_file(gx_ctx* ctx, const char* path, const uint8_t* sync16, int32_t n_parts, int32_t part,
                           uint64_t* bytes_written);

/* ---- R3: partitioner ---------------------------------------------------------------------------
 * Batched KmerPartitionComputerFactory.partition over the record stream (Java 31-polynomial hash,
 * abs, % n_parts): writes one int32 per node into host_parts (n_nodes entries). */
int gx_partition_records(gx_ctx* ctx, int32_t n_parts, int32_t* host_parts);

/* ---- graph statistics --------------------------------------------------------
 * The per-node counters of the reference's post-build statistics job (genomix/genomix-hadoop/src/main/java/edu/uci/ics/genomix/
 * hadoop/utils/GraphStatistics.java:78-131) as one fused reduction over this rank's nodes, without leaving the GPU:
 * inDegree = |RF|+|RR|, outDegree = |FF|+|FR| (Node.java:820-848). */
typedef struct gx_graph_stats {
    uint64_t nodes;
    uint64_t degree_total, degree_max;
    uint64_t degree_bins[17];        /* inDegree + outDegree, 0..16 */
    uint64_t coverage_total, coverage_max;
    uint64_t coverage_bins[257];     /* Math.round(coverage) 0..255, last bin = 256 and above */
    uint64_t unflipped_read_ids, flipped_read_ids;
    uint64_t self_edges[4];          /* totals/selfEdge-FF,FR,RF,RR */
    uint64_t path_nodes;             /* inDegree == 1 && outDegree == 1 */
    uint64_t tips_forward, tips_reverse, tips_both, tips_one;
    /* GraphStatistics.java:87-119. Index 0 = DIR.FORWARD, 1 = DIR.REVERSE: "<x>-with-<DIR>" counts nodes with
     * degree(DIR) != 0. Every node of a graph build holds one k-letter k-mer, so kmerLength[-with-DIR] follows from the
     * node counts. scaffoldSeedScore = Node.calculateSeedScore (Node.java:859-862) of the nodes whose coverage lies in
     * COVERAGE_DIST_MEAN +- COVERAGE_DIST_STD = 0 +- 1000 (GraphStatistics.java:75-76,93-96). */
    uint64_t kmer_length_total, kmer_length_max;
    uint64_t nodes_with_dir[2], coverage_with_dir_total[2], coverage_with_dir_max[2];
    uint64_t seed_nodes, seed_score_total, seed_score_max;
    uint64_t seed_nodes_with_dir[2], seed_score_with_dir_total[2], seed_score_with_dir_max[2];
} gx_graph_stats;
int gx_graph_statistics(gx_ctx* ctx, gx_graph_stats* out);   /* after gx_finish */

/* The unclipped "coverage-bins" group of GraphStatistics (one counter per Math.round(coverage) value): bins[c] = nodes of
 * coverage c for c < n_bins (call gx_graph_statistics first for coverage_max). After gx_finish. */
int gx_coverage_histogram(gx_ctx* ctx, uint64_t* host_bins, uint64_t n_bins);

/* The driver's coverage cut-off (GenomixDriver.setCutoffCoverageByFittingMixture, GenomixDriver.java:120-137, running
 * FittingMixture.fittingMixture, genomix-driver/.../mixture/model/FittingMixture.java:92-216, on this rank's coverage bins):
 * an exponential + normal mixture fitted with `iterations` EM rounds (the driver uses 10); *cutoff = the first coverage at
 * which the normal component outweighs the exponential one, 0 if there is none (the driver then leaves
 * REMOVE_BAD_COVERAGE_MIN_COVERAGE unset). GX_ERR_STATE ("No information for coverage!") on an empty graph.
 * exp_mean / normal_mean / normal_std (may be NULL) receive the fitted parameters (GenomixDriver.cur_*). */
int gx_coverage_cutoff(gx_ctx* ctx, int32_t iterations, int64_t* cutoff, double* exp_mean, double* normal_mean, double* normal_std);

/* ---- multi-GPU (n_ranks > 1): hash-partitioned exchange ----------------------------------------
 * One process per GPU. Bootstrap: rank 0 calls gx_mg_unique_id, the host side broadcasts the 128
 * bytes (torch.distributed / MPI / Hyracks RPC), every rank calls gx_mg_init. After the last
 * gx_push_*, every rank calls gx_mg_exchange (collective): staged k-mer and read-head records are
 * routed to owner = floor(mix(key) * n_ranks / 2^64) with an all-to-all-v over NVLink (copy-engine pushes into CUDA-IPC
 * inboxes; GENOMIX_GB_NO_IPC=1: grouped ncclSend/ncclRecv) and upserted. The call is pipelined: the transfer of the last
 * staged chunk is started but its records are upserted by the next gx_mg_exchange or by gx_finish, so with one exchange per
 * pushed chunk the transfer runs under the next chunk's split. */
int gx_mg_unique_id(uint8_t out_id[128]);
int gx_mg_init(gx_ctx* ctx, const uint8_t id[128]);
int gx_mg_exchange(gx_ctx* ctx);

/* ---- timing hooks (device-side, CUDA events on the ctx's stream) ------------------------------ */
/* Milliseconds spent by the ctx's kernels since gx_reset/gx_create, by phase:
 * [0] line index + parse, [1] region upserts (single GPU), [2] exchange (whole call, upserts included), [3] finish (heads + emit),
 * [4] h2d copies, [5] exchange: NVLink pushes + arrival barrier on the communication stream (overlaps the next chunk's split),
 * [6] exchange: upserts of own + received records, [7] split (bucket sizes + placement). */
int gx_phase_ms(gx_ctx* ctx, float out_ms[8]);
/* Number of kernel launches issued by this ctx since creation. */
uint64_t gx_kernel_launches(gx_ctx* ctx);
/* Make the ctx enqueue on an external CUDA stream (cudaStream_t as void*), e.g. torch's current stream. */
int gx_set_stream(gx_ctx* ctx, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GENOMIX_GB_H */
