// gx_parse.cuh -- line index + line parser kernels (k-mer-width independent).
//
// Replaces, for a chunk of text resident in HBM:
//   * hadoop TextInputFormat record splitting feeding HDFSReadOperatorDescriptor.initialize
//     (hyracks-hdfs-core/.../dataflow/HDFSReadOperatorDescriptor.java:128-133), and
//   * the per-line front half of ReadsKeyValueParserFactory.parse
//     (genomix-hyracks/.../graph/dataflow/ReadsKeyValueParserFactory.java:95-148): split on \t, field
//     count check, Long.parseLong, [ACGTacgt]+ per mate, the readId / k>=len guards.
#pragma once
#include "gx_internal.cuh"
#include "gx_scan.cuh"

namespace gx {

struct Counters {
    u64 chunk_lines;     // lines in the current chunk (set by the line-index scan)
    u64 chunk_reads;     // split mates in the current chunk
    u64 chunk_occ;       // k-mer occurrences in the current chunk
    u64 chunk_store;     // read-store bytes reserved by the current chunk
    u64 lines, reads, bases, occurrences;  // cumulative
    u64 head_cursor;     // next free index of the heads array
    u64 store_cursor;    // next free byte of the read store
    u64 distinct;        // occupied table slots
    u64 error;           // min over (global line index << 8 | code); ~0 = none
    u64 heads_missing;   // internal consistency: heads whose key was not found at finish (must stay 0)
    u64 read_heads;      // ReadHeadInfo entries after de-duplication
    u64 table_overflow;  // upserts that ran out of probe budget (must stay 0)
    u64 chunk_max_line_occ;  // largest k-mer occurrence count of one line in the current chunk
    // spill area: (key, mask) records whose upsert ran out of probe budget because the table filled up faster than
    // the host predicted; the host grows the table and inserts them afterwards (gx_api.cu handle_spills)
    u64 spill_count;
    u64 spill_cap;
    u64* spill_keys;
    unsigned short* spill_meta;
    u32* spill_counts;   // occurrence count carried by each spilled record
    u64 deferred_count;  // work items of the region upsert postponed because the table reached its load limit
    u64 split_overflow;  // a bucket's estimated room in the record arena was too small (the host redoes the chunk with exact counts)
    u64 upsert_ticket;   // next work item of the running region upsert (gx_split.cuh upsert_regions_kernel)
    u64 big_groups;      // read-head groups too large for one thread (gx_emit.cuh heads_group_big_kernel)
    u64 big_tiles;       // write-pass tiles too large for one warp (gx_emit.cuh emit_write_big_kernel)
    u64 scratch[2];
};

enum LineError : u32 {  // low byte of Counters::error; ordered like gx_status
    LE_FORMAT = 4, LE_NUMBER = 5, LE_TOO_SHORT = 6, LE_READID = 7
};

static constexpr int LI_THREADS = 256;
static constexpr int LI_BYTES_PER_THREAD = 32;
static constexpr int LI_TILE = LI_THREADS * LI_BYTES_PER_THREAD;  // 8 KB of text per CTA

// line-terminator byte-mask of the 32 bytes [pos, pos+32) of text, bytes >= n excluded: bit i set <=> text[pos+i] ends a
// line, i.e. it is '\n', or a '\r' that is not followed by '\n' (hadoop LineReader.readLine and BufferedReader.readLine
// both accept \n, \r and \r\n; of a \r\n pair the '\n' is the terminator and the '\r' is dropped by the line parser).
__device__ __forceinline__ u32 newline_mask32(const uint8_t* __restrict__ text, u64 n, u64 pos) {
    u32 nl = 0, cr = 0;
    if (pos >= n) return 0;
    const uint8_t* p = text + pos;
    if (pos + 32 <= n && (((uintptr_t)p) & 3) == 0) {
        const u32* w = reinterpret_cast<const u32*>(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u32 x = __ldg(w + i);
            const u32 eq = __vcmpeq4(x, 0x0a0a0a0au);  // 0xff per matching byte
            nl |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * i);
            const u32 ec = __vcmpeq4(x, 0x0d0d0d0du);
            cr |= ((ec & 1u) | ((ec >> 7) & 2u) | ((ec >> 14) & 4u) | ((ec >> 21) & 8u)) << (4 * i);
        }
    } else {
        const int lim = (int)min((u64)32, n - pos);
        for (int i = 0; i < lim; ++i) {
            const uint8_t c = __ldg(p + i);
            nl |= (u32)(c == '\n') << i;
            cr |= (u32)(c == '\r') << i;
        }
    }
    if (cr) {   // rare: which of the '\r' are followed by '\n'?
        u32 next_nl = nl >> 1;
        if ((cr >> 31) && pos + 32 < n && __ldg(p + 32) == '\n') next_nl |= 1u << 31;
        nl |= cr & ~next_nl;
    }
    return nl;
}

static __global__ void __launch_bounds__(LI_THREADS) count_newlines_kernel(const uint8_t* __restrict__ text, u64 n,
                                                                    u64* __restrict__ tile_sums) {
    const u64 pos = (u64)blockIdx.x * LI_TILE + (u64)threadIdx.x * LI_BYTES_PER_THREAD;
    const u64 c = __popc(newline_mask32(text, n, pos));
    const u64 tot = block_reduce_sum<LI_THREADS>(c);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// nl_pos[i] = chunk-relative offset of the i-th line terminator. A final unterminated line gets a
// virtual terminator at offset n (written by the host-side launch of finish_line_index_kernel).
static __global__ void __launch_bounds__(LI_THREADS) write_newlines_kernel(const uint8_t* __restrict__ text, u64 n,
                                                                    const u64* __restrict__ tile_base,
                                                                    u32* __restrict__ nl_pos) {
    const u64 pos = (u64)blockIdx.x * LI_TILE + (u64)threadIdx.x * LI_BYTES_PER_THREAD;
    u32 m = newline_mask32(text, n, pos);
    u64 tot;
    u64 idx = tile_base[blockIdx.x] + block_scan_excl<LI_THREADS>((u64)__popc(m), &tot);
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        nl_pos[idx++] = (u32)(pos + b);
    }
}

// after the tile-sum scan: account for a final unterminated line, publish the line count
static __global__ void finish_line_index_kernel(const uint8_t* __restrict__ text, u64 n, const u64* __restrict__ total_nl,
                                         u32* __restrict__ nl_pos, Counters* __restrict__ ctr) {
    u64 lines = *total_nl;
    if (n > 0 && text[n - 1] != '\n' && text[n - 1] != '\r') { nl_pos[lines] = (u32)n; lines += 1; }
    ctr->chunk_lines = lines;
    ctr->chunk_reads = 0;
    ctr->chunk_occ = 0;
    ctr->chunk_store = 0;
    ctr->chunk_max_line_occ = 0;
}

static __global__ void set_chunk_lines_kernel(Counters* ctr, u64 n) {
    ctr->chunk_lines = n;
    ctr->chunk_reads = 0;
    ctr->chunk_occ = 0;
    ctr->chunk_store = 0;
    ctr->chunk_max_line_occ = 0;
}

__device__ __forceinline__ void report_line_error(Counters* ctr, u64 global_line, u32 code) {
    atomicMin(&ctr->error, (global_line << 8) | (u64)code);
}

static constexpr int PL_THREADS = 256;

// Per-mate checks in the order the reference hits them (ReadsKeyValueParserFactory.java:125-145): regex, then
// makeUUID's readId guard (ReadHeadInfo.java:110-113), then k >= length (:152-155). Fills the split flags and
// this line's contribution to the counters.
__device__ __forceinline__ void validate_mates(LineDesc& d, bool has_mate_field, bool neg, u64 mag, bool ok0, bool ok1, int k,
                                               u64 gline, Counters* ctr, u32& n_split, u32& store_bytes, u64& bases, u64& occ) {
    const bool v0 = ok0 && d.len[0] > 0;
    const bool v1 = has_mate_field && ok1 && d.len[1] > 0;
    const bool id_bad = neg || mag >= (1ull << 29);
    bool dead = false;
    if (v0) {
        if (id_bad) { report_line_error(ctr, gline, LE_READID); dead = true; }
        else if ((u32)k >= d.len[0]) { report_line_error(ctr, gline, LE_TOO_SHORT); dead = true; }
    }
    if (!dead && v1) {
        if (id_bad) { report_line_error(ctr, gline, LE_READID); dead = true; }
        else if ((u32)k >= d.len[1]) { report_line_error(ctr, gline, LE_TOO_SHORT); dead = true; }
    }
    if (!dead) {
        if (v0) { d.flags |= 1u; ++n_split; bases += d.len[0]; occ += d.len[0] - k + 1; }
        if (v1) { d.flags |= 2u; ++n_split; bases += d.len[1]; occ += d.len[1] - k + 1; }
        if (n_split) store_bytes = (d.len[0] + 3) / 4 + (d.len[1] + 3) / 4;
    }
}

// Reserve head indices, read-store bytes and flat occurrence indices for the CTA's lines (one atomic each per CTA),
// update the job counters, and store the descriptor.
template <int NT>
__device__ __forceinline__ void reserve_and_store(LineDesc& d, bool in_range, u64 i, u32 n_split, u32 store_bytes, u64 bases,
                                                  u64 occ, LineDesc* __restrict__ desc, Counters* __restrict__ ctr) {
    u64 tot;
    const u64 packed = ((u64)store_bytes << 20) | (u64)n_split;
    const u64 ex = block_scan_excl<NT>(packed, &tot);
    __shared__ u64 base_heads, base_store, base_occ;
    u64 occ_tot;
    const u64 occ_ex = block_scan_excl<NT>(occ, &occ_tot);
    if (threadIdx.x == 0) {
        base_heads = atomicAdd(&ctr->head_cursor, tot & 0xfffffull);
        base_store = atomicAdd(&ctr->store_cursor, tot >> 20);
    }
    const u64 b_sum = block_reduce_sum<NT>(bases);
    const u64 o_sum = block_reduce_sum<NT>(occ);
    {   // per-line maximum: warp reduce, one atomic per warp that has something to report
        u64 mx = occ;
#pragma unroll
        for (int dl = 16; dl > 0; dl >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, dl));
        if ((threadIdx.x & 31) == 0 && mx > 0 && mx > ctr->chunk_max_line_occ) atomicMax(&ctr->chunk_max_line_occ, mx);
    }
    if (threadIdx.x == 0) {
        atomicAdd(&ctr->chunk_reads, tot & 0xfffffull);
        atomicAdd(&ctr->chunk_store, tot >> 20);
        base_occ = atomicAdd(&ctr->chunk_occ, o_sum);
        atomicAdd(&ctr->reads, tot & 0xfffffull);
        atomicAdd(&ctr->bases, b_sum);
        atomicAdd(&ctr->occurrences, o_sum);
    }
    __syncthreads();
    if (in_range) {
        u64 h = base_heads + (ex & 0xfffffull);
        const u64 s = base_store + (ex >> 20);
        if (d.flags & 1u) d.head_idx[0] = h++;
        if (d.flags & 2u) d.head_idx[1] = h++;
        d.store[0] = s;
        d.store[1] = s + (d.len[0] + 3) / 4;
        d.occ_base = base_occ + occ_ex;
        desc[i] = d;
    }
}

// One thread per line.
static __global__ void __launch_bounds__(PL_THREADS) parse_lines_kernel(const uint8_t* __restrict__ text,
                                                                 const u32* __restrict__ nl_pos,
                                                                 u64 first_global_line, int k,
                                                                 LineDesc* __restrict__ desc,
                                                                 Counters* __restrict__ ctr) {
    const u64 i = (u64)blockIdx.x * PL_THREADS + threadIdx.x;
    const u64 n_lines = ctr->chunk_lines;  // written by finish_line_index_kernel earlier on the stream
    u32 n_split = 0, store_bytes = 0;
    u64 bases = 0, occ = 0;
    LineDesc d;
    d.off[0] = d.off[1] = 0; d.len[0] = d.len[1] = 0; d.read_id = 0; d.flags = 0; d.pad = 0;
    d.store[0] = d.store[1] = 0; d.head_idx[0] = d.head_idx[1] = 0; d.occ_base = 0;
    if (i < n_lines) {
        const u32 start = (i == 0) ? 0u : nl_pos[i - 1] + 1u;
        u32 end = nl_pos[i];
        if (end > start && __ldg(text + end - 1) == '\r') --end;  // hadoop LineReader drops the CR of CRLF
        // ---- String.split("\t"): field f spans [fs[f], fe[f]); trailing empty fields are dropped
        u32 fs[3], fe[3];
        fs[0] = start; fs[1] = fs[2] = end; fe[0] = fe[1] = fe[2] = end;
        int field = 0, last_nonempty = -1;
        u32 fstart = start;
        bool ok0 = true, ok1 = true;  // [ACGTacgt]* so far for fields 1 and 2
        for (u32 p = start; p <= end; ++p) {
            const u32 c = (p < end) ? (u32)__ldg(text + p) : (u32)'\t';  // virtual tab closes the last field
            if (c == '\t') {
                if (p > fstart) last_nonempty = field;
                if (field < 3) { fs[field] = fstart; fe[field] = p; }
                ++field;
                fstart = p + 1;
            } else if (field == 1 || field == 2) {
                bool ok;
                (void)code_of(c, ok);
                if (field == 1) ok0 = ok0 && ok; else ok1 = ok1 && ok;
            }
        }
        const int nf = last_nonempty + 1;
        const u64 gline = first_global_line + i;
        bool dead = false;
        if (nf != 2 && nf != 3) { report_line_error(ctr, gline, LE_FORMAT); dead = true; }
        // ---- Long.parseLong(field 0)
        u64 mag = 0; bool neg = false;
        if (!dead) {
            u32 p = fs[0];
            const u32 e = fe[0];
            if (p < e) { const u32 c = __ldg(text + p); if (c == '-' || c == '+') { neg = (c == '-'); ++p; } }
            bool okn = p < e;
            for (; p < e && okn; ++p) {
                const u32 c = __ldg(text + p);
                if (c < '0' || c > '9') { okn = false; break; }
                const u64 dgt = c - '0';
                if (mag > (0x8000000000000000ull - dgt) / 10ull) { okn = false; break; }  // |value| <= 2^63
                mag = mag * 10ull + dgt;
            }
            if (okn && !neg && mag > 0x7fffffffffffffffull) okn = false;
            if (!okn) { report_line_error(ctr, gline, LE_NUMBER); dead = true; }
        }
        if (!dead) {
            d.read_id = neg ? (u64)(-(long long)mag) : mag;
            d.off[0] = fs[1]; d.len[0] = fe[1] - fs[1];
            if (nf == 3) { d.off[1] = fs[2]; d.len[1] = fe[2] - fs[2]; d.flags |= 4u; }
            validate_mates(d, nf == 3, neg, mag, ok0, ok1, k, gline, ctr, n_split, store_bytes, bases, occ);
        }
    }
    reserve_and_store<PL_THREADS>(d, i < n_lines, i, n_split, store_bytes, bases, occ, desc, ctr);
}

// ---------------------------------------------------------------------------------------------------------
// R6 fused with R1: fastq records -> the same LineDesc the text parser produces.
// GenomixDriver.convertAndUploadFastqToHDFS (GenomixDriver.java:665-714): record i of the file is 0-based line 4i+1,
// its id is the post-incremented line number 4i+2, the sequence (and the mate's, read in lock step from the second
// file) is String.trim()med. The text line it would have written is then parsed by ReadsKeyValueParserFactory.parse,
// so an empty (trimmed) sequence leaves the line with too few fields -> the same IllegalStateException.
static __global__ void __launch_bounds__(PL_THREADS) parse_fastq_kernel(const uint8_t* __restrict__ text,
                                                                 const u32* __restrict__ nl1, u64 n_lines1,
                                                                 u32 base2, const u32* __restrict__ nl2, int paired,
                                                                 u64 first_record, u64 first_global_line, int k,
                                                                 LineDesc* __restrict__ desc, Counters* __restrict__ ctr) {
    const u64 i = (u64)blockIdx.x * PL_THREADS + threadIdx.x;
    const u64 n_records = (n_lines1 + 2) / 4;  // records whose sequence line 4i+1 exists
    u32 n_split = 0, store_bytes = 0;
    u64 bases = 0, occ = 0;
    LineDesc d;
    d.off[0] = d.off[1] = 0; d.len[0] = d.len[1] = 0; d.read_id = 0; d.flags = 0; d.pad = 0;
    d.store[0] = d.store[1] = 0; d.head_idx[0] = d.head_idx[1] = 0; d.occ_base = 0;
    if (i < n_records) {
        const u64 j = 4 * i + 1;
        bool ok[2] = {true, true};
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (m == 1 && !paired) break;
            const u32* nl = m ? nl2 : nl1;
            const u32 b = m ? base2 : 0u;
            u32 s = b + nl[j - 1] + 1u, e = b + nl[j];
            while (s < e && __ldg(text + s) <= ' ') ++s;          // String.trim(): strips chars <= U+0020 at both ends
            while (e > s && __ldg(text + e - 1) <= ' ') --e;
            d.off[m] = s; d.len[m] = e - s;
            for (u32 p = s; p < e; ++p) { bool c_ok; (void)code_of(__ldg(text + p), c_ok); ok[m] = ok[m] && c_ok; }
        }
        const u64 id = 4 * (first_record + i) + 2;
        d.read_id = id;
        const u64 gline = first_global_line + i;
        // fields of the text line "<id>\t<seq>[\t<mate>]" after String.split drops trailing empty strings
        const int nf = paired ? (d.len[1] ? 3 : (d.len[0] ? 2 : 1)) : (d.len[0] ? 2 : 1);
        if (nf < 2) report_line_error(ctr, gline, LE_FORMAT);
        else {
            if (nf == 3) d.flags |= 4u; else { d.len[1] = 0; }
            validate_mates(d, nf == 3, false, id, ok[0], ok[1], k, gline, ctr, n_split, store_bytes, bases, occ);
        }
    }
    reserve_and_store<PL_THREADS>(d, i < n_records, i, n_split, store_bytes, bases, occ, desc, ctr);
}

}  // namespace gx
