"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's goldens. GPU only."""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _gx():
    import genomix_b200 as gx
    return gx


def oracle_canonical(k, text):
    from genomix_b200 import types as T
    recs = O.graph_records(k, O.build_graph(k, text))
    return {key: T.Node.read(val, 0)[0].canonical_bytes() for key, val in recs.items()}


def oracle_canonical_c(k, text):
    """same as oracle_canonical through the C twin (for inputs too large for the pure-Python oracle)"""
    from genomix_b200 import types as T
    from oracle import c_oracle as CO
    return T.canonical_records(CO.build_graph_records(k, text, 4)[0])


def gpu_canonical(k, text, **kw):
    gx = _gx()
    stream = gx.build_graph(k, text, **kw)
    return gx.types.canonical_records(stream), stream


def random_reads_text(rng, n_reads, min_len, max_len, paired=False, genome_len=400, lower=False):
    genome = rng.choice(list(b"ACGT"), size=genome_len).astype(np.uint8)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    lines = []
    for i in range(n_reads):
        def one():
            L = int(rng.integers(min_len, max_len + 1))
            s = int(rng.integers(0, genome_len - L + 1))
            r = bytes(genome[s: s + L].tolist())
            if rng.integers(0, 2):
                r = bytes(comp[c] for c in reversed(r))
            if lower and rng.integers(0, 4) == 0:
                r = r.lower()
            return r
        line = b"%d\t%s" % (4 * i + 2, one())
        if paired:
            line += b"\t" + one()
        lines.append(line)
    return b"\n".join(lines) + b"\n"


@pytest.mark.parametrize("name,k,text,expected", golden_cases(), ids=[c[0] for c in golden_cases()])
def test_reference_goldens(name, k, text, expected):
    """The reference's own expected files, under the reference's own comparison rule (TestUtils.java:67-181)."""
    gx = _gx()
    stream = gx.build_graph(k, text)
    O.compare_unordered(expected, gx.types.records_to_text(stream))
    # and byte-level against the oracle after canonical sorting
    assert gx.types.canonical_records(stream) == oracle_canonical(k, text)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 15, 16, 21, 31, 32, 33, 47, 55, 63, 64, 65, 80, 91, 96, 97, 127, 128])
def test_random_reads_match_oracle(k):
    """bit-exact node set / coverage / edge sets / read-head sets for every key width and word boundary"""
    rng = np.random.default_rng(1000 + k)
    text = random_reads_text(rng, 60, k + 1, k + 40, paired=(k % 2 == 1), genome_len=max(300, 3 * k), lower=True)
    got, _ = gpu_canonical(k, text)
    assert got == oracle_canonical(k, text)


@pytest.mark.parametrize("k,regions", [(3, 2), (21, 7), (31, 64), (55, 5), (64, 1024), (91, 13), (128, 3)])
def test_region_counts_match_oracle(k, regions):
    """the region split with forced region counts (1 .. the multisplit's maximum) on a small input, whole and chunked,
    with growth in the middle of the region walk"""
    gx = _gx()
    rng = np.random.default_rng(2000 + k)
    text = random_reads_text(rng, 300, k + 1, k + 60, paired=(k % 2 == 1), genome_len=max(3000, 30 * k))
    want = oracle_canonical(k, text)
    got, _ = gpu_canonical(k, text, table_regions=regions)
    assert got == want
    got, _ = gpu_canonical(k, text, table_regions=regions, chunk_bytes=3000)
    assert got == want
    with gx.GraphBuilder(k, table_regions=regions, chunk_bytes=20000, min_capacity=8192, start_small=True) as gb:
        gb.push_lines(text)
        gb.finish()
        assert gb.stats()["table_grows"] >= (1 if len(want) > 8192 else 0)
        assert gx.types.canonical_records(gb.records()) == want


def test_low_complexity_and_palindromes_even_k():
    """even k: palindromic k-mers tie to FORWARD; tandem repeats give self loops and multi-edges"""
    for k in (2, 4, 6, 32, 64):
        unit = b"ACGT" * (k // 2 + 8)
        text = b"2\t" + unit + b"\n6\t" + b"AT" * (k + 5) + b"\n10\t" + b"A" * (k + 9) + b"\n14\t" + b"GC" * (k + 3) + b"T" * k + b"\n"
        got, _ = gpu_canonical(k, text)
        assert got == oracle_canonical(k, text)


def test_chunked_push_equals_single_push():
    gx = _gx()
    rng = np.random.default_rng(5)
    text = random_reads_text(rng, 400, 40, 90, paired=True)
    want = oracle_canonical(21, text)
    got, _ = gpu_canonical(21, text, chunk_bytes=4096)   # many internal chunks
    assert got == want
    with gx.GraphBuilder(21) as gb:                     # many push calls, ragged split points
        lines = text.split(b"\n")[:-1]
        for i in range(0, len(lines), 37):
            gb.push_lines(b"\n".join(lines[i: i + 37]) + (b"\n" if i % 2 == 0 else b""))
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == want
        st = gb.stats()
        assert st["lines"] == len(lines) and st["reads"] == 2 * len(lines)
        assert st["distinct_kmers"] == len(want) == gb.num_nodes


def test_table_growth_rehash():
    gx = _gx()
    rng = np.random.default_rng(6)
    text = random_reads_text(rng, 3000, 60, 100, genome_len=200000)
    want = oracle_canonical(31, text)
    with gx.GraphBuilder(31, chunk_bytes=20000, min_capacity=8192) as gb:
        gb.push_lines(text)
        gb.finish()
        assert gb.stats()["table_grows"] >= 1
        assert gx.types.canonical_records(gb.records()) == want


@pytest.mark.parametrize("k", [21, 55, 91])
def test_deferral_and_regrow_when_the_table_starts_too_small(k):
    """test hook (reserved[2] bit 8): no sizing heuristics, the table starts at 8192 slots for ~190k distinct keys: the
    region upsert defers work items at the load limit, the host doubles the table and re-launches them, several times"""
    gx = _gx()
    rng = np.random.default_rng(77 + k)
    text = random_reads_text(rng, 4000, k + 20, k + 80, genome_len=400000)
    want = oracle_canonical_c(k, text)
    with gx.GraphBuilder(k, min_capacity=8192, start_small=True) as gb:
        gb.push_lines(text)
        gb.finish()
        st = gb.stats()
        assert st["table_grows"] >= 4
        assert st["distinct_kmers"] <= 0.75 * st["table_capacity"]
        assert gx.types.canonical_records(gb.records()) == want
    # a wrong (far too small) hint is only a hint
    with gx.GraphBuilder(k, expected_kmers=1000, min_capacity=8192) as gb:
        gb.push_lines(text)
        gb.finish()
        assert gb.stats()["table_grows"] >= 1
        assert gx.types.canonical_records(gb.records()) == want


def test_edge_inputs():
    gx = _gx()
    # empty job
    with gx.GraphBuilder(5) as gb:
        gb.finish()
        assert gb.num_nodes == 0 and gb.records() == b""
    # reads with non-ACGT are skipped whole; an invalid mate still shows up as the other's mate sequence;
    # CRLF, no final newline, trailing empty fields, lowercase
    text = b"2\tACGTNACGT\n6\tACNGT\tCCGTAAC\r\n10\tacgtacg\t\t\n14\t\tGGATCCA\n18\tTTGACCA\tNNNN"
    got, stream = gpu_canonical(3, text)
    assert got == oracle_canonical(3, text)
    # duplicate read ids collapse in the TreeSet: first line wins
    text = b"2\tACGTAC\n2\tACGTAG\n2\tACGTAC\n"
    got, _ = gpu_canonical(3, text)
    assert got == oracle_canonical(3, text)
    # many heads on one node (EdgeSizePressureTest shape): same read on many lines, heapsort path
    text = b"".join(b"%d\tGATTACAGATTACA\n" % (4 * (300 - i) + 2) for i in range(300))
    got, _ = gpu_canonical(5, text)
    assert got == oracle_canonical(5, text)


def test_lone_carriage_return_ends_a_line():
    """hadoop LineReader.readLine and BufferedReader.readLine (GenomixDriver.java:723-735) accept \\n, \\r and \\r\\n."""
    gx = _gx()
    rng = np.random.default_rng(5)
    reads = [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(9, 40))).tolist()) for _ in range(200)]
    eols = [b"\n", b"\r", b"\r\n"]
    text = b"".join(b"%d\t%s" % (4 * i + 2, r) + eols[int(rng.integers(0, 3))] for i, r in enumerate(reads))
    for t in (text, text.rstrip(b"\r\n"), b"\r".join(b"%d\t%s" % (4 * i + 2, r) for i, r in enumerate(reads)) + b"\r",
              b"\r".join(b"%d\t%s" % (4 * i + 2, r) for i, r in enumerate(reads))):   # also: \r-only files
        want = oracle_canonical(5, t)
        assert len(want) > 100
        assert gpu_canonical(5, t)[0] == want
        assert oracle_canonical_c(5, t) == want
    # a '\r' on a 32-byte boundary of the line index whose '\n' starts the next 32 bytes is one terminator, not two
    line = b"2\t" + b"ACGT" * 7 + b"A"          # 31 bytes + '\r' = 32
    t = line + b"\r\n" + b"6\tGGATTACAGG\r\n"
    assert len(line) == 31
    with gx.GraphBuilder(5) as gb:
        gb.push_lines(t)
        gb.finish()
        assert gb.stats()["lines"] == 2
        assert gx.types.canonical_records(gb.records()) == oracle_canonical(5, t)
    # fastq with \r-terminated lines
    fq = b"".join(b"@r%d\r%s\r+\r%s\r" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    with gx.GraphBuilder(5) as gb:
        gb.push_fastq(fq)
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == oracle_canonical(5, O.fastq_to_readids(fq))


def test_long_read_windows():
    """FrameSizePressureTest shape: one read far longer than a warp window"""
    rng = np.random.default_rng(9)
    read = bytes(rng.choice(list(b"ACGT"), size=6000).tolist())
    text = b"2\t" + read + b"\n6\t" + read[100:4000] + b"\t" + read[50:2500] + b"\n"
    for k in (21, 55):
        got, _ = gpu_canonical(k, text)
        assert got == oracle_canonical(k, text)


@pytest.mark.parametrize("bad,status", [
    (b"2\tACGTA\n7\n", -4), (b"\n", -4), (b"2\tACGT\tACGT\tACGT\n", -4),
    (b"x1\tACGTA\n", -5), (b"\tACGTA\n", -5), (b"99999999999999999999\tACGTA\n", -5),
    (b"2\tACG\n", -6), (b"2\tACGTA\tAC\n", -6),
    (b"%d\tACGTA\n" % (1 << 29), -7), (b"-4\tACGTA\n", -7),
])
def test_errors_match_reference(bad, status):
    gx = _gx()
    with pytest.raises(O.GraphBuildError):
        O.build_graph(3, bad)
    with pytest.raises(gx.GenomixError) as ei:
        gx.build_graph(3, bad)
    assert ei.value.status == status


def test_errors_only_when_reference_throws():
    # the readId / length guards only fire for mates that pass the regex
    for ok in (b"%d\tACNGT\n" % (1 << 30), b"2\tAN\n"):
        assert O.build_graph(3, ok) == {}
        assert _gx().build_graph(3, ok) == b""


def test_frames_and_partitioner():
    gx = _gx()
    rng = np.random.default_rng(12)
    text = random_reads_text(rng, 200, 30, 60, paired=True)
    k = 21
    nb = (k + 3) // 4
    with gx.GraphBuilder(k) as gb:
        gb.push_lines(text)
        gb.finish()
        stream = gb.records()
        recs = list(gx.types.iter_records(stream))
        # R3: Java partition hash over the Kmer bytes
        parts = gb.partition(7)
        assert [O.java_partition(key[4:], 7) for key, _ in recs] == parts.tolist()
        # R5: frames decode back to the same tuples, FrameTupleAccessor-style
        import struct
        tuples = []
        fs = 4096
        for frame in gb.iter_frames(fs):
            (n,) = struct.unpack_from(">i", frame, fs - 4)
            start = 0
            for t in range(n):
                (end,) = struct.unpack_from(">i", frame, fs - 4 - 4 * (t + 1))
                f0, f1 = struct.unpack_from(">ii", frame, start)
                assert f0 == nb
                tuples.append((frame[start + 8: start + 8 + f0], frame[start + 8 + f0: start + 8 + f1]))
                assert start + 8 + f1 == end
                start = end
        assert tuples == [(key[4:], val) for key, val in recs]
        with pytest.raises(gx.GenomixError):
            list(gb.iter_frames(32))


def test_cfg1_full_size():
    """BASELINE config 1 at full size (10 kb genome, 5000 x 100 bp, k=21) against the C oracle if built, else
    the Python oracle on a prefix."""
    gx = _gx()
    w = gx.synth.CONFIGS["cfg1"]
    text = gx.synth.readid_text(w).tobytes()
    lines = text.split(b"\n")[:300]
    sub = b"\n".join(lines) + b"\n"
    got, _ = gpu_canonical(w.k, sub)
    assert got == oracle_canonical(w.k, sub)
    with gx.GraphBuilder(w.k) as gb:
        gb.push_lines(text)
        gb.finish()
        st = gb.stats()
        assert st["kmer_occurrences"] == 5000 * 80
        nodes = gx.types.canonical_records(gb.records())
        # size-independent properties: coverage sums to the occurrence count, edges are symmetric
        tot = 0
        for key, val in nodes.items():
            n, _ = gx.types.Node.read(val, 0)
            tot += int(n.coverage)
        assert tot == st["kmer_occurrences"]


def test_graph_statistics_match_reference_definitions():
    """gx_graph_statistics / gx_coverage_histogram against the oracle's restatement of GraphStatistics.map
    (GraphStatistics.java:78-131) evaluated on the decoded records, and gx_coverage_cutoff against its restatement of the
    driver's FittingMixture cut-off (GenomixDriver.java:120-137)"""
    gx = _gx()
    rng = np.random.default_rng(31)
    text = random_reads_text(rng, 500, 40, 70, paired=True, genome_len=1500) + b"9000\t" + b"AC" * 30 + b"\n9004\t" + b"A" * 40 + b"\n"
    text += b"".join(b"%d\t%s\n" % (4 * (5000 + i) + 2, b"GATTACAGGATCCATTTGCA" * 3) for i in range(1500))   # coverage > 1000 and > 256
    for k in (4, 21, 33):
        with gx.GraphBuilder(k) as gb:
            gb.push_lines(text)
            gb.finish()
            got = gb.graph_statistics()
            hist = gb.coverage_histogram()
            cut = gb.coverage_cutoff(10)
            nodes = {key: gx.types.Node.read(val, 0)[0] for key, val in gx.types.iter_records(gb.records())}
        ref = O.graph_statistics(k, nodes)
        T, M, B = ref["totals"], ref["maximum"], ref["bins"]
        want = {
            "nodes": T["nodes"], "degree_total": T["degree"], "degree_max": M["degree"],
            "degree_bins": [B["degree"].get(i, 0) for i in range(17)],
            "coverage_total": T["coverage"], "coverage_max": M["coverage"],
            "coverage_bins": [B["coverage"].get(i, 0) for i in range(256)] + [sum(v for c, v in B["coverage"].items() if c >= 256)],
            "unflipped_read_ids": T["unflippedReadIds"], "flipped_read_ids": T["flippedReadIds"],
            "self_edges": [T.get("selfEdge-" + n, 0) for n in ("FF", "FR", "RF", "RR")],
            "path_nodes": T.get("pathNode", 0), "tips_forward": T.get("tips-FORWARD", 0), "tips_reverse": T.get("tips-REVERSE", 0),
            "tips_both": T.get("tips-BOTH", 0), "tips_one": T.get("tips-ONE", 0),
            "kmer_length_total": T["kmerLength"], "kmer_length_max": M["kmerLength"],
            "nodes_with_dir": [sum(B.get("kmerLength-with-" + d, {}).values()) for d in ("FORWARD", "REVERSE")],
            "coverage_with_dir_total": [T.get("coverage-with-" + d, 0) for d in ("FORWARD", "REVERSE")],
            "coverage_with_dir_max": [M.get("coverage-with-" + d, 0) for d in ("FORWARD", "REVERSE")],
            "seed_nodes": sum(B.get("scaffoldSeedScore", {}).values()),
            "seed_score_total": T.get("scaffoldSeedScore", 0), "seed_score_max": M.get("scaffoldSeedScore", 0),
            "seed_nodes_with_dir": [sum(B.get("scaffoldSeedScore-with-" + d, {}).values()) for d in ("FORWARD", "REVERSE")],
            "seed_score_with_dir_total": [T.get("scaffoldSeedScore-with-" + d, 0) for d in ("FORWARD", "REVERSE")],
            "seed_score_with_dir_max": [M.get("scaffoldSeedScore-with-" + d, 0) for d in ("FORWARD", "REVERSE")],
        }
        assert got == want, k
        assert want["self_edges"] != [0, 0, 0, 0] and want["coverage_max"] > 1000
        assert want["seed_nodes"] < want["nodes"]                  # the coverage window excludes the deep nodes
        # coverage-bins, unclipped
        assert [int(x) for x in hist] == [B["coverage"].get(i, 0) for i in range(M["coverage"] + 1)]
        # the cut-off: the driver expands the bins into one entry per node (GraphStatistics.getCoverageStats)
        data = [float(c) for c, n in sorted(B["coverage"].items()) for _ in range(n)]
        w_cut, w_em, w_nm, w_ns = O.fitting_mixture(data, float(M["coverage"]), 10)
        assert cut["cutoff"] == w_cut
        assert abs(cut["exp_mean"] - w_em) <= 1e-9 * abs(w_em) and abs(cut["normal_mean"] - w_nm) <= 1e-9 * abs(w_nm)
        assert abs(cut["normal_std"] - w_ns) <= 1e-9 * abs(w_ns)
    with gx.GraphBuilder(5) as gb:   # empty graph: "No information for coverage!"
        gb.finish()
        with pytest.raises(gx.GenomixError):
            gb.coverage_cutoff()


@pytest.mark.parametrize("k", [31, 55, 91, 128])
def test_hot_keys_under_contention(k):
    """thousands of identical and low-complexity reads: every warp hammers the same few slots (CAS races on insertion,
    the claim/publish protocol of KW >= 3 spinning on LOCK, RED.ADD/RED.OR on one value word)"""
    gx = _gx()
    rng = np.random.default_rng(k)
    base = bytes(rng.choice(list(b"ACGT"), size=k + 30).tolist())
    lines = []
    for i in range(6000):
        r = base if i % 3 else b"A" * (k + 20)
        if i % 7 == 0:
            r = r[: k + 5]
        lines.append(b"%d\t%s" % (4 * i + 2, r))
    text = b"\n".join(lines) + b"\n"
    got, _ = gpu_canonical(k, text)
    assert got == oracle_canonical_c(k, text)


def test_reset_and_small_record_batches():
    """gx_reset reuses the allocations for a new job; gx_next_records with a small buffer returns whole records only"""
    gx = _gx()
    rng = np.random.default_rng(99)
    t1 = random_reads_text(rng, 300, 40, 80, paired=True)
    t2 = random_reads_text(rng, 200, 30, 50)
    with gx.GraphBuilder(21) as gb:
        gb.push_lines(t1)
        gb.finish()
        whole = gb.records()
        parts = list(gb.iter_record_batches(batch_bytes=700))
        assert len(parts) > 10 and b"".join(parts) == whole
        for p in parts:
            assert sum(8 + len(k_) + len(v) for k_, v in gx.types.iter_records(p)) == len(p)   # no record is cut
        with pytest.raises(gx.GenomixError):
            list(gb.iter_record_batches(batch_bytes=16))
        assert gx.types.canonical_records(whole) == oracle_canonical(21, t1)
        with pytest.raises(gx.GenomixError):
            gb.push_lines(t2)                     # finished job: must reset first
        gb.reset()
        gb.push_lines(t2)
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == oracle_canonical(21, t2)
        assert gb.stats()["lines"] == 200


@pytest.mark.parametrize("k,paired", [(21, False), (55, True)])
def test_one_node_with_1e5_read_heads(k, paired):
    """EdgeSizePressureTest shape: 10^5 reads that all start with the same k-mer -> ONE node carries 10^5 ReadHeadInfos
    (both sets, duplicates of (readId, mate) among them). The heads are ordered by a CTA-wide sort and written one
    thread per head, so the finish phase stays in the milliseconds; bit-exact against the C oracle."""
    gx = _gx()
    rng = np.random.default_rng(k)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    head = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
    rc_head = bytes(comp[c] for c in reversed(head))
    lines = []
    n = 100_000
    for i in range(n):
        tail = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(3, 12))).tolist())
        read = head + tail if i % 3 else rc_head + tail            # both startReads and endReads of the node
        rid = 4 * (i if i % 11 else i // 2) + 2                     # some (readId, mate) twice: TreeSet keeps one
        line = b"%d\t%s" % (rid, read)
        if paired:
            line += b"\t" + bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(k + 1, k + 20))).tolist())
        lines.append(line)
    text = b"\n".join(lines) + b"\n"
    with gx.GraphBuilder(k) as gb:
        times = []
        for _ in range(2):   # the first build also allocates every buffer inside the timed phase: judge the warm one
            gb.reset()
            gb.push_lines(text)
            gb.finish()
            times.append(gb.phase_ms()["finish"])
        got = gx.types.canonical_records(gb.records())
        heads = gb.stats()["read_heads"]
    want = oracle_canonical_c(k, text)
    assert got == want
    assert heads > n // 2
    print(f"finish phase with 1e5 heads on one node: {times[0]:.1f} ms cold, {times[1]:.1f} ms warm")
    assert min(times) < 250.0, f"finish phase took {times} ms"   # measured: 8.7 ms


@pytest.mark.parametrize("k,paired", [(31, False), (55, True), (91, False)])
def test_streamed_records_equal_resident_records(k, paired, monkeypatch):
    """gx_config.reserved[2] bit 0: records serialised slice by slice inside gx_next_records -- the same records as the
    device-resident stream, the same bytes whatever the caller's buffer size; frames work on top of it. (Two builds of
    the same input may order their records differently -- slot order depends on who wins an insertion race -- so builds
    are compared canonically, pulls of one build byte by byte.)"""
    gx = _gx()
    rng = np.random.default_rng(1000 + k)
    text = random_reads_text(rng, 4000, k + 5, k + 90, paired=paired, genome_len=20000)
    want = oracle_canonical_c(k, text)
    with gx.GraphBuilder(k) as gb2:
        gb2.push_lines(text)
        gb2.finish()
        resident = gb2.records()
        n_frames = len(list(gb2.iter_frames(32768)))
    assert gx.types.canonical_records(resident) == want
    monkeypatch.setenv("GENOMIX_GB_SLICE", "100000")   # many slices even for this small job
    with gx.GraphBuilder(k, stream_records=True) as gb:
        gb.push_lines(text)
        gb.finish()
        assert gb.record_bytes == len(resident)
        whole = gb.records()
        assert gx.types.canonical_records(whole) == want
        assert gb.records() == whole
        assert b"".join(gb.iter_record_batches(batch_bytes=1 << 16)) == whole
        assert b"".join(gb.iter_record_batches(batch_bytes=5000)) == whole
        with pytest.raises(gx.GenomixError):
            gb.records_device()
        frames = list(gb.iter_frames(32768))
        assert len(frames) == n_frames
        # the frames carry the same tuples in the same order as the record stream
        recs = list(gx.types.iter_records(whole))
        assert sum(int.from_bytes(f[-4:], "big") for f in frames) == len(recs)


@pytest.mark.parametrize("k,paired", [(5, False), (31, True), (55, False), (91, True)])
def test_push_records_merges_like_the_aggregator(k, paired):
    """gx_push_records = the merge half of AggregateKmerAggregateFactory (:128-144) on serialised Nodes: graphs built
    separately from two halves of the input, folded into one job, equal the graph of the whole input -- edge lists united,
    read-head sets united (the same (readId, mate) on both sides stays once), coverages added."""
    gx = _gx()
    rng = np.random.default_rng(4000 + k)
    t1 = random_reads_text(rng, 700, k + 3, k + 60, paired=paired, genome_len=3000)
    # second half: other reads of the same genome-free random source plus some lines of the first half again
    t2 = random_reads_text(rng, 500, k + 3, k + 60, paired=paired, genome_len=3000) + b"\n".join(t1.split(b"\n")[:50]) + b"\n"
    want = oracle_canonical_c(k, t1 + t2)
    a, b = gx.build_graph(k, t1), gx.build_graph(k, t2)
    # records + records into an empty job
    with gx.GraphBuilder(k) as gb:
        gb.push_records(a)
        gb.push_records(b)
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == want
        assert gb.stats()["kmer_occurrences"] == sum(int(gx.types.Node.read(v, 0)[0].coverage) for v in want.values())
    # lines, then records on top (and in the other order), in small pieces
    for first_lines in (True, False):
        with gx.GraphBuilder(k, min_capacity=8192) as gb:
            if first_lines:
                gb.push_lines(t1)
            pieces, pos, offs = [], 0, [0]
            for key, val in gx.types.iter_records(b):
                offs.append(offs[-1] + 8 + len(key) + len(val))
            for cut in offs[::97] + [offs[-1]]:
                if cut > pos:
                    gb.push_records(b[pos:cut])
                    pos = cut
            if not first_lines:
                gb.push_lines(t1)
            gb.finish()
            assert gx.types.canonical_records(gb.records()) == want, first_lines
    # what is not a graph-build record of this k is refused, not folded in wrongly
    with gx.GraphBuilder(k) as gb:
        with pytest.raises(gx.GenomixError) as ei:
            gb.push_records(a[:-3])
        assert ei.value.status == -4
    other = gx.build_graph(k + 1, t1)
    with gx.GraphBuilder(k) as gb:
        with pytest.raises(gx.GenomixError) as ei:
            gb.push_records(other)
        assert ei.value.status == -4


@pytest.mark.parametrize("k", [21, 55])
def test_split_falls_back_to_exact_counts_when_the_sample_misleads(k):
    """The split sizes its buckets from every 16th line. Here exactly those lines are ordinary reads while all the others repeat
    one low-complexity read, so one bucket receives hundreds of thousands of records the sample never saw: the placement flags
    the overflow, the chunk is split again with exact counts, the graph is the oracle's."""
    gx = _gx()
    rng = np.random.default_rng(k)
    hot = b"A" * (k + 40)
    lines = []
    for i in range(20000):
        r = bytes(rng.choice(list(b"ACGT"), size=k + 30).tolist()) if i % 16 == 0 else hot
        lines.append(b"%d\t%s" % (4 * i + 2, r))
    text = b"\n".join(lines) + b"\n"
    with gx.GraphBuilder(k) as gb:
        gb.push_lines(text)
        gb.finish()
        st = gb.stats()
        got = gx.types.canonical_records(gb.records())
    assert st["split_redos"] == 1
    assert got == oracle_canonical_c(k, text)
    # ordinary data of the same size: the estimate holds
    text2 = random_reads_text(rng, 20000, k + 5, k + 40, genome_len=50000)
    with gx.GraphBuilder(k) as gb:
        gb.push_lines(text2)
        gb.finish()
        assert gb.stats()["split_redos"] == 0
        assert gx.types.canonical_records(gb.records()) == oracle_canonical_c(k, text2)


def test_push_lines_device_cuts_chunks_at_line_ends():
    """gx_push_lines_device on text resident in HBM, with an internal chunk size far below the text: the chunk ends are found
    on the device, behind a '\\n' or a lone '\\r', never inside a "\\r\\n" pair."""
    import torch
    gx = _gx()
    rng = np.random.default_rng(77)
    reads = [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(30, 90))).tolist()) for _ in range(3000)]
    for eols in ([b"\n"], [b"\r\n"], [b"\n", b"\r", b"\r\n"]):
        text = b"".join(b"%d\t%s" % (4 * i + 2, r) + eols[i % len(eols)] for i, r in enumerate(reads))
        want = oracle_canonical_c(21, text)
        dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
        for chunk in (4096, 70000):
            with gx.GraphBuilder(21, chunk_bytes=chunk) as gb:
                gb.push_lines_device(dev.data_ptr(), dev.numel())
                gb.finish()
                assert gb.stats()["lines"] == 3000
                assert gx.types.canonical_records(gb.records()) == want


@pytest.mark.parametrize("k,paired", [(3, False), (21, True), (31, False), (55, True), (91, False)])
def test_sort_output_is_kmer_pointable_order(k, paired):
    """gx_config.sort_output: the records leave in the order of KmerPointable.compare (KmerPointable.java:94-107: unsigned
    byte order of the Kmer bytes), the order of the reference's part files; same record set, streamed or resident."""
    gx = _gx()
    rng = np.random.default_rng(500 + k)
    text = random_reads_text(rng, 1500, k + 2, k + 70, paired=paired, genome_len=6000)
    want = oracle_canonical_c(k, text) if k > 3 else oracle_canonical(k, text)
    for kw in ({}, {"stream_records": True}):
        with gx.GraphBuilder(k, sort_output=True, **kw) as gb:
            gb.push_lines(text)
            gb.finish()
            stream = gb.records()
            keys = [key[4:] for key, _ in gx.types.iter_records(stream)]        # Kmer bytes behind the VKmer length header
            assert keys == sorted(keys) and len(set(keys)) == len(keys) == len(want)
            assert gx.types.canonical_records(stream) == want
            if not kw:
                tuples = []
                for frame in gb.iter_frames(16384):
                    n = int.from_bytes(frame[-4:], "big")
                    ends = [int.from_bytes(frame[len(frame) - 4 * (i + 2): len(frame) - 4 * (i + 1)], "big") for i in range(n)]
                    start = 0
                    for e in ends:
                        nbk = int.from_bytes(frame[start:start + 4], "big")
                        tuples.append(bytes(frame[start + 8:start + 8 + nbk]))
                        start = e
                assert tuples == keys                                            # frames carry the same order
