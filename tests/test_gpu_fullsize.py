"""BASELINE.json configs at FULL size on the GPU, checked through size-independent properties and, where the C oracle
finishes in about a minute on the box's cores, bit-exactly through the canonical fingerprint (a checksum of per-record
checksums that is invariant under record order and VKmerList order -- the reference's comparison rule)."""
import os

import numpy as np
import pytest

from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def gpu_build_numpy(gx, k, text_np, **kw):
    import ctypes as C
    with gx.GraphBuilder(k, **kw) as gb:
        gb.push_lines(text_np)
        gb.finish()
        n = gb.record_bytes
        out = np.empty(max(n, 1), dtype=np.uint8)
        cursor, used, pos = C.c_uint64(0), C.c_size_t(0), 0
        while pos < n:
            gb._check(gb._lib.gx_next_records(gb._ctx, C.byref(cursor), C.c_void_p(out.ctypes.data + pos), n - pos, C.byref(used)))
            pos += used.value
        return out[:n], gb.stats()


def check_properties(stream, stats, n_reads_split):
    fp = CO.canonical_fingerprint(stream)
    assert fp.records == stats["distinct_kmers"]
    assert fp.coverage_total == stats["kmer_occurrences"]          # float coverages are exact integers below 2^24
    assert fp.heads == stats["read_heads"] == n_reads_split         # one ReadHeadInfo per split mate (ids are unique)
    sym, n_edges = CO.edge_symmetry(stream)
    assert sym == 0 and n_edges == fp.edges and n_edges > 0
    return fp


def test_cfg1_full_size_bit_exact():
    import genomix_b200 as gx
    w = gx.synth.CONFIGS["cfg1"]
    text = gx.synth.readid_text(w)
    stream, st = gpu_build_numpy(gx, w.k, text)
    assert st["kmer_occurrences"] == gx.synth.occurrences(w) == 400000
    fp = check_properties(stream, st, w.n_reads)
    want, ost = CO.build_graph_records(w.k, text, os.cpu_count() or 1, as_numpy=True)
    assert CO.canonical_fingerprint(want).key() == fp.key()
    assert ost["nodes"] == fp.records


def test_cfg2_full_size_properties_and_fingerprint():
    """configs[1] (the bench workload) at full size: 1.53 M reads, 1.84e8 k-mer occurrences, ~5.2e7 nodes."""
    import psutil
    import genomix_b200 as gx
    w = gx.synth.CONFIGS["cfg2"]
    text = gx.synth.readid_text(w)
    stream, st = gpu_build_numpy(gx, w.k, text)
    assert st["kmer_occurrences"] == gx.synth.occurrences(w)
    fp = check_properties(stream, st, w.n_reads)
    # invariance: other chunking, other region counts, a capacity hint, and a table that has to grow several times
    del stream
    for kw in ({"chunk_bytes": 16 << 20}, {"table_regions": 9, "chunk_bytes": 48 << 20}, {"expected_kmers": 55_000_000},
               {"start_small": True, "chunk_bytes": 64 << 20}):
        s2, st2 = gpu_build_numpy(gx, w.k, text, **kw)
        assert CO.canonical_fingerprint(s2).key() == fp.key(), kw
        del s2
    # bit-exact against the C oracle at full size when the host can hold its tuples (~25 GB), else on the first quarter
    big = psutil.virtual_memory().available > (70 << 30)
    n_sub = w.n_reads if big else w.n_reads // 4
    sub = text if big else gx.synth.readid_text(w, n_reads=n_sub)
    want, ost = CO.build_graph_records(w.k, sub, os.cpu_count() or 1, as_numpy=True)
    if big:
        got_key = fp.key()
    else:
        s3, _ = gpu_build_numpy(gx, w.k, sub)
        got_key = CO.canonical_fingerprint(s3).key()
    assert CO.canonical_fingerprint(want).key() == got_key


def test_cfg4s_k55_high_coverage_fingerprint():
    """The single-GPU twin of configs[3] (the config the north-star target is quoted on): k=55 (two key words, 128-bit CAS),
    100x coverage (every key is hot: ~64 occurrences), 4.28e8 occurrences, 6.7e6 nodes. Properties at full size; bit-exact
    against the C oracle on as many reads as the host can hold the oracle's tuples for."""
    import psutil
    import genomix_b200 as gx
    w = gx.synth.CONFIGS["cfg4s"]
    text = gx.synth.readid_text(w)
    stream, st = gpu_build_numpy(gx, w.k, text)
    assert st["kmer_occurrences"] == gx.synth.occurrences(w)
    fp = check_properties(stream, st, w.n_reads)
    del stream
    s2, _ = gpu_build_numpy(gx, w.k, text, chunk_bytes=96 << 20, table_regions=5)     # other chunking / region count: same graph
    assert CO.canonical_fingerprint(s2).key() == fp.key()
    del s2
    avail = psutil.virtual_memory().available
    n_sub = w.n_reads if avail > (160 << 30) else (w.n_reads // 4 if avail > (48 << 30) else w.n_reads // 16)
    sub = text if n_sub == w.n_reads else gx.synth.readid_text(w, n_reads=n_sub)
    want, ost = CO.build_graph_records(w.k, sub, os.cpu_count() or 1, as_numpy=True)
    if n_sub == w.n_reads:
        got_key = fp.key()
    else:
        s3, _ = gpu_build_numpy(gx, w.k, sub)
        got_key = CO.canonical_fingerprint(s3).key()
    assert CO.canonical_fingerprint(want).key() == got_key


@pytest.mark.parametrize("name", ["cfg3s", "cfg5s"])
def test_scaled_multi_gpu_configs_properties(name):
    """k=55 paired-end and k=91 high-error twins of configs[2] and configs[4] (single-GPU sized): properties at size,
    bit-exact fingerprint on a prefix the oracle finishes quickly."""
    import genomix_b200 as gx
    w = gx.synth.CONFIGS[name]
    n_reads = w.n_reads // 4
    text = gx.synth.readid_text(w, n_reads=n_reads)
    stream, st = gpu_build_numpy(gx, w.k, text)
    assert st["kmer_occurrences"] == gx.synth.occurrences(w, n_reads)
    check_properties(stream, st, n_reads * (2 if w.paired else 1))
    n_sub = 20000
    sub = gx.synth.readid_text(w, n_reads=n_sub)
    got, _ = gpu_build_numpy(gx, w.k, sub)
    want, _ = CO.build_graph_records(w.k, sub, os.cpu_count() or 1, as_numpy=True)
    assert CO.canonical_fingerprint(want).key() == CO.canonical_fingerprint(got).key()
