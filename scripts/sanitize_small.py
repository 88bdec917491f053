"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck); not a test of results
beyond parity with the oracle on tiny inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "tests")
import numpy as np
import genomix_b200 as gx
from test_gpu_parity import random_reads_text, oracle_canonical_c

for k, paired in ((21, False), (55, True), (91, False)):
    rng = np.random.default_rng(k)
    text = random_reads_text(rng, 400, k + 3, k + 80, paired=paired, genome_len=4000)
    head = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
    text += b"".join(b"%d\t%s\n" % (4 * (9000 + i) + 2, head + bytes(rng.choice(list(b"ACGT"), size=5).tolist())) for i in range(600))
    want = oracle_canonical_c(k, text)
    for kw in ({}, {"start_small": True, "min_capacity": 8192, "chunk_bytes": 20000}, {"stream_records": True, "table_regions": 7}):
        with gx.GraphBuilder(k, **kw) as gb:
            gb.push_lines(text)
            gb.finish()
            got = gx.types.canonical_records(gb.records())
            assert got == want, (k, kw)
            gb.graph_statistics(); gb.coverage_histogram(); gb.coverage_cutoff()
            list(gb.iter_frames(65536))
    with gx.GraphBuilder(k, sort_output=True) as gb:             # radix sort of the node list
        gb.push_lines(text)
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == want
    import torch
    dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    with gx.GraphBuilder(k, chunk_bytes=30000) as gb:             # chunk ends found on the device
        gb.push_lines_device(dev.data_ptr(), dev.numel())
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == want
    a = gx.build_graph(k, text)
    with gx.GraphBuilder(k) as gb:
        gb.push_records(a)
        gb.push_lines(text)
        gb.finish()
        assert len(gx.types.canonical_records(gb.records())) == len(want)
    print("ok", k, flush=True)

# one node with 20000 read heads: CTA-wide head sort, multi-window tile written by emit_write_big_kernel
k = 21
rng = np.random.default_rng(7)
head = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
text = b"".join(b"%d\t%s\t%s\n" % (4 * i + 2, head + bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(3, 9))).tolist()),
                                   bytes(rng.choice(list(b"ACGT"), size=30).tolist())) for i in range(20000))
want = oracle_canonical_c(k, text)
with gx.GraphBuilder(k) as gb:
    gb.push_lines(text)
    gb.finish()
    assert gx.types.canonical_records(gb.records()) == want
print("ok big node", flush=True)
