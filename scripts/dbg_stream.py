import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomix_b200 as gx
sys.path.insert(0, "tests")
from test_gpu_parity import random_reads_text
k = 31
rng = np.random.default_rng(1000 + k)
text = random_reads_text(rng, 4000, k + 5, k + 90, paired=False, genome_len=20000)
whole = gx.build_graph(k, text)
for sl in ("100000", "1000000000"):
    os.environ["GENOMIX_GB_SLICE"] = sl
    with gx.GraphBuilder(k, stream_records=True) as gb:
        gb.push_lines(text); gb.finish()
        got = gb.records()
        a = np.frombuffer(whole, dtype=np.uint8); b = np.frombuffer(got, dtype=np.uint8)
        d = np.flatnonzero(a != b)
        print("slice", sl, "len", len(whole), len(got), "ndiff", d.size, "first", d[:20], "last", d[-5:] if d.size else None)
        got2 = gb.records()
        b2 = np.frombuffer(got2, dtype=np.uint8)
        print("  second pull ndiff", np.flatnonzero(a != b2).size)
        # record boundaries
        offs = [0]
        for kk, v in gx.types.iter_records(whole):
            offs.append(offs[-1] + 8 + len(kk) + len(v))
        offs = np.array(offs)
        if d.size:
            for x in d[:5]:
                i = np.searchsorted(offs, x, side="right") - 1
                print("   diff at", x, "record", i, "rec start", offs[i], "len", offs[i+1]-offs[i], "rel", x - offs[i], "i%32", i % 32)
