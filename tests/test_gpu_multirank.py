"""Multi-GPU parity: 2 (and 4 if present) ranks, hash-partitioned NCCL exchange, union of the ranks' records must
equal the oracle's graph of the whole input. Needs >= 2 GPUs (skipped otherwise)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_union_of_ranks_equals_oracle(world, tmp_path):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import genomix_b200 as gx
    cases = []
    for name, factor, nreads in (("cfg3", 20000, 600), ("cfg2", 2000, 1500), ("cfg5", 200000, 300)):
        w = gx.synth.scaled(gx.synth.CONFIGS[name], factor)
        cases.append((w.k, gx.synth.readid_text(w, n_reads=nreads).tobytes()))
    inp, outp = tmp_path / "cases.pkl", tmp_path / "out.pkl"
    pickle.dump(cases, open(inp, "wb"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mg_worker.py"), str(inp), str(outp)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    results = pickle.load(open(outp, "rb"))
    for (k, text), got in zip(cases, results):
        want = gx.types.canonical_records(CO.build_graph_records(k, text, 4)[0])
        if got != want:
            missing = [key for key in want if key not in got]
            extra = [key for key in got if key not in want]
            differ = [key for key in want if key in got and got[key] != want[key]]
            msg = [f"k={k}: want {len(want)} nodes, got {len(got)}; missing {len(missing)}, extra {len(extra)}, differing {len(differ)}"]
            for key in differ[:3]:
                msg.append("want " + str(gx.types.Node.read(want[key], 0)[0]))
                msg.append("got  " + str(gx.types.Node.read(got[key], 0)[0]))
            raise AssertionError("\n".join(msg))
