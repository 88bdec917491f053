"""CPU check of the identities behind the CUDA kernels: tests/host/kmer_identities.cu compiles the SAME __host__ __device__
helpers the kernels use (gx_internal.cuh: 2-bit packing into 64-bit words, word-parallel reverse complement, canonical =
integer min, 16-bit edge masks, neighbour reconstruction from (key, type, base)) as host code and emulates the aggregation;
the result must equal the oracle's literal byte-wise restatement of the reference. Runs without a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O

SRC = os.path.join(ROOT, "tests", "host", "kmer_identities.cu")


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host") / "kmer_identities")
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-o", out, SRC])
    return out


def oracle_nodes(k, reads):
    text = b"".join(b"%d\t%s\n" % (4 * i + 2, r) for i, r in enumerate(reads))
    table = O.build_graph(k, text)
    out = {}
    for key, node in table.items():
        out[key.hex()] = (int(node.coverage), [sorted(e[4:].hex() for e in node.edges[t]) for t in range(4)])
    return out


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 16, 21, 31, 32, 33, 55, 63, 64, 65, 91, 96, 97, 128])
def test_word_parallel_identities_match_oracle(k, host_binary):
    rng = np.random.default_rng(500 + k)
    genome = bytes(rng.choice(list(b"ACGT"), size=max(200, 3 * k)).tolist())
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    reads = []
    for i in range(25):
        L = int(rng.integers(k + 1, k + 30))
        s = int(rng.integers(0, len(genome) - L + 1))
        r = genome[s: s + L]
        if i % 2:
            r = bytes(comp[c] for c in reversed(r))
        if i % 5 == 0:
            r = r.lower()
        reads.append(r)
    # even k: palindromes; low complexity: self loops
    reads += [b"ACGT" * (k // 2 + 4), b"A" * (k + 6), b"AT" * (k + 3)]
    inp = b"".join(b"%d\t%s\n" % (k, r) for r in reads)
    res = subprocess.run([host_binary], input=inp, capture_output=True, check=True)
    got = {}
    for line in res.stdout.decode().splitlines():
        key, count, *lists = line.split(" ")
        got[key] = (int(count), [sorted(l.split(",")) if l != "-" else [] for l in lists])
    assert got == oracle_nodes(k, reads)
