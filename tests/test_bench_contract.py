"""bench.py contract on the CPU box: the reference arm runs, prints one JSON line with the required keys; the synthetic
workloads are deterministic and carry the reference's read ids."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from genomix_b200 import synth


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample-reads", "3000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "kmers/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] > 1e5


def test_synthetic_workloads_are_seeded_and_well_formed():
    w = synth.scaled(synth.CONFIGS["cfg3"], 50000)
    a = synth.readid_text(w, n_reads=200).tobytes()
    b = synth.readid_text(w, n_reads=200).tobytes()
    assert a == b
    lines = a.split(b"\n")[:-1]
    assert len(lines) == 200
    for i, ln in enumerate(lines):
        rid, m0, m1 = ln.split(b"\t")
        assert int(rid) == 4 * i + 2 and len(m0) == len(m1) == w.read_len and set(m0 + m1) <= set(b"ACGT")
    # multi-GPU shards: disjoint id ranges, independent reads, same genome
    s0 = synth.shard_text(synth.CONFIGS["cfg1"], 0, 50).tobytes().split(b"\n")[:-1]
    s1 = synth.shard_text(synth.CONFIGS["cfg1"], 1, 50).tobytes().split(b"\n")[:-1]
    assert [int(l.split(b"\t")[0]) for l in s0] == [4 * i + 2 for i in range(50)]
    assert [int(l.split(b"\t")[0]) for l in s1] == [4 * (50 + i) + 2 for i in range(50)]
    assert s0 != s1
    assert synth.occurrences(synth.CONFIGS["cfg2"]) == 184000080


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench
    job, insert, b_occ, b_dist = bench.algorithmic_bytes(31, 150, 1000, 100)
    assert abs(b_occ - 25.25) < 1e-9 and b_dist == 73
    assert abs(bench.algorithmic_bytes(55, 150, 1, 0)[2] - (150 / 96 + 32)) < 1e-9
    assert bench.algorithmic_bytes(91, 250, 0, 1)[3] == 150
    assert bench.algorithmic_bytes(21, 100, 0, 1)[3] == 67
