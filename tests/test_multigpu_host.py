"""CPU-only coverage of the N>1 host logic: line-aligned sharding and the torch.distributed bootstrap over gloo
(world_size 2). No GPU, no compute calls."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from genomix_b200 import multigpu


def test_shards_partition_the_lines():
    rng = np.random.default_rng(0)
    lines = [b"%d\t%s" % (4 * i + 2, bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(5, 80))).tolist())) for i in range(500)]
    for tail in (b"\n", b""):
        text = b"\n".join(lines) + tail
        for world in (1, 2, 3, 8, 64):
            shards = [bytes(multigpu.shard_lines(text, r, world)) for r in range(world)]
            assert b"".join(shards) == text
            for s in shards[:-1]:
                assert s == b"" or s.endswith(b"\n")
    arr = np.frombuffer(b"\n".join(lines) + b"\n", dtype=np.uint8)
    parts = [multigpu.shard_lines(arr, r, 4) for r in range(4)]
    assert sum(p.size for p in parts) == arr.size


class _FakeBuilder:
    """stands in for GraphBuilder: records what the bootstrap hands to the C ABI"""

    def __init__(self):
        self.uid = None

    def mg_unique_id(self):
        return (np.arange(128) * 7 % 251).astype(np.uint8)

    def mg_init(self, uid):
        self.uid = np.array(uid, copy=True)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gb = _FakeBuilder()
    multigpu.bootstrap_nccl(gb, dist)
    # every rank's shard, gathered: the host-side contract of the distributed build
    text = b"".join(b"%d\tACGTACGT\n" % i for i in range(100))
    shard = bytes(multigpu.shard_lines(text, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    q.put((rank, gb.uid.tolist(), b"".join(gathered) == text))
    dist.destroy_process_group()


def test_bootstrap_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (np.arange(128) * 7 % 251).astype(np.uint8).tolist()
    for rank, uid, ok in out:
        assert uid == want and ok
