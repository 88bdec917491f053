"""Host-side plumbing for the multi-GPU graph build: one process per GPU, reads sharded by byte range,
k-mers hash-partitioned to their owner GPU inside libgenomix_gb (all-to-all-v over NVLink; NCCL for the collectives). torch.distributed is used
only to bootstrap (broadcast of the NCCL unique id) -- any backend works for that, gloo included.

Reference counterpart: the partition layout of JobGen (genomix-hyracks/.../graph/job/JobGen.java:61-79: one
HDFS split per partition thread) and the M:N hash connector of JobGenBuildBrujinGraph.java:132-133.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_bytes: int, rank: int, world: int) -> Tuple[int, int]:
    """Nominal byte range of `rank`'s split (equal parts, like HDFS splits before line alignment)."""
    return (n_bytes * rank) // world, (n_bytes * (rank + 1)) // world


def shard_lines(text, rank: int, world: int):
    """This rank's whole lines of `text` (bytes-like or numpy uint8 array): the nominal split moved forward to the
    next line start at both ends, exactly how LineRecordReader treats split boundaries -- every line lands in
    exactly one shard."""
    mv = memoryview(text)
    n = len(mv)
    lo, hi = shard_range(n, rank, world)

    def align(pos: int) -> int:
        if pos <= 0:
            return 0
        if pos >= n:
            return n
        # a line starting exactly at pos belongs to this split only if the previous byte is a newline
        if mv[pos - 1] == 0x0A:
            return pos
        chunk = 1 << 16
        p = pos
        while p < n:
            seg = bytes(mv[p: p + chunk])
            i = seg.find(b"\n")
            if i >= 0:
                return p + i + 1
            p += chunk
        return n

    a, b = align(lo), align(hi)
    return text[a:b]


def bootstrap_nccl(gb, dist, device=None) -> None:
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed broadcasts it."""
    import numpy as np
    import torch
    rank = dist.get_rank()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.from_numpy(gb.mg_unique_id().copy())
    if device is not None:
        uid = uid.to(device)
    dist.broadcast(uid, 0)
    gb.mg_init(np.ascontiguousarray(uid.cpu().numpy()))


def build_graph_distributed(kmer_length: int, text_shard, dist, device_index: int, expected_kmers: int = 0) -> bytes:
    """Collective one-call build: every rank passes its shard, gets back the records of the nodes it owns."""
    import torch
    from .graphbuild import GraphBuilder
    world, rank = dist.get_world_size(), dist.get_rank()
    with GraphBuilder(kmer_length, device=device_index, rank=rank, n_ranks=world, expected_kmers=expected_kmers) as gb:
        if world > 1:
            bootstrap_nccl(gb, dist, torch.device("cuda", device_index) if dist.get_backend() == "nccl" else None)
        gb.push_lines(text_shard)
        if world > 1:
            gb.mg_exchange()
        gb.finish()
        return gb.records()
