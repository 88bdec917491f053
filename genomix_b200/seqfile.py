"""Reader of uncompressed SequenceFile v6 files -- the consumer's view of the graph-build output
(pregelix BinaryVertexInputFormat opens part files with SequenceFile.Reader:
genomix/genomix-pregelix/src/main/java/edu/uci/ics/genomix/pregelix/base/BinaryVertexInputFormat.java:25,93-99).
The writer lives in the library (gx_write_sequence_file). Container format: hadoop-core 0.20.2 SequenceFile,
as observed in the reference's own fixtures.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Tuple

SYNC_ESCAPE = -1
SYNC_INTERVAL = 2000  # 100 * (4 + 16): SequenceFile.SYNC_INTERVAL


@dataclass
class SequenceFile:
    key_class: str
    value_class: str
    sync: bytes
    header_len: int
    records: List[Tuple[bytes, bytes]] = field(default_factory=list)
    record_offsets: List[int] = field(default_factory=list)   # file offset of each record's length word
    sync_offsets: List[int] = field(default_factory=list)     # file offset of each sync escape


def _read_vint(buf: bytes, off: int) -> Tuple[int, int]:
    """WritableUtils.readVInt"""
    first = struct.unpack_from(">b", buf, off)[0]
    if first >= -112:
        return first, off + 1
    neg = first < -120
    n = (-119 - first) if neg else (-111 - first)
    val = 0
    for i in range(n - 1):
        val = (val << 8) | buf[off + 1 + i]
    return (~val if neg else val), off + n


def read_sequence_file(data: bytes) -> SequenceFile:
    if data[:3] != b"SEQ" or data[3] != 6:
        raise ValueError("not a SequenceFile v6")
    off = 4
    n, off = _read_vint(data, off)
    key_class = data[off: off + n].decode(); off += n
    n, off = _read_vint(data, off)
    value_class = data[off: off + n].decode(); off += n
    compressed, block = data[off], data[off + 1]
    off += 2
    if compressed or block:
        raise ValueError("compressed SequenceFiles are not produced by graph build")
    (n_meta,) = struct.unpack_from(">i", data, off); off += 4
    for _ in range(n_meta):
        for _ in range(2):
            n, off = _read_vint(data, off); off += n
    sync = data[off: off + 16]; off += 16
    sf = SequenceFile(key_class, value_class, sync, off)
    while off < len(data):
        (rec_len,) = struct.unpack_from(">i", data, off)
        if rec_len == SYNC_ESCAPE:
            if data[off + 4: off + 20] != sync:
                raise ValueError(f"bad sync marker at {off}")
            sf.sync_offsets.append(off)
            off += 20
            continue
        (key_len,) = struct.unpack_from(">i", data, off + 4)
        sf.record_offsets.append(off)
        sf.records.append((data[off + 8: off + 8 + key_len], data[off + 8 + key_len: off + 8 + rec_len]))
        off += 8 + rec_len
    return sf


def expected_sync_offsets(sf: SequenceFile) -> List[int]:
    """Where SequenceFile.Writer.checkAndWriteSync (called at the start of every append) puts sync escapes for this
    file's record sizes: before a record when pos >= lastSyncPos + SYNC_INTERVAL, lastSyncPos starting at 0."""
    out = []
    pos, last = sf.header_len, 0
    for key, val in sf.records:
        if pos >= last + SYNC_INTERVAL:
            out.append(pos)
            pos += 20
            last = pos
        pos += 8 + len(key) + len(val)
    return out
