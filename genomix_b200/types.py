"""Host-side mirrors of the reference's wire types for the graph-build path: Kmer, VKmer, VKmerList,
ReadHeadInfo, Node -- decode / encode / toString only (what a consumer of the record stream needs).

Reference (paths relative to /root/reference/genomix/genomix-data/src/main/java/edu/uci/ics/genomix/data/):
  types/Kmer.java:225-242, types/VKmer.java:366-391, types/VKmerList.java:62-68,314-336,
  types/ReadHeadInfo.java:100-127,152-176,205-212,236-240, types/Node.java:408-487,525-537,
  types/ExternalableTreeSet.java:220-267, utils/KmerUtil.java:21-48, utils/GeneCode.java:24-27.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Tuple

import numpy as np

GENE_SYMBOL = "ACGT"
EDGE_NAMES = ("FF", "FR", "RF", "RR")  # EDGETYPE.java:6-9
_LUT = np.array([[GENE_SYMBOL[(b >> (2 * j)) & 3] for j in range(4)] for b in range(256)], dtype="U1")


def bytes_for_k(k: int) -> int:
    """KmerUtil.getByteNumFromK (KmerUtil.java:21-27)."""
    return (k + 3) // 4


def kmer_to_string(k: int, data: bytes) -> str:
    """KmerUtil.recoverKmerFrom (KmerUtil.java:34-48): letter i sits at bits 2*(i%4) of byte nb-1-i/4."""
    if k <= 0:
        return ""
    arr = np.frombuffer(data, dtype=np.uint8)[::-1]
    return "".join(_LUT[arr].reshape(-1)[:k])


def string_to_kmer(s: str) -> bytes:
    """Kmer.setFromStringBytes (Kmer.java:225-242); non-ACGT letters pack as A (GeneCode.java:29-50)."""
    k = len(s)
    nb = bytes_for_k(k)
    out = bytearray(nb)
    for i, ch in enumerate(s):
        code = {"A": 0, "C": 1, "G": 2, "T": 3}.get(ch.upper(), 0)
        out[nb - 1 - i // 4] |= code << (2 * (i % 4))
    return bytes(out)


@dataclass
class VKmer:
    k: int
    data: bytes

    @staticmethod
    def read(buf: bytes, off: int) -> Tuple["VKmer", int]:
        (k,) = struct.unpack_from(">i", buf, off)
        nb = bytes_for_k(k)
        return VKmer(k, bytes(buf[off + 4: off + 4 + nb])), off + 4 + nb

    def write(self) -> bytes:
        return struct.pack(">i", self.k) + self.data

    def __str__(self) -> str:
        return kmer_to_string(self.k, self.data)


@dataclass
class ReadHeadInfo:
    value: int
    this_seq: VKmer
    mate_seq: Optional[VKmer]

    @property
    def read_id(self) -> int:
        return self.value & ((1 << 35) - 1)

    @property
    def mate_id(self) -> int:
        return (self.value >> 35) & 1

    @property
    def library_id(self) -> int:
        return (self.value >> 36) & 0xF

    @property
    def offset(self) -> int:
        off = (self.value >> 40) & 0xFFFFFF
        return -(off & 0x7FFFFF) if off & 0x800000 else off

    @staticmethod
    def read(buf: bytes, off: int) -> Tuple["ReadHeadInfo", int]:
        flags = buf[off]
        (value,) = struct.unpack_from(">Q", buf, off + 1)
        this_seq, off = VKmer.read(buf, off + 9)
        mate = None
        if flags & 1:
            mate, off = VKmer.read(buf, off)
        return ReadHeadInfo(value, this_seq, mate), off

    def write(self) -> bytes:
        has_mate = self.mate_seq is not None and self.mate_seq.k > 0
        out = bytes([1 if has_mate else 0]) + struct.pack(">Q", self.value) + self.this_seq.write()
        return out + (self.mate_seq.write() if has_mate else b"")

    def __str__(self) -> str:
        return (f"{self.read_id}-{self.offset}_{self.mate_id}-{self.library_id} readSeq: {self.this_seq} "
                f"mateReadSeq: {self.mate_seq if self.mate_seq is not None else 'null'}")


def _read_head_set(buf: bytes, off: int) -> Tuple[List[ReadHeadInfo], int]:
    """ExternalableTreeSet.readFields (ExternalableTreeSet.java:220-234), whole-body form only."""
    if buf[off] != 1:
        raise ValueError("ReadHeadSet stored by path reference; graph build always writes the whole body")
    (n,) = struct.unpack_from(">i", buf, off + 1)
    off += 5
    out = []
    for _ in range(n):
        rh, off = ReadHeadInfo.read(buf, off)
        out.append(rh)
    return out, off


def java_float_str(f: float) -> str:
    """java.lang.Float.toString for non-negative finite floats."""
    f32 = np.float32(f)
    if f32 == 0:
        return "0.0"
    if 1e-3 <= float(f32) < 1e7:
        s = np.format_float_positional(f32, unique=True, trim="0")
        return s + "0" if s.endswith(".") else s
    mant, exp = np.format_float_scientific(f32, unique=True, trim="0").split("e")
    if mant.endswith("."):
        mant += "0"
    return f"{mant}E{int(exp)}"


@dataclass
class Node:
    edges: List[Optional[List[VKmer]]] = field(default_factory=lambda: [None, None, None, None])
    unflipped: Optional[List[ReadHeadInfo]] = None
    flipped: Optional[List[ReadHeadInfo]] = None
    internal_kmer: Optional[VKmer] = None
    coverage: Optional[float] = None

    @staticmethod
    def read(buf: bytes, off: int = 0) -> Tuple["Node", int]:
        """Node.readFields (Node.java:429-456)."""
        n = Node()
        active = buf[off]
        off += 1
        for et in range(4):
            if active & (1 << et):
                (cnt,) = struct.unpack_from(">i", buf, off)
                off += 4
                lst = []
                for _ in range(cnt):
                    v, off = VKmer.read(buf, off)
                    lst.append(v)
                n.edges[et] = lst
        if active & (1 << 4):
            n.unflipped, off = _read_head_set(buf, off)
        if active & (1 << 5):
            n.flipped, off = _read_head_set(buf, off)
        if active & (1 << 6):
            n.internal_kmer, off = VKmer.read(buf, off)
        if active & (1 << 7):
            (n.coverage,) = struct.unpack_from(">f", buf, off)
            off += 4
        return n, off

    def write(self) -> bytes:
        """Node.write (Node.java:408-427)."""
        active = 0
        body = b""
        for et in range(4):
            if self.edges[et]:
                active |= 1 << et
                body += struct.pack(">i", len(self.edges[et])) + b"".join(v.write() for v in self.edges[et])
        for bit, s in ((4, self.unflipped), (5, self.flipped)):
            if s:
                active |= 1 << bit
                body += b"\x01" + struct.pack(">i", len(s)) + b"".join(r.write() for r in s)
        if self.internal_kmer is not None and self.internal_kmer.k > 0:
            active |= 1 << 6
            body += self.internal_kmer.write()
        if self.coverage is not None:
            active |= 1 << 7
            body += struct.pack(">f", self.coverage)
        return bytes([active]) + body

    def __str__(self) -> str:
        """Node.toString (Node.java:525-537)."""
        out = ["{"]
        for et in range(4):
            lst = self.edges[et]
            out.append(f"{EDGE_NAMES[et]}:" + ("null" if lst is None else "[" + ",".join(str(v) for v in lst) + "]") + "\t")

        def rs(s):
            return "null" if s is None else "[" + ",".join(str(r) for r in s) + "]"

        out.append("5':" + rs(self.unflipped) + ", ~5':" + rs(self.flipped) + "\t")
        out.append("kmer:" + ("null" if self.internal_kmer is None else str(self.internal_kmer)) + "\t")
        out.append("cov:" + ("null" if self.coverage is None else java_float_str(self.coverage) + "x") + "}")
        return "".join(out)

    def canonical_bytes(self) -> bytes:
        """Serialisation with each edge list sorted by VKmer bytes: the order inside a VKmerList is
        java.util.HashSet iteration order in the reference (VKmerList.java:117-133) and is excluded from
        equality by the reference's own comparator (utils/TestUtils.java:67-181)."""
        n = Node([None if e is None else sorted(e, key=lambda v: v.write()) for e in self.edges],
                 self.unflipped, self.flipped, self.internal_kmer, self.coverage)
        return n.write()


def iter_records(stream: bytes) -> Iterator[Tuple[bytes, bytes]]:
    """Split the record stream (int32be recordLength | int32be keyLength | key | value) into (key, value)."""
    off, n = 0, len(stream)
    mv = memoryview(stream)
    while off < n:
        rec_len, key_len = struct.unpack_from(">ii", stream, off)
        yield bytes(mv[off + 8: off + 8 + key_len]), bytes(mv[off + 8 + key_len: off + 8 + rec_len])
        off += 8 + rec_len


def records_to_text(stream: bytes) -> List[str]:
    """key.toString() + '\\t' + value.toString() per record, the dump format of
    GenomixClusterManager.java:345-387 that the reference's graph-build tests compare."""
    out = []
    for key, value in iter_records(stream):
        vk, _ = VKmer.read(key, 0)
        node, _ = Node.read(value, 0)
        out.append(f"{vk}\t{node}")
    return out


def canonical_records(stream: bytes) -> dict:
    """{key bytes: canonical Node bytes} -- the multiset of records after canonical sorting."""
    out = {}
    for key, value in iter_records(stream):
        if key in out:
            raise ValueError(f"duplicate key in record stream: {key.hex()}")
        out[key] = Node.read(value, 0)[0].canonical_bytes()
    return out
