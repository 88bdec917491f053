"""CPU oracle for the Genomix graph-build path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module. The product path (genomix_b200/) never does; it fails loudly when the CUDA
library is missing.

This is a literal, byte-level restatement of the reference's Java algorithm (the reference is
100 % Java and there is no JVM here, so it cannot be run). Every function cites the reference
file:line it follows; paths are relative to /root/reference/genomix/.

  GH = genomix-hyracks/src/main/java/edu/uci/ics/genomix/hyracks
  GD = genomix-data/src/main/java/edu/uci/ics/genomix/data

Parity pin: reproduces all 9 golden files the reference's own tests hold for this path (all
k=3; tests/golden/, checked by tests/test_oracle_golden.py) under the reference's own comparison
rule (GD/utils/TestUtils.java:67-181), plus the unit-level known answers of KmerFixedTest,
VKmerFixedTest and ReadHeadInfoTest. For k > 4 (multi-byte keys) the canonical-orientation choice
depends on hadoop-core 0.20.2 `WritableComparator.compareBytes` (unsigned lexicographic; jar not
vendored under /root/reference), so multi-byte parity is pinned by code reading + KmerFixedTest
only: "parity unpinned by reference goldens for k > 4".

Pure-Python loops: use for small inputs only (the C twin oracle/gx_oracle.c is the fast one).
"""
from __future__ import annotations

import re
import struct
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------- GeneCode
# GD/utils/GeneCode.java:24-50  A=0 C=1 G=2 T=3; any other byte maps to 0 (the switch has no default)
_CODE = {ord("A"): 0, ord("a"): 0, ord("C"): 1, ord("c"): 1, ord("G"): 2, ord("g"): 2, ord("T"): 3, ord("t"): 3}
GENE_SYMBOL = "ACGT"


def code_from_symbol(ch: int) -> int:
    return _CODE.get(ch, 0)


def paired_code_from_symbol(ch: int) -> int:
    """GD/utils/GeneCode.java:52-61 (3 - code)."""
    return 3 - code_from_symbol(ch)


def byte_num_from_k(k: int) -> int:
    """GD/utils/KmerUtil.java:21-27."""
    return k // 4 + (1 if k % 4 else 0)


# ----------------------------------------------------------------------------- Kmer
def kmer_from_string_bytes(k: int, s: bytes, start: int) -> bytes:
    """GD/types/Kmer.java:225-242 setFromStringBytes (also VKmer.java:462-479)."""
    nb = byte_num_from_k(k)
    out = bytearray(nb)
    l = 0
    bitcount = 0
    bcount = nb - 1
    i = start
    while i < start + k and i < len(s):
        l |= (code_from_symbol(s[i]) << bitcount) & 0xFF
        bitcount += 2
        if bitcount == 8:
            out[bcount] = l
            bcount -= 1
            l = 0
            bitcount = 0
        i += 1
    if bcount >= 0:
        out[0] = l
    return bytes(out)


def kmer_reversed_from_string_bytes(k: int, s: bytes, start: int) -> bytes:
    """GD/types/Kmer.java:253-272 setReversedFromStringBytes (reverse complement, packed)."""
    nb = byte_num_from_k(k)
    out = bytearray(nb)
    l = 0
    bitcount = 0
    bcount = nb - 1
    i = start + k - 1
    while i >= start and i < len(s):
        l |= (paired_code_from_symbol(s[i]) << bitcount) & 0xFF
        bitcount += 2
        if bitcount == 8:
            out[bcount] = l
            bcount -= 1
            l = 0
            bitcount = 0
        i -= 1
    if bcount >= 0:
        out[0] = l
    return bytes(out)


def kmer_shift_with_next_code(k: int, kb: bytes, c: int) -> bytes:
    """GD/types/Kmer.java:292-303 shiftKmerWithNextCode + clearLeadBit :332-336."""
    nb = len(kb)
    b = bytearray(kb)
    for i in range(nb - 1, 0, -1):
        inn = b[i - 1] & 0x03
        b[i] = ((b[i] >> 2) & 0x3F) | (inn << 6)
    pos = ((k - 1) % 4) << 1
    b[0] = (((b[0] >> 2) & 0x3F) | (c << pos)) & 0xFF
    if k % 4 != 0:
        b[0] &= (1 << ((k % 4) << 1)) - 1
    return bytes(b)


def compare_bytes(a: bytes, b: bytes) -> int:
    """hadoop-core 0.20.2 WritableComparator.compareBytes (unsigned lexicographic, byte difference);
    used through BinaryComparable.compareTo by GH/graph/dataflow/ReadsKeyValueParserFactory.java:163,181.
    Pinned by VKmerFixedTest.java:549-550 (CTA < GTA)."""
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return a[i] - b[i]
    return len(a) - len(b)


def recover_kmer(k: int, kb: bytes) -> str:
    """GD/utils/KmerUtil.java:34-48 recoverKmerFrom (== Kmer/VKmer.toString)."""
    if k < 1 or len(kb) == 0:
        return ""
    out = []
    byte_id = len(kb) - 1
    cur = kb[byte_id]
    for g in range(k):
        if g % 4 == 0 and g > 0:
            byte_id -= 1
            cur = kb[byte_id]
        out.append(GENE_SYMBOL[(cur >> ((g % 4) * 2)) & 3])
    return "".join(out)


def vkmer_bytes(k: int, kb: bytes) -> bytes:
    """GD/types/VKmer.java:389-391 write: int k (big-endian) + ceil(k/4) bytes."""
    return struct.pack(">i", k) + kb


def vkmer_from_string(s: bytes) -> bytes:
    """GD/types/VKmer.java:140-142 setAsCopy(String) -> setFromStringBytes(len, bytes, 0)."""
    return vkmer_bytes(len(s), kmer_from_string_bytes(len(s), s, 0))


# ----------------------------------------------------------------------------- ReadHeadInfo
def make_uuid(mate: int, library: int, read_id: int, offset: int) -> int:
    """GD/types/ReadHeadInfo.java:100-127 makeUUID. Bits: offset:24 | library:4 | mate:1 | readId:35,
    but the readId guard masks to 29 bits (`~(-1l << (64-35))`, :110-113)."""
    if mate != (mate & ~(-1 << 63)):
        raise ValueError("mateId will lose bits")
    if library != (library & ~(-1 << 60)):
        raise ValueError("libraryId will lose bits")
    if read_id != (read_id & ~(-1 << 29)):
        raise ValueError(f"readId {read_id} will lose some of its bits when saved")
    if abs(offset) > (1 << 23) - 1:
        raise ValueError("offset will lose bits")
    if offset < 0:
        offset = -offset | (1 << 23)
    return ((offset << 40) + (library << 36) + (mate << 35) + read_id) & 0xFFFFFFFFFFFFFFFF


def uuid_fields(v: int):
    """GD/types/ReadHeadInfo.java:152-176 getMateId/getLibraryId/getReadId/getOffset."""
    read_id = v & ((1 << 35) - 1)
    mate = (v >> 35) & 1
    lib = (v >> 36) & 0xF
    off = (v >> 40) & 0xFFFFFF
    if off & (1 << 23):
        off = -(off & ((1 << 23) - 1))
    return mate, lib, read_id, off


@dataclass
class ReadHead:
    value: int  # the 64-bit uuid
    this_seq: bytes  # serialised VKmer (4-byte k + packed)
    mate_seq: bytes | None  # serialised VKmer or None

    def sort_key(self):
        """GD/types/ReadHeadInfo.java:247-264 compareTo: offset, library, mate, readId."""
        mate, lib, rid, off = uuid_fields(self.value)
        return (off, lib, mate, rid)

    def write(self) -> bytes:
        """GD/types/ReadHeadInfo.java:197-212: flags byte (bit0 = mate seq present and non-empty),
        long value, VKmer this, [VKmer mate]."""
        has_mate = self.mate_seq is not None and struct.unpack(">i", self.mate_seq[:4])[0] > 0
        out = bytes([1 if has_mate else 0]) + struct.pack(">Q", self.value) + self.this_seq
        if has_mate:
            out += self.mate_seq
        return out

    def to_string(self) -> str:
        """GD/types/ReadHeadInfo.java:236-240 toString."""
        mate, lib, rid, off = uuid_fields(self.value)
        k = struct.unpack(">i", self.this_seq[:4])[0]
        s = f"{rid}-{off}_{mate}-{lib} readSeq: {recover_kmer(k, self.this_seq[4:])} mateReadSeq: "
        if self.mate_seq is None:
            return s + "null"
        km = struct.unpack(">i", self.mate_seq[:4])[0]
        return s + recover_kmer(km, self.mate_seq[4:])


# ----------------------------------------------------------------------------- Node
FF, FR, RF, RR = 0, 1, 2, 3  # GD/types/EDGETYPE.java:6-9
MIRROR = {FF: RR, FR: FR, RF: RF, RR: FF}  # EDGETYPE.java:44-58
FORWARD, REVERSE = 0, 1  # GD/types/DIR.java
EDGE_NAMES = ["FF", "FR", "RF", "RR"]


def java_float_to_string(f) -> str:
    """java.lang.Float.toString for the non-negative finite values coverage takes."""
    f = np.float32(f)
    if f == 0:
        return "0.0"
    if 1e-3 <= float(f) < 1e7:
        s = np.format_float_positional(f, unique=True, trim="0")
        if s.endswith("."):
            s += "0"
        return s
    s = np.format_float_scientific(f, unique=True, trim="0")  # e.g. 1.2345678e+07
    mant, exp = s.split("e")
    if mant.endswith("."):
        mant += "0"
    return f"{mant}E{int(exp)}"


@dataclass
class Node:
    """GD/types/Node.java:123-128: 4 x VKmerList + 2 x ReadHeadSet + optional kmer + Float coverage."""
    edges: list = field(default_factory=lambda: [[], [], [], []])  # lists of serialised VKmer, insertion order
    unflipped: dict = field(default_factory=dict)  # sort_key -> ReadHead (TreeSet semantics: first kept)
    flipped: dict = field(default_factory=dict)
    coverage: np.float32 | None = None

    def write(self) -> bytes:
        """GD/types/Node.java:408-427 write + getActiveFields :466-487; VKmerList.write VKmerList.java:314-316;
        ExternalableTreeSet.write ExternalableTreeSet.java:236-267 with forceWriteEntireBody(true)
        (GH/graph/dataflow/AggregateKmerAggregateFactory.java:65): boolean true, int size, elements."""
        active = 0
        body = b""
        for et in range(4):
            if len(self.edges[et]) > 0:
                active |= 1 << et
                body += struct.pack(">i", len(self.edges[et])) + b"".join(self.edges[et])
        for bit, s in ((1 << 4, self.unflipped), (1 << 5, self.flipped)):
            if len(s) > 0:
                active |= bit
                body += b"\x01" + struct.pack(">i", len(s))
                for key in sorted(s):
                    body += s[key].write()
        # internalKmer (bit 6) is never set by graph build (pregelix fills it on load)
        if self.coverage is not None:
            active |= 1 << 7
            body += struct.pack(">f", float(self.coverage))
        return bytes([active]) + body

    def to_string(self) -> str:
        """GD/types/Node.java:525-537 toString as printed after a readFields round trip (absent lists are
        null: Node.reset; GD/cluster/GenomixClusterManager.java:345-387 dumps key.toString()\\tvalue.toString())."""
        parts = ["{"]
        for et in range(4):
            if len(self.edges[et]) == 0:
                lst = "null"
            else:
                lst = "[" + ",".join(recover_kmer(struct.unpack(">i", e[:4])[0], e[4:]) for e in self.edges[et]) + "]"
            parts.append(f"{EDGE_NAMES[et]}:{lst}\t")

        def rs(s):
            if len(s) == 0:
                return "null"
            return "[" + ",".join(s[key].to_string() for key in sorted(s)) + "]"

        parts.append("5':" + rs(self.unflipped))
        parts.append(", ~5':" + rs(self.flipped) + "\t")
        parts.append("kmer:null\t")
        parts.append("cov:" + ("null" if self.coverage is None else java_float_to_string(self.coverage) + "x") + "}")
        return "".join(parts)


def aggregate_into(acc: Node, inp: Node, first: bool) -> None:
    """GH/graph/dataflow/AggregateKmerAggregateFactory.java:93-125 (init) and :128-144 (aggregate):
    per edge type VKmerList.unionUpdate (set union by VKmer bytes, GD/types/VKmerList.java:117-133),
    ReadHeadSet.unionUpdate (TreeSet.addAll: an equal element already present is kept),
    coverage assigned on init, added (float) on aggregate."""
    for et in range(4):
        for e in inp.edges[et]:
            if e not in acc.edges[et]:
                acc.edges[et].append(e)
    for dst, src in ((acc.unflipped, inp.unflipped), (acc.flipped, inp.flipped)):
        for key, rh in src.items():
            if key not in dst:
                dst[key] = rh
    if first:
        acc.coverage = np.float32(inp.coverage)
    else:
        acc.coverage = np.float32(acc.coverage + np.float32(inp.coverage))


# ----------------------------------------------------------------------------- parser (A1, A2, A4)
class GraphBuildError(Exception):
    """Stands for the unchecked exceptions that kill the reference job."""


_GENE = re.compile(rb"[ACGTacgt]+")
_LONG = re.compile(rb"[+-]?[0-9]+")


def java_split_tab(line: bytes) -> list:
    """String.split("\\t") with limit 0: trailing empty strings are removed."""
    parts = line.split(b"\t")
    while len(parts) > 1 and parts[-1] == b"":
        parts.pop()
    return parts


def split_reads(k: int, read_head: ReadHead, letters: bytes, emit) -> None:
    """GH/graph/dataflow/ReadsKeyValueParserFactory.java:150-196 SplitReads, with
    setEdgesForCurAndNext :209-233 and writeToFrame :198-207. `emit(key_bytes, Node)` stands for InsertToFrame."""
    if k >= len(letters):
        raise GraphBuildError(f"kmersize (k={k}) is larger than the read length ({len(letters)})")
    cur = Node(coverage=np.float32(1))
    cur_f = kmer_from_string_bytes(k, letters, 0)
    cur_r = kmer_reversed_from_string_bytes(k, letters, 0)
    cur_dir = FORWARD if compare_bytes(cur_f, cur_r) <= 0 else REVERSE
    mate, lib, rid, _ = uuid_fields(read_head.value)
    if cur_dir == FORWARD:
        cur.unflipped[read_head.sort_key()] = read_head
    else:
        rh = ReadHead(make_uuid(mate, lib, rid, k - 1), read_head.this_seq, read_head.mate_seq)  # resetOffset(K-1) :168
        cur.flipped[rh.sort_key()] = rh
    nxt = Node(coverage=np.float32(1))
    nxt_f = cur_f
    for i in range(k, len(letters)):
        nxt_f = kmer_shift_with_next_code(k, nxt_f, code_from_symbol(letters[i]))
        nxt_r = kmer_reversed_from_string_bytes(k, letters, i - k + 1)
        nxt_dir = FORWARD if compare_bytes(nxt_f, nxt_r) <= 0 else REVERSE
        # setEdgesForCurAndNext :209-233
        if cur_dir == FORWARD and nxt_dir == FORWARD:
            cur.edges[FF].append(vkmer_bytes(k, nxt_f))
            nxt.edges[RR].append(vkmer_bytes(k, cur_f))
        elif cur_dir == FORWARD and nxt_dir == REVERSE:
            cur.edges[FR].append(vkmer_bytes(k, nxt_r))
            nxt.edges[FR].append(vkmer_bytes(k, cur_f))
        elif cur_dir == REVERSE and nxt_dir == FORWARD:
            cur.edges[RF].append(vkmer_bytes(k, nxt_f))
            nxt.edges[RF].append(vkmer_bytes(k, cur_r))
        else:
            cur.edges[RR].append(vkmer_bytes(k, nxt_r))
            nxt.edges[FF].append(vkmer_bytes(k, cur_r))
        emit(cur_f if cur_dir == FORWARD else cur_r, cur)
        cur_f, cur_r, cur, cur_dir = nxt_f, nxt_r, nxt, nxt_dir
        nxt = Node(coverage=np.float32(1))
    emit(cur_f if cur_dir == FORWARD else cur_r, cur)


def parse_line(k: int, line: bytes, emit) -> None:
    """GH/graph/dataflow/ReadsKeyValueParserFactory.java:95-148 parse. libraryId is always 0 (:98-106:
    Matcher.group(0) without matches() throws IllegalStateException, which is caught)."""
    library = 0
    raw = java_split_tab(line)
    if len(raw) == 2:
        id_txt, mate0, mate1 = raw[0], raw[1], None
    elif len(raw) == 3:
        id_txt, mate0, mate1 = raw[0], raw[1], raw[2]
    else:
        raise GraphBuildError(f"input format is not correct! saw {line!r} which has {len(raw)} elements")
    if not _LONG.fullmatch(id_txt):
        raise GraphBuildError(f"NumberFormatException: {id_txt!r}")
    read_id = int(id_txt)
    if not (-(1 << 63) <= read_id < (1 << 63)):
        raise GraphBuildError(f"NumberFormatException: {id_txt!r}")

    def uuid(mate):
        try:
            return make_uuid(mate, library, read_id, 0)
        except ValueError as e:
            raise GraphBuildError(str(e))

    if _GENE.fullmatch(mate0):
        this_seq = vkmer_from_string(mate0)
        mate_seq = vkmer_from_string(mate1) if mate1 is not None else None
        split_reads(k, ReadHead(uuid(0), this_seq, mate_seq), mate0, emit)
    if mate1 is not None and _GENE.fullmatch(mate1):
        # :140-145 -- the mate sequence is the raw mate-0 text even if it failed the regex
        split_reads(k, ReadHead(uuid(1), vkmer_from_string(mate1), vkmer_from_string(mate0)), mate1, emit)


def split_lines(text: bytes) -> list:
    """hadoop TextInputFormat / LineReader.readLine (hadoop-core 0.20.2, org.apache.hadoop.util.LineReader): a record ends
    at \\n, \\r or \\r\\n; a final unterminated line is still a record."""
    lines = re.split(rb"\r\n|\n|\r", text)
    if lines and lines[-1] == b"":
        lines.pop()
    return lines


def build_graph(k: int, text: bytes) -> dict:
    """The whole job of GH/graph/job/JobGenBuildBrujinGraph.java:79-90: parse -> sort -> local aggregate ->
    hash repartition -> global aggregate. Sort + two-level grouping only make equal keys meet; the result is
    the per-key aggregate, returned as {key bytes: Node}."""
    table: dict = {}

    def emit(key: bytes, node: Node):
        acc = table.get(key)
        if acc is None:
            acc = Node()
            table[key] = acc
            aggregate_into(acc, node, True)
        else:
            aggregate_into(acc, node, False)

    for line in split_lines(text):
        parse_line(k, line, emit)
    return table


def graph_text_lines(k: int, table: dict) -> list:
    """key.toString() + '\\t' + value.toString() per record (GD/cluster/GenomixClusterManager.java:345-387)."""
    return [recover_kmer(k, key) + "\t" + node.to_string() for key, node in table.items()]


def graph_records(k: int, table: dict) -> dict:
    """{VKmer key bytes: Node bytes} as KmerNodePairSequenceWriterFactory would append them
    (GH/graph/dataflow/KmerNodePairSequenceWriterFactory.java:79-94)."""
    return {vkmer_bytes(k, key): node.write() for key, node in table.items()}


# ----------------------------------------------------------------------------- A0 fastq conversion
def java_trim(b: bytes) -> bytes:
    """String.trim(): strips code points <= U+0020 from both ends."""
    s, e = 0, len(b)
    while s < e and b[s] <= 0x20:
        s += 1
    while e > s and b[e - 1] <= 0x20:
        e -= 1
    return b[s:e]


def java_read_lines(data: bytes) -> list:
    """BufferedReader.readLine(): lines end at \n, \r or \r\n; a final unterminated line counts."""
    lines = re.split(rb"\r\n|\n|\r", data)
    if lines and lines[-1] == b"":
        lines.pop()
    return lines


def fastq_to_readids(mate1: bytes, mate2: bytes | None = None) -> bytes:
    """genomix-driver/.../GenomixDriver.java:665-714 convertAndUploadFastqToHDFS: 0-based line j with j % 4 == 1 is
    written as '<j+1>\t<line.trim()>[\t<mateLine.trim()>]\n'; paired files of different line counts are an IOException."""
    l1 = java_read_lines(mate1)
    out = []
    if mate2 is not None:
        l2 = java_read_lines(mate2)
        if len(l1) != len(l2):
            raise GraphBuildError("IOException: Fastq files didn't have the same number of lines")
        for j, (a, b) in enumerate(zip(l1, l2)):
            if j % 4 == 1:
                out.append(b"%d\t%s\t%s\n" % (j + 1, java_trim(a), java_trim(b)))
    else:
        for j, a in enumerate(l1):
            if j % 4 == 1:
                out.append(b"%d\t%s\n" % (j + 1, java_trim(a)))
    return b"".join(out)


# ----------------------------------------------------------------------------- A8 partition hash
def java_partition(key: bytes, n_parts: int) -> int:
    """GH/data/primitive/KmerPartitionComputerFactory.java:28-33,39-52."""
    h = 1
    for b in key:
        sb = b - 256 if b >= 128 else b
        h = (31 * h + sb) & 0xFFFFFFFF
    if h >= 1 << 31:
        h -= 1 << 32
    if h < 0:
        h = -(h + 1)
    return h % n_parts


# ----------------------------------------------------------------------------- reference comparison rule
def compare_unordered(expected_lines, actual_lines, unordered=(1, 2, 3, 4)) -> None:
    """GD/utils/TestUtils.java:67-181 compareFilesWithUnOrderedFields(expected, actual, sorted=true, {1,2,3,4}):
    sort lines ignoring the unordered fields, then compare: ordered fields verbatim, unordered fields as sorted
    lists of ACGT tokens. Raises AssertionError on mismatch."""
    uo = set(unordered)

    def sort_key(line):
        parts = line.split("\t")
        return [p for i, p in enumerate(parts) if i not in uo]

    exp = sorted(expected_lines, key=sort_key)
    act = sorted(actual_lines, key=sort_key)
    assert len(exp) == len(act), f"line count differs: expected {len(exp)} actual {len(act)}"
    for e, a in zip(exp, act):
        if e == a:
            continue
        fe, fa = e.split("\t"), a.split("\t")
        assert len(fe) == len(fa), f"field count differs:\n< {e}\n> {a}"
        for i, (x, y) in enumerate(zip(fe, fa)):
            if i in uo:
                tx = sorted(t for t in re.split(r"[^ATCG]+", x))
                ty = sorted(t for t in re.split(r"[^ATCG]+", y))
                assert tx == ty, f"unordered field {i} differs:\n< {e}\n> {a}"
            else:
                assert x == y, f"field {i} differs:\n< {e}\n> {a}"


# ---------------------------------------------------------------------------------------------
# Post-build statistics: GraphStatistics.map and the driver's FittingMixture cut-off, restated literally.

def graph_statistics(k: int, nodes: dict) -> dict:
    """GraphStatistics.map (genomix-hadoop/.../utils/GraphStatistics.java:78-131) over {key bytes: Node}: the Hadoop counter
    groups as nested dicts -- "totals", "maximum" and one "<name>-bins" group per updateStats name (:133-141)."""
    totals, maximum, bins = {}, {}, {}

    def incr(name, by=1):
        totals[name] = totals.get(name, 0) + by

    def update(name, value):
        bins.setdefault(name, {})
        bins[name][value] = bins[name].get(value, 0) + 1
        incr(name, value)
        maximum[name] = max(maximum.get(name, 0), value)

    mean, std = 0.0, 1000.0   # COVERAGE_DIST_MEAN / _STD (:75-76)
    for key, node in nodes.items():
        sizes = [len(e) if e else 0 for e in node.edges]
        out_deg, in_deg = sizes[0] + sizes[1], sizes[2] + sizes[3]   # DIR.FORWARD = {FF, FR}, DIR.REVERSE = {RF, RR} (Node.java:820-836)
        n_un, n_fl = len(node.unflipped or []), len(node.flipped or [])
        incr("nodes")
        update("degree", in_deg + out_deg)
        update("kmerLength", k)                      # the internal kmer of a graph-build node is empty: the key's length (:88-89)
        cov_round = int(java_round(node.coverage))
        update("coverage", cov_round)
        update("unflippedReadIds", n_un)
        update("flippedReadIds", n_fl)
        seed = k * (n_un + n_fl)                     # Node.calculateSeedScore (Node.java:859-862)
        in_window = mean - std <= node.coverage <= mean + std
        if in_window:
            update("scaffoldSeedScore", seed)
        for et, name in enumerate(("FF", "FR", "RF", "RR")):
            for e in (node.edges[et] or []):
                if e.write() == key:
                    incr("selfEdge-" + name)
        if in_deg == 1 and out_deg == 1:
            incr("pathNode")
        for d, deg in (("FORWARD", out_deg), ("REVERSE", in_deg)):
            if deg == 0:
                incr("tips-" + d)
            else:
                update("kmerLength-with-" + d, k)
                update("coverage-with-" + d, cov_round)
                if in_window:
                    update("scaffoldSeedScore-with-" + d, seed)
        if in_deg == 0 and out_deg == 0:
            incr("tips-BOTH")
        if (in_deg == 0) != (out_deg == 0):
            incr("tips-ONE")
    return {"totals": totals, "maximum": maximum, "bins": bins}


def java_round(x: float) -> int:
    """Math.round(float): floor(x + 0.5)"""
    import math
    return int(math.floor(x + 0.5))


def fitting_mixture(data, max_coverage: float, iterations: int):
    """FittingMixture.fittingMixture (genomix-driver/.../mixture/model/FittingMixture.java:92-216), one entry of `data` per
    node exactly like GraphStatistics.getCoverageStats builds it; commons-math3 densities. Returns
    (cutoff, exp_mean, normal_mean, normal_std)."""
    import math

    def exp_density(mean, x):
        return 0.0 if x < 0 else math.exp(-x / mean) / mean

    def normal_density(mean, sd, x):
        z = (x - mean) / sd
        return math.exp(-0.5 * z * z) / (sd * math.sqrt(2 * math.pi))

    n = len(data)
    e_mean, n_mean, n_sd, p_exp, p_norm = 5.0, 20.0, 5.0, 0.5, 0.5
    m_exp, m_norm = [0.0] * n, [0.0] * n

    def expectation():
        nonlocal p_exp, p_norm
        for i in range(n):
            a = exp_density(e_mean, data[i]) * p_exp
            if a == 0:
                a = 0.000000001
            b = normal_density(n_mean, n_sd, data[i]) * p_norm
            if b == 0:
                b = 0.000000001
            m_exp[i], m_norm[i] = a / (a + b), b / (a + b)
        es, ns = math.fsum(m_exp), math.fsum(m_norm)
        p_exp, p_norm = es / (es + ns), ns / (es + ns)
        return es, ns

    es, ns = expectation()
    for _ in range(iterations):
        e_mean = math.fsum(m_exp[i] * data[i] for i in range(n)) / es
        n_mean = math.fsum(m_norm[i] * data[i] for i in range(n)) / ns
        n_sd = math.sqrt(math.fsum(m_norm[i] * (data[i] - n_mean) ** 2 for i in range(n)) / ns)
        if n_sd == 0:
            n_sd = float(np.float32(0.000000001))
        es, ns = expectation()
    cov = 1.0
    while cov < max_coverage:
        if p_exp * exp_density(e_mean, cov) < p_norm * normal_density(n_mean, n_sd, cov):
            return int(cov), e_mean, n_mean, n_sd
        cov += 1
    return 0, e_mean, n_mean, n_sd
