"""Host side of the fastq front end: cut (paired) fastq files into whole-record chunks and feed gx_push_fastq.
Stands where the reference has GenomixDriver.setupHDFSInput / convertAndUploadFastqToHDFS / openFile
(genomix/genomix-driver/src/main/java/edu/uci/ics/genomix/driver/GenomixDriver.java:621-735); the conversion itself (ids,
trim, pairing, validation) runs on the GPU inside the library."""
from __future__ import annotations

import gzip
from typing import Iterator, Optional, Tuple

import numpy as np


def open_fastq(path: str):
    """openFile (GenomixDriver.java:723-735): '.gz' files are gunzipped."""
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def _take_records(buf: bytearray, n_lines_wanted: Optional[int], final: bool) -> Tuple[int, int]:
    """Largest prefix of buf made of whole 4-line records (at most n_lines_wanted lines): (bytes, lines)."""
    arr = np.frombuffer(buf, dtype=np.uint8)
    nl = np.flatnonzero(arr == 0x0A)
    n_lines = nl.size
    if final and len(buf) and (nl.size == 0 or nl[-1] != len(buf) - 1):
        n_lines += 1  # unterminated last line
    take = n_lines if final else (n_lines // 4) * 4
    if n_lines_wanted is not None:
        take = min(take, n_lines_wanted)
    if take == 0:
        return 0, 0
    if take > nl.size:  # includes the unterminated last line
        return len(buf), take
    return int(nl[take - 1]) + 1, take


def iter_chunks(path1: str, path2: Optional[str] = None, chunk_bytes: int = 128 << 20) -> Iterator[Tuple[bytes, Optional[bytes], int]]:
    """Yield (r1_chunk, r2_chunk_or_None, first_record) with the same whole records in both chunks."""
    f1 = open_fastq(path1)
    f2 = open_fastq(path2) if path2 else None
    b1, b2 = bytearray(), bytearray()
    first_record = 0
    eof1 = eof2 = False
    starved = False   # the last round produced nothing: read on even if a buffer is already chunk-sized
    while True:
        if not eof1 and (len(b1) < chunk_bytes or starved):
            d = f1.read(chunk_bytes)
            eof1 = len(d) == 0
            b1 += d
        if f2 is not None and not eof2 and (len(b2) < chunk_bytes or starved):
            d = f2.read(chunk_bytes)
            eof2 = len(d) == 0
            b2 += d
        final = eof1 and (f2 is None or eof2)
        n1, lines1 = _take_records(b1, None, eof1)
        if f2 is None:
            if n1 == 0 and final:
                return
            starved = n1 == 0
            if n1:
                yield bytes(b1[:n1]), None, first_record
                del b1[:n1]
                first_record += (lines1 + 2) // 4
            continue
        n2, lines2 = _take_records(b2, None, eof2)
        # one file exhausted while the other still has lines: the reference throws at this point
        # (GenomixDriver.java:684-687); waiting for more data from the exhausted side would never end
        if not final and ((eof1 and lines1 == 0 and lines2 > 0) or (eof2 and lines2 == 0 and lines1 > 0)):
            from .graphbuild import GenomixError
            raise GenomixError(-4,  # GX_ERR_FORMAT
                               f"IOException: Fastq files {path1} and {path2} didn't have the same number of lines!")
        lines = min(lines1, lines2) if not final else max(lines1, lines2)
        if not final:
            lines = (lines // 4) * 4
            n1, _ = _take_records(b1, lines, eof1)
            n2, _ = _take_records(b2, lines, eof2)
        if lines == 0 and final:
            return
        starved = lines == 0
        if lines:
            yield bytes(b1[:n1]), bytes(b2[:n2]), first_record   # unequal line counts at EOF surface as GX_ERR_FORMAT
            del b1[:n1]
            del b2[:n2]
            first_record += (lines + 2) // 4
        if final:
            return


def build_graph_from_fastq(kmer_length: int, path1: str, path2: Optional[str] = None, device: int = 0, **kw) -> bytes:
    from .graphbuild import GraphBuilder
    with GraphBuilder(kmer_length, device=device, **kw) as gb:
        for r1, r2, first in iter_chunks(path1, path2):
            gb.push_fastq(r1, r2, first)
        gb.finish()
        return gb.records()
