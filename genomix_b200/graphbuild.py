"""GraphBuilder: thin Python owner of one gx_ctx (include/genomix_gb.h). Every compute call goes through
the C ABI into the CUDA kernels; nothing here computes on the CPU.

Stands where the reference has GenomixHyracksDriver.runJob + JobGenBuildBrujinGraph
(genomix-hyracks/.../graph/driver/GenomixHyracksDriver.java:84-135, graph/job/JobGenBuildBrujinGraph.java:79-90).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterator, Optional

import numpy as np

from . import _lib
from ._lib import GxConfig, GxStats, STATUS_NAMES


class GenomixError(RuntimeError):
    """A gx_status < 0; `.status` is the code, the message is what the reference would have thrown."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


class GraphBuilder:
    def __init__(self, kmer_length: int, device: int = 0, rank: int = 0, n_ranks: int = 1,
                 expected_kmers: int = 0, chunk_bytes: int = 0, min_capacity: int = 0,
                 start_small: bool = False, table_regions: int = 0, stream_records: bool = False,
                 sort_output: bool = False):
        """expected_kmers: optional hint (distinct k-mers this rank will own); chunk_bytes: internal chunk size;
        min_capacity / start_small: smallest table and "no sizing heuristics" (tests of the growth path);
        table_regions: regions of the region-sorted build (0 = sized for L2); stream_records: records are serialised on
        demand while they are copied to the host (no device-resident record stream); sort_output: records in KmerPointable
        order, like the reference's part files (default: table-slot order)."""
        self._lib = _lib.load()
        cfg = GxConfig()
        cfg.abi_version = _lib.GX_ABI_VERSION
        cfg.kmer_length = kmer_length
        cfg.device = device
        cfg.rank = rank
        cfg.n_ranks = n_ranks
        cfg.expected_kmers = expected_kmers
        cfg.sort_output = 1 if sort_output else 0
        cfg.reserved[0] = chunk_bytes or int(os.environ.get("GENOMIX_GB_CHUNK", "0"))
        cfg.reserved[1] = min_capacity
        cfg.reserved[2] = (256 if start_small else 0) | (1 if stream_records else 0)
        cfg.reserved[3] = table_regions or int(os.environ.get("GENOMIX_GB_REGIONS", "0"))
        self.kmer_length = kmer_length
        self._ctx = C.c_void_p()
        st = self._lib.gx_create(C.byref(cfg), C.byref(self._ctx))
        if st != 0:
            msg = self._lib.gx_last_error(None).decode()
            self._ctx = None
            raise GenomixError(st, msg)

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, st: int) -> None:
        if st != 0:
            raise GenomixError(st, self._lib.gx_last_error(self._ctx).decode())

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.gx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self) -> None:
        self._check(self._lib.gx_reset(self._ctx))

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._lib.gx_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    # -- input ------------------------------------------------------------------------------
    def push_lines(self, text) -> None:
        """text: bytes / bytearray / numpy uint8 array / pinned torch uint8 tensor (host memory)."""
        ptr, n, keep = _host_ptr(text)
        self._check(self._lib.gx_push_lines(self._ctx, ptr, n))
        del keep

    def push_lines_device(self, dev_ptr: int, n_bytes: int) -> None:
        self._check(self._lib.gx_push_lines_device(self._ctx, C.c_void_p(dev_ptr), n_bytes))

    def push_fastq(self, r1, r2=None, first_record: int = 0) -> None:
        """Whole-record-aligned chunk(s) of uncompressed fastq (host buffers); r2 is the mate file's chunk holding the same
        records. Read ids are the reference's 4*i+2 with i counted from first_record (GenomixDriver.java:665-714)."""
        p1, n1, k1 = _host_ptr(r1)
        if r2 is None:
            p2, n2, k2 = None, 0, None
        else:
            p2, n2, k2 = _host_ptr(r2)
        self._check(self._lib.gx_push_fastq(self._ctx, p1, n1, p2, n2, first_record))
        del k1, k2

    def push_records(self, records) -> None:
        """Fold a stream of serialised `VKmer | Node` records (the framing records() returns) into the job: the merge
        half of the reference's aggregator (AggregateKmerAggregateFactory.java:128-144) on the GPU."""
        ptr, n, keep = _host_ptr(records)
        self._check(self._lib.gx_push_records(self._ctx, ptr, n))
        del keep

    def finish(self) -> None:
        self._check(self._lib.gx_finish(self._ctx))

    # -- multi-GPU exchange (n_ranks > 1) ------------------------------------------------------
    def mg_unique_id(self) -> np.ndarray:
        """128-byte NCCL unique id; call on rank 0 and broadcast it to the other ranks."""
        out = np.zeros(128, dtype=np.uint8)
        st = self._lib.gx_mg_unique_id(C.c_void_p(out.ctypes.data))
        if st != 0:
            raise GenomixError(st, "ncclGetUniqueId failed")
        return out

    def mg_init(self, unique_id) -> None:
        uid = np.ascontiguousarray(np.asarray(unique_id, dtype=np.uint8))
        assert uid.size == 128
        self._check(self._lib.gx_mg_init(self._ctx, C.c_void_p(uid.ctypes.data)))

    def mg_exchange(self) -> None:
        """Collective: route staged k-mer and read-head records to their owner GPUs and fold them in."""
        self._check(self._lib.gx_mg_exchange(self._ctx))

    # -- output -----------------------------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        return int(self._lib.gx_num_nodes(self._ctx))

    @property
    def record_bytes(self) -> int:
        return int(self._lib.gx_record_bytes(self._ctx))

    def records(self) -> bytes:
        """The whole record stream copied to the host."""
        n = self.record_bytes
        if n < 0:
            raise GenomixError(-9, "records() before finish()")
        buf = np.empty(max(n, 1), dtype=np.uint8)
        cursor, used = C.c_uint64(0), C.c_size_t(0)
        pos = 0
        while pos < n:
            self._check(self._lib.gx_next_records(self._ctx, C.byref(cursor), C.c_void_p(buf.ctypes.data + pos), n - pos,
                                                  C.byref(used)))
            if used.value == 0:
                break
            pos += used.value
        return buf[:n].tobytes()

    def iter_record_batches(self, batch_bytes: int = 64 << 20) -> Iterator[bytes]:
        buf = np.empty(batch_bytes, dtype=np.uint8)
        cursor, used = C.c_uint64(0), C.c_size_t(0)
        while True:
            self._check(self._lib.gx_next_records(self._ctx, C.byref(cursor), C.c_void_p(buf.ctypes.data), batch_bytes,
                                                  C.byref(used)))
            if used.value == 0:
                return
            yield buf[: used.value].tobytes()

    def records_device(self):
        """(device pointer of the record stream, device pointer of the n_nodes+1 record offsets)."""
        p, o = C.c_void_p(), C.c_void_p()
        self._check(self._lib.gx_records_device(self._ctx, C.byref(p), C.byref(o)))
        return p.value, o.value

    def iter_frames(self, frame_size: int) -> Iterator[bytes]:
        """Hyracks frames of (Kmer, Node) tuples (FrameTupleAppender.java:57-70 layout)."""
        frame = np.zeros(frame_size, dtype=np.uint8)
        cursor, n_tuples = C.c_uint64(0), C.c_int32(0)
        while True:
            frame[:] = 0
            self._check(self._lib.gx_next_frame(self._ctx, C.byref(cursor), C.c_void_p(frame.ctypes.data), frame_size,
                                                C.byref(n_tuples)))
            if n_tuples.value == 0:
                return
            yield frame.tobytes()

    def write_sequence_file(self, path: str, sync: bytes | None = None, n_parts: int = 0, part: int = 0) -> int:
        """Write `part-<part>` as an uncompressed SequenceFile v6 <VKmer,Node> (what KmerNodePairSequenceWriterFactory
        produces); returns the bytes written. n_parts == 0 writes every record."""
        written = C.c_uint64(0)
        sp = None
        if sync is not None:
            assert len(sync) == 16
            sp = C.cast(C.create_string_buffer(sync, 16), C.c_void_p)
        self._check(self._lib.gx_write_sequence_file(self._ctx, path.encode(), sp, n_parts, part, C.byref(written)))
        return int(written.value)

    def partition(self, n_parts: int) -> np.ndarray:
        """KmerPartitionComputerFactory.partition for every emitted record (Java hash, abs, % n_parts)."""
        out = np.empty(max(self.num_nodes, 1), dtype=np.int32)
        self._check(self._lib.gx_partition_records(self._ctx, n_parts, C.c_void_p(out.ctypes.data)))
        return out[: self.num_nodes]

    # -- introspection ------------------------------------------------------------------------
    def stats(self) -> dict:
        s = GxStats()
        self._check(self._lib.gx_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()

    def graph_statistics(self) -> dict:
        """Node counters of the reference's GraphStatistics job, reduced on the GPU (after finish())."""
        g = _lib.GxGraphStats()
        self._check(self._lib.gx_graph_statistics(self._ctx, C.byref(g)))
        return g.as_dict()

    def coverage_histogram(self) -> np.ndarray:
        """GraphStatistics' "coverage-bins": nodes per Math.round(coverage) value, 0 .. coverage_max."""
        n = self.graph_statistics()["coverage_max"] + 1
        out = np.zeros(n, dtype=np.uint64)
        self._check(self._lib.gx_coverage_histogram(self._ctx, C.c_void_p(out.ctypes.data), C.c_uint64(n)))
        return out

    def coverage_cutoff(self, iterations: int = 10) -> dict:
        """GenomixDriver.setCutoffCoverageByFittingMixture: {"cutoff": minimum coverage or 0, fitted mixture parameters}."""
        cut, em, nm, ns = C.c_int64(0), C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(self._lib.gx_coverage_cutoff(self._ctx, iterations, C.byref(cut), C.byref(em), C.byref(nm), C.byref(ns)))
        return {"cutoff": int(cut.value), "exp_mean": em.value, "normal_mean": nm.value, "normal_std": ns.value}

    def phase_ms(self) -> dict:
        arr = (C.c_float * 8)()
        self._check(self._lib.gx_phase_ms(self._ctx, C.byref(arr)))
        return {"parse": arr[0], "insert": arr[1], "exchange": arr[2], "finish": arr[3], "h2d": arr[4],
                "exchange_comm": arr[5], "exchange_insert": arr[6], "split": arr[7]}

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.gx_kernel_launches(self._ctx))


def _host_ptr(obj):
    """(pointer, n_bytes, keep-alive) for bytes-like / numpy / torch host buffers."""
    if isinstance(obj, (bytes, bytearray)):
        arr = np.frombuffer(obj, dtype=np.uint8)
        return C.c_void_p(arr.ctypes.data), arr.size, (obj, arr)
    if isinstance(obj, np.ndarray):
        arr = np.ascontiguousarray(obj).view(np.uint8).reshape(-1)
        return C.c_void_p(arr.ctypes.data), arr.size, arr
    if hasattr(obj, "data_ptr"):  # torch tensor on the host
        if obj.is_cuda:
            raise TypeError("device tensor passed to push_lines; use push_lines_device")
        t = obj.contiguous()
        return C.c_void_p(t.data_ptr()), t.numel() * t.element_size(), t
    raise TypeError(f"unsupported text buffer type {type(obj)}")


def build_graph(kmer_length: int, text, device: int = 0, **kw) -> bytes:
    """One-call graph build: text lines in, record stream out."""
    with GraphBuilder(kmer_length, device=device, **kw) as gb:
        gb.push_lines(text)
        gb.finish()
        return gb.records()
