"""ctypes loader of oracle/gx_oracle.c (the fast C twin of oracle.py). TEST INFRASTRUCTURE ONLY --
see the header of oracle.py for who may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_bin", "libgx_oracle.so")


class OracleStats(C.Structure):
    _fields_ = [("lines", C.c_uint64), ("reads", C.c_uint64), ("occurrences", C.c_uint64), ("nodes", C.c_uint64),
                ("err", C.c_int32), ("err_line", C.c_uint64)]


class OracleError(Exception):
    def __init__(self, status, line):
        super().__init__(f"reference job would fail with status {status} at input line {line}")
        self.status = status
        self.line = line


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "gx_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.gxo_build.restype = C.c_int
        lib.gxo_build.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                  C.POINTER(OracleStats)]
        lib.gxo_free.argtypes = [C.c_void_p]
        lib.gxo_java_partition.restype = C.c_int
        lib.gxo_java_partition.argtypes = [C.c_char_p, C.c_int, C.c_int]
        _lib = lib
    return _lib


class Fingerprint(C.Structure):
    _fields_ = [("sum", C.c_uint64), ("xor", C.c_uint64), ("records", C.c_uint64), ("heads", C.c_uint64), ("edges", C.c_uint64),
                ("coverage_total", C.c_double), ("bad", C.c_int32)]

    def key(self):
        return (self.sum, self.xor, self.records, self.heads, self.edges, self.coverage_total)


def canonical_fingerprint(stream) -> Fingerprint:
    """Order-independent fingerprint of a record stream (see gx_oracle.c): equal iff the streams hold the same records after
    canonical sorting. Accepts bytes or a numpy uint8 array."""
    import numpy as np
    lib = load()
    arr = np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else np.ascontiguousarray(stream)
    fp = Fingerprint()
    lib.gxo_canonical_fingerprint.restype = C.c_int
    lib.gxo_canonical_fingerprint.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Fingerprint)]
    rc = lib.gxo_canonical_fingerprint(C.c_void_p(arr.ctypes.data), arr.size, C.byref(fp))
    if rc != 0:
        raise ValueError(f"malformed record stream (code {rc})")
    return fp


def edge_symmetry(stream):
    """(signed sum, edge count): the sum is 0 iff every edge X -t-> Y has its mirror Y -mirror(t)-> X in the stream."""
    import numpy as np
    lib = load()
    arr = np.frombuffer(stream, dtype=np.uint8) if isinstance(stream, (bytes, bytearray)) else np.ascontiguousarray(stream)
    s, n = C.c_uint64(), C.c_uint64()
    lib.gxo_edge_symmetry.restype = C.c_int
    lib.gxo_edge_symmetry.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rc = lib.gxo_edge_symmetry(C.c_void_p(arr.ctypes.data), arr.size, C.byref(s), C.byref(n))
    if rc != 0:
        raise ValueError(f"malformed record stream (code {rc})")
    return int(s.value), int(n.value)


def build_graph_records(k: int, text, n_threads: int = 1, as_numpy: bool = False):
    """Record stream (same framing as the CUDA path) and stats dict for `text` (bytes or numpy uint8 array)."""
    import numpy as np
    lib = load()
    arr = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text)
    out, out_len, st = C.c_void_p(), C.c_size_t(), OracleStats()
    rc = lib.gxo_build(C.c_void_p(arr.ctypes.data), arr.size, k, n_threads, C.byref(out), C.byref(out_len), C.byref(st))
    if rc != 0:
        raise OracleError(rc, int(st.err_line))
    try:
        data = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(out_len.value,)).copy() if as_numpy \
            else C.string_at(out, out_len.value)
    finally:
        lib.gxo_free(out)
    return data, {"lines": st.lines, "reads": st.reads, "occurrences": st.occurrences, "nodes": st.nodes}
