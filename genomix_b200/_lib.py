"""ctypes binding of libgenomix_gb.so (include/genomix_gb.h). Fails loudly: there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GENOMIX_GB_LIB") or os.path.join(_HERE, "libgenomix_gb.so")  # env: tuning variants only

GX_ABI_VERSION = 2

# every symbol include/genomix_gb.h declares (tests check that the .so exports each one)
EXPORTS = [
    "gx_abi_version", "gx_create", "gx_destroy", "gx_reset", "gx_last_error", "gx_get_stats",
    "gx_push_lines", "gx_push_lines_device", "gx_push_fastq", "gx_push_records", "gx_finish",
    "gx_num_nodes", "gx_record_bytes", "gx_next_records", "gx_records_device", "gx_next_frame",
    "gx_partition_records", "gx_write_sequence_file", "gx_graph_statistics", "gx_coverage_histogram", "gx_coverage_cutoff",
    "gx_mg_unique_id", "gx_mg_init", "gx_mg_exchange",
    "gx_phase_ms", "gx_kernel_launches", "gx_set_stream",
]

STATUS_NAMES = {
    0: "GX_OK", -1: "GX_ERR_INVALID", -2: "GX_ERR_CUDA", -3: "GX_ERR_NOMEM", -4: "GX_ERR_FORMAT",
    -5: "GX_ERR_NUMBER", -6: "GX_ERR_READ_TOO_SHORT", -7: "GX_ERR_READID_RANGE", -8: "GX_ERR_BUFFER",
    -9: "GX_ERR_STATE",
}


class GxConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("kmer_length", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32),
        ("n_ranks", C.c_int32), ("sort_output", C.c_int32), ("expected_kmers", C.c_uint64),
        ("reserved", C.c_uint64 * 4),
    ]


class GxStats(C.Structure):
    _fields_ = [
        ("lines", C.c_uint64), ("reads", C.c_uint64), ("bases", C.c_uint64), ("kmer_occurrences", C.c_uint64),
        ("distinct_kmers", C.c_uint64), ("read_heads", C.c_uint64), ("record_bytes", C.c_uint64),
        ("table_capacity", C.c_uint64), ("table_grows", C.c_uint64), ("exchanged_records", C.c_uint64),
        ("split_redos", C.c_uint64), ("reserved", C.c_uint64 * 5),
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


class GxGraphStats(C.Structure):
    _fields_ = [
        ("nodes", C.c_uint64), ("degree_total", C.c_uint64), ("degree_max", C.c_uint64), ("degree_bins", C.c_uint64 * 17),
        ("coverage_total", C.c_uint64), ("coverage_max", C.c_uint64), ("coverage_bins", C.c_uint64 * 257),
        ("unflipped_read_ids", C.c_uint64), ("flipped_read_ids", C.c_uint64), ("self_edges", C.c_uint64 * 4),
        ("path_nodes", C.c_uint64), ("tips_forward", C.c_uint64), ("tips_reverse", C.c_uint64), ("tips_both", C.c_uint64),
        ("tips_one", C.c_uint64),
        ("kmer_length_total", C.c_uint64), ("kmer_length_max", C.c_uint64),
        ("nodes_with_dir", C.c_uint64 * 2), ("coverage_with_dir_total", C.c_uint64 * 2), ("coverage_with_dir_max", C.c_uint64 * 2),
        ("seed_nodes", C.c_uint64), ("seed_score_total", C.c_uint64), ("seed_score_max", C.c_uint64),
        ("seed_nodes_with_dir", C.c_uint64 * 2), ("seed_score_with_dir_total", C.c_uint64 * 2),
        ("seed_score_with_dir_max", C.c_uint64 * 2),
    ]

    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = int(v) if isinstance(v, int) else [int(x) for x in v]
        return out


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library built in-tree by `make -C genomix_b200/csrc` (or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C genomix_b200/csrc` (nvcc, sm_100a). "
            "genomix_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, u8p, sz, u64, i32 = C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_int32
    sig = {
        "gx_abi_version": (C.c_int, []),
        "gx_create": (C.c_int, [C.POINTER(GxConfig), C.POINTER(vp)]),
        "gx_destroy": (None, [vp]),
        "gx_reset": (C.c_int, [vp]),
        "gx_last_error": (C.c_char_p, [vp]),
        "gx_get_stats": (C.c_int, [vp, C.POINTER(GxStats)]),
        "gx_push_lines": (C.c_int, [vp, u8p, sz]),
        "gx_push_lines_device": (C.c_int, [vp, u8p, sz]),
        "gx_push_fastq": (C.c_int, [vp, u8p, sz, u8p, sz, u64]),
        "gx_push_records": (C.c_int, [vp, u8p, sz]),
        "gx_finish": (C.c_int, [vp]),
        "gx_num_nodes": (C.c_int64, [vp]),
        "gx_record_bytes": (C.c_int64, [vp]),
        "gx_next_records": (C.c_int, [vp, C.POINTER(u64), u8p, sz, C.POINTER(sz)]),
        "gx_records_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "gx_next_frame": (C.c_int, [vp, C.POINTER(u64), u8p, i32, C.POINTER(i32)]),
        "gx_partition_records": (C.c_int, [vp, i32, vp]),
        "gx_write_sequence_file": (C.c_int, [vp, C.c_char_p, u8p, i32, i32, C.POINTER(u64)]),
        "gx_graph_statistics": (C.c_int, [vp, C.POINTER(GxGraphStats)]),
        "gx_coverage_histogram": (C.c_int, [vp, vp, u64]),
        "gx_coverage_cutoff": (C.c_int, [vp, i32, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "gx_mg_unique_id": (C.c_int, [u8p]),
        "gx_mg_init": (C.c_int, [vp, u8p]),
        "gx_mg_exchange": (C.c_int, [vp]),
        "gx_phase_ms": (C.c_int, [vp, C.POINTER(C.c_float * 8)]),
        "gx_kernel_launches": (u64, [vp]),
        "gx_set_stream": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.gx_abi_version() != GX_ABI_VERSION:
        raise RuntimeError(f"libgenomix_gb ABI {lib.gx_abi_version()} != binding {GX_ABI_VERSION}")
    _lib = lib
    return lib
