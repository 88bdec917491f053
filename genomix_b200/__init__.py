"""genomix_b200 -- B200-native (sm_100a) graph build for Genomix, behind the reference's operator surface.

The compute lives in libgenomix_gb.so (genomix_b200/csrc, C ABI in include/genomix_gb.h). This package is the
host-side mirror of the reference's interfaces for that one path; it has no CPU fallback.
"""
from .graphbuild import GenomixError, GraphBuilder, build_graph  # noqa: F401
from . import types, synth  # noqa: F401

__all__ = ["GraphBuilder", "GenomixError", "build_graph", "types", "synth"]
