"""CPU-only: the C-ABI library loads and exports every symbol include/genomix_gb.h declares; the product fails
loudly without a GPU; the host-side mirrors round-trip the oracle's bytes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_cases
from oracle import oracle as O


def header_symbols():
    src = open(os.path.join(ROOT, "include", "genomix_gb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from genomix_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/genomix_gb.h but not exported"
    assert sorted(_lib.EXPORTS) == syms
    assert lib.gx_abi_version() == _lib.GX_ABI_VERSION


def test_struct_layouts_match_header():
    from genomix_b200 import _lib
    assert C.sizeof(_lib.GxConfig) == 6 * 4 + 8 + 4 * 8
    assert C.sizeof(_lib.GxStats) == 16 * 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import genomix_b200 as gx
    with pytest.raises(gx.GenomixError) as ei:
        gx.GraphBuilder(21)
    assert ei.value.status == -2 and "no CPU path" in str(ei.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "genomix_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "gx_oracle" not in txt, f


@pytest.mark.parametrize("name,k,text,expected", golden_cases(), ids=[c[0] for c in golden_cases()])
def test_host_types_roundtrip_oracle_bytes(name, k, text, expected):
    """genomix_b200.types decodes the oracle's Node bytes to the same text the oracle prints, and re-encodes them."""
    from genomix_b200 import types as T
    table = O.build_graph(k, text)
    recs = O.graph_records(k, table)
    stream = b"".join(
        (4 + len(key) - 4 + len(val)).to_bytes(4, "big") + len(key).to_bytes(4, "big") + key + val for key, val in recs.items())
    assert sorted(T.records_to_text(stream)) == sorted(O.graph_text_lines(k, table))
    for key, val in T.iter_records(stream):
        node, end = T.Node.read(val, 0)
        assert end == len(val) and node.write() == val
    O.compare_unordered(expected, T.records_to_text(stream))


def test_kmer_string_helpers():
    from genomix_b200 import types as T
    rng = np.random.default_rng(3)
    for k in (1, 3, 4, 5, 21, 31, 32, 55, 91):
        s = "".join(rng.choice(list("ACGT"), size=k))
        assert T.string_to_kmer(s) == O.kmer_from_string_bytes(k, s.encode(), 0)
        assert T.kmer_to_string(k, T.string_to_kmer(s)) == s


def test_header_is_plain_c(tmp_path):
    """the boundary is a C ABI: the header must compile as C (no C++ or torch types in the signatures)"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "genomix_gb.h"\nint main(void) { gx_config c; gx_stats s; gx_graph_stats g; (void)c; (void)s; (void)g; '
                   'return sizeof(c) + sizeof(s) + sizeof(g) == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
