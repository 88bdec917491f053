"""The C twin of the oracle (oracle/gx_oracle.c) against the goldens and against oracle.py. CPU only."""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import c_oracle as CO
from oracle import oracle as O


def canon(stream):
    from genomix_b200 import types as T
    return T.canonical_records(stream)


def py_canon(k, text):
    from genomix_b200 import types as T
    return {key: T.Node.read(v, 0)[0].canonical_bytes() for key, v in O.graph_records(k, O.build_graph(k, text)).items()}


@pytest.mark.parametrize("name,k,text,expected", golden_cases(), ids=[c[0] for c in golden_cases()])
@pytest.mark.parametrize("threads", [1, 4])
def test_c_oracle_reproduces_reference_golden(name, k, text, expected, threads):
    from genomix_b200 import types as T
    recs, st = CO.build_graph_records(k, text, threads)
    O.compare_unordered(expected, T.records_to_text(recs))
    assert canon(recs) == py_canon(k, text)


@pytest.mark.parametrize("k", [1, 4, 21, 32, 55, 91, 128])
def test_c_oracle_equals_python_oracle(k):
    rng = np.random.default_rng(k)
    genome = rng.choice(list(b"ACGT"), size=600).astype(np.uint8)
    lines = []
    for i in range(40):
        L = int(rng.integers(k + 1, k + 50))
        s = int(rng.integers(0, 600 - L))
        line = b"%d\t%s" % (4 * i + 2, bytes(genome[s: s + L].tolist()))
        if i % 3 == 0:
            L2 = int(rng.integers(k + 1, k + 50))
            line += b"\t" + bytes(genome[s: s + L2].tolist()[::-1])
        lines.append(line)
    text = b"\n".join(lines) + b"\n"
    want = py_canon(k, text)
    for threads in (1, 3):
        recs, st = CO.build_graph_records(k, text, threads)
        assert canon(recs) == want
        assert st["nodes"] == len(want)


@pytest.mark.parametrize("bad,status", [(b"2\tACGTA\n7\n", -4), (b"x\tACGTA\n", -5), (b"2\tACG\n", -6),
                                        (b"%d\tACGTA\n" % (1 << 29), -7)])
def test_c_oracle_errors(bad, status):
    with pytest.raises(CO.OracleError) as ei:
        CO.build_graph_records(3, bad, 2)
    assert ei.value.status == status


def test_java_partition_matches():
    lib = CO.load()
    rng = np.random.default_rng(0)
    for _ in range(200):
        key = bytes(rng.integers(0, 256, size=int(rng.integers(1, 24)), dtype=np.uint8).tolist())
        assert lib.gxo_java_partition(key, len(key), 13) == O.java_partition(key, 13)


def test_differential_fuzz_python_vs_c_oracle():
    """random small inputs incl. malformed lines, invalid letters, empty fields, CRLF: both restatements must agree on the
    graph, or both must fail with the same error class at the same line"""
    rng = np.random.default_rng(2024)
    alphabet = list(b"ACGTacgtN")
    status_of = {"format": -4, "NumberFormat": -5, "larger than the read": -6, "lose some of its bits": -7}
    n_err = n_ok = 0
    for trial in range(120):
        k = int(rng.integers(1, 9))
        lines = []
        for i in range(int(rng.integers(1, 8))):
            kind = int(rng.integers(0, 12))
            seq = bytes(rng.choice(alphabet[:8] if kind else alphabet, size=int(rng.integers(k + 1, k + 12))).tolist())
            mate = bytes(rng.choice(alphabet[:8], size=int(rng.integers(k + 1, k + 12))).tolist())
            rid = b"%d" % (4 * i + 2)
            if kind == 1:
                line = rid + b"\t" + seq + b"\t" + mate
            elif kind == 2:
                line = rid + b"\t" + seq + b"\t"             # trailing empty field is dropped
            elif kind == 3:
                line = rid + b"\t\t" + mate                  # empty mate 0
            elif kind == 4 and trial % 3 == 0:
                line = rid                                   # too few fields
            elif kind == 5 and trial % 3 == 0:
                line = b"x" + rid + b"\t" + seq              # bad number
            elif kind == 6 and trial % 3 == 0:
                line = rid + b"\t" + seq[:k]                 # k >= length
            elif kind == 7 and trial % 3 == 0:
                line = b"%d\t" % ((1 << 29) + i) + seq       # id loses bits
            else:
                line = rid + b"\t" + seq
            lines.append(line)
        text = (b"\r\n" if trial % 5 == 0 else b"\n").join(lines) + (b"\n" if trial % 2 else b"")
        try:
            want = py_canon(k, text)
            err = None
        except O.GraphBuildError as e:
            want, err = None, str(e)
        if err is None:
            recs, _ = CO.build_graph_records(k, text, 1 + trial % 3)
            assert canon(recs) == want
            n_ok += 1
        else:
            with pytest.raises(CO.OracleError) as ei:
                CO.build_graph_records(k, text, 1)
            expected = [v for key, v in status_of.items() if key in err]
            assert expected and ei.value.status == expected[0], (err, ei.value.status)
            n_err += 1
    assert n_ok > 40 and n_err > 10


def test_line_terminators_python_vs_c_oracle():
    """\\n, \\r and \\r\\n all end a record (hadoop LineReader.readLine); both oracles agree, for every thread count."""
    import numpy as np
    from genomix_b200 import types as T
    rng = np.random.default_rng(11)
    reads = [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(6, 30))).tolist()) for _ in range(120)]
    eols = [b"\n", b"\r", b"\r\n"]
    text = b"".join(b"%d\t%s" % (4 * i + 2, r) + eols[int(rng.integers(0, 3))] for i, r in enumerate(reads))
    want = {key: T.Node.read(val, 0)[0].canonical_bytes() for key, val in O.graph_records(4, O.build_graph(4, text)).items()}
    assert len(O.split_lines(text)) == 120
    for threads in (1, 3, 8):
        got, st = CO.build_graph_records(4, text, threads)
        assert st["lines"] == 120
        assert T.canonical_records(got) == want
