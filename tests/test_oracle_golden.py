"""Pins the oracle (oracle/oracle.py) on every golden vector and known-answer test the reference holds for the
graph-build path. CPU only."""
import struct

import numpy as np
import pytest

from conftest import golden_cases
from oracle import oracle as O


@pytest.mark.parametrize("name,k,text,expected", golden_cases(), ids=[c[0] for c in golden_cases()])
def test_oracle_reproduces_reference_golden(name, k, text, expected):
    """StepByStepTest.TestGroupby / BatchTestFilesTest expected files under TestUtils' comparison rule."""
    table = O.build_graph(k, text)
    O.compare_unordered(expected, O.graph_text_lines(k, table))


def test_nine_goldens_present():
    assert len(golden_cases()) == 9


def test_kmer_fixed_known_answers():
    """KmerFixedTest.java:30-61: AATAGAA -> shift G -> ATAGAAG; reverse of ATAGAAG = CTTCTAT."""
    k = 7
    kb = O.kmer_from_string_bytes(k, b"AATAGAA", 0)
    assert O.recover_kmer(k, kb) == "AATAGAA"
    assert kb.hex() == "0230"
    kb2 = O.kmer_shift_with_next_code(k, kb, O.code_from_symbol(ord("G")))
    assert O.recover_kmer(k, kb2) == "ATAGAAG"
    assert O.recover_kmer(k, O.kmer_reversed_from_string_bytes(k, b"ATAGAAG", 0)) == "CTTCTAT"
    # KmerFixedTest.TestCompressKmer: sliding window over AGCTGACCG for k=3..10 round-trips through toString
    s = b"AGCTGACCGT"
    for kk in range(3, 11):
        for st in range(0, len(s) - kk + 1):
            assert O.recover_kmer(kk, O.kmer_from_string_bytes(kk, s, st)) == s[st: st + kk].decode()


def test_rolling_shift_equals_repack():
    rng = np.random.default_rng(7)
    for k in (3, 4, 5, 8, 21, 31, 32, 33, 55, 64, 91):
        s = bytes(rng.choice(list(b"ACGT"), size=k + 40).tolist())
        kb = O.kmer_from_string_bytes(k, s, 0)
        for i in range(k, len(s)):
            kb = O.kmer_shift_with_next_code(k, kb, O.code_from_symbol(s[i]))
            assert kb == O.kmer_from_string_bytes(k, s, i - k + 1)


def test_vkmer_compare_direction():
    """VKmerFixedTest.java:549-550: CTA < GTA (-1 / +1) pins the byte-compare direction."""
    a = O.kmer_from_string_bytes(3, b"CTA", 0)
    b = O.kmer_from_string_bytes(3, b"GTA", 0)
    assert O.compare_bytes(a, b) < 0 and O.compare_bytes(b, a) > 0


def test_canonical_is_lexicographic_min():
    """SURVEY.md headline fact 5: byte order of the reversed 2-bit layout == string order of the letters read
    from the last one backwards, so pack(s) <= pack(rc s) <=> s <= rc(s) compared as strings."""
    rng = np.random.default_rng(11)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for k in (4, 5, 6, 7, 8, 9, 21, 31, 55, 91):
        for _ in range(200):
            s = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
            rc = bytes(comp[c] for c in reversed(s))
            f = O.kmer_from_string_bytes(k, s, 0)
            r = O.kmer_reversed_from_string_bytes(k, s, 0)
            assert r == O.kmer_from_string_bytes(k, rc, 0)
            assert (O.compare_bytes(f, r) <= 0) == (s <= rc)


def test_read_head_info_bitfields():
    """ReadHeadInfoTest.java:10-55: round trip of (mate, library, readId, offset) incl. negative offsets."""
    for mate, lib, rid, off in [(0, 0, 1, 0), (1, 3, 97, 25), (1, 15, (1 << 29) - 1, (1 << 23) - 1), (0, 7, 12345, -5),
                                (1, 0, 0, -((1 << 23) - 1))]:
        v = O.make_uuid(mate, lib, rid, off)
        assert O.uuid_fields(v) == (mate, lib, rid, off)
    with pytest.raises(ValueError):
        O.make_uuid(0, 0, 1 << 29, 0)
    with pytest.raises(ValueError):
        O.make_uuid(0, 0, -1, 0)
    with pytest.raises(ValueError):
        O.make_uuid(0, 0, 1, 1 << 23)


def test_derived_known_answer_bytes():
    """SURVEY.md §8c derived bytes: packing, Java partition hash, smalltest Node bytes."""
    assert O.kmer_from_string_bytes(3, b"CAG", 0).hex() == "21"
    assert O.kmer_from_string_bytes(21, b"ACGT" * 5 + b"A", 0).hex() == "00e4e4e4e4e4"
    assert O.java_partition(bytes.fromhex("21"), 8) == 0
    assert O.java_partition(bytes.fromhex("0230"), 8) == 7
    assert O.java_partition(bytes.fromhex("00e4e4e4e4e4"), 8) == 5
    recs = O.graph_records(3, O.build_graph(3, b"1\tCAGCCA\tCGTCGA\n"))
    cag = recs[O.vkmer_bytes(3, bytes.fromhex("21"))]
    assert cag.hex() == ("91" "00000001" "00000003" "18" "01" "00000001" "01" "0000000000000001" "00000006" "0161"
                         "00000006" "0279" "3f800000")
    acg = recs[O.vkmer_bytes(3, O.kmer_from_string_bytes(3, b"ACG", 0))]
    assert acg.hex() == ("a8" "00000001" "00000003" "12" "01" "00000001" "01" "0000020800000001" "00000006" "0279"
                         "00000006" "0161" "3f800000")


def test_parse_errors_follow_reference():
    for bad in (b"1\n", b"\n", b"1\tACGT\tACGT\tACGT\n", b"\t\t\n"):
        with pytest.raises(O.GraphBuildError):
            O.build_graph(3, bad)
    with pytest.raises(O.GraphBuildError):
        O.build_graph(3, b"x1\tACGT\n")          # NumberFormatException
    with pytest.raises(O.GraphBuildError):
        O.build_graph(4, b"1\tACGT\n")           # k >= read length
    with pytest.raises(O.GraphBuildError):
        O.build_graph(3, b"%d\tACGT\n" % (1 << 29))  # readId loses bits
    # a read with a non-ACGT letter is skipped whole, not an error; trailing empty fields are dropped
    assert O.build_graph(3, b"1\tACNGT\n") == {}
    assert len(O.build_graph(3, b"1\tACGTA\t\t\n")) > 0
    # invalid mate 0 still becomes mate 1's mateReadSequence, packed with N -> A
    t = O.build_graph(3, b"5\tACNGT\tCCGTA\n")
    heads = [rh for n in t.values() for rh in list(n.unflipped.values()) + list(n.flipped.values())]
    assert len(heads) == 1 and heads[0].to_string().endswith("readSeq: CCGTA mateReadSeq: ACAGT")


def test_java_float_to_string():
    assert O.java_float_to_string(1.0) == "1.0"
    assert O.java_float_to_string(52.0) == "52.0"
    assert O.java_float_to_string(1.0e7) == "1.0E7"
    assert O.java_float_to_string(12345678.0) == "1.2345678E7"


def test_fitting_mixture_cutoff_separates_error_kmers():
    """FittingMixture.fittingMixture restated (FittingMixture.java:92-216): an exponential error peak at low coverage and a
    normal genomic peak; the cut-off is the first coverage at which the normal component outweighs the exponential one."""
    rng = np.random.default_rng(0)
    data = [float(x) for x in np.concatenate([np.round(rng.exponential(3, 5000)) + 1, np.round(rng.normal(50, 10, 5000))]).clip(1)]
    cut, e_mean, n_mean, n_sd = O.fitting_mixture(data, max(data), 10)
    assert 15 <= cut <= 35 and 3 < e_mean < 6 and 48 < n_mean < 53 and 8 < n_sd < 11
    # no second component: every coverage is explained by the exponential -> 0 (the driver then sets no cut-off)
    assert O.fitting_mixture([1.0] * 50, 1.0, 10)[0] == 0
    # permutation invariance (the Hadoop counters arrive in string order of their names)
    shuffled = list(data)
    rng.shuffle(shuffled)
    assert O.fitting_mixture(shuffled, max(data), 10)[0] == cut
