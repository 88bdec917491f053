"""fastq front end: oracle restatement on CPU, gx_push_fastq against it on the GPU."""
import gzip

import numpy as np
import pytest

from oracle import oracle as O


def make_fastq(rng, n, L, messy=False):
    recs = []
    for i in range(n):
        seq = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(L, L + 30))).tolist())
        if messy and i % 5 == 0:
            seq = b"  " + seq + b" \t"
        if messy and i % 7 == 0:
            seq = seq[:3] + b"N" + seq[4:]
        eol = b"\r\n" if messy and i % 3 == 0 else b"\n"
        recs.append(b"@read%d" % i + eol + seq + eol + b"+" + eol + b"I" * len(seq) + eol)
    return b"".join(recs)


def test_oracle_fastq_conversion_ids_and_trim():
    fq = b"@a\n ACGTA \n+\nIIIII\n@b\nGGGTTC\r\n+\nIIIIII\n@c\nTTTT"
    assert O.fastq_to_readids(fq) == b"2\tACGTA\n6\tGGGTTC\n10\tTTTT\n"
    assert O.fastq_to_readids(fq, fq) == b"2\tACGTA\tACGTA\n6\tGGGTTC\tGGGTTC\n10\tTTTT\tTTTT\n"
    with pytest.raises(O.GraphBuildError):
        O.fastq_to_readids(fq, fq + b"\n@d\n")


def test_chunker_covers_every_record(tmp_path):
    from genomix_b200 import fastq
    rng = np.random.default_rng(1)
    fq1, fq2 = make_fastq(rng, 300, 40), make_fastq(rng, 300, 40)
    p1, p2 = tmp_path / "a.fq", tmp_path / "b.fq.gz"
    p1.write_bytes(fq1)
    with gzip.open(p2, "wb") as f:
        f.write(fq2)
    for chunk in (500, 4096, 1 << 20):
        got1, got2, firsts = b"", b"", []
        for r1, r2, first in fastq.iter_chunks(str(p1), str(p2), chunk):
            assert r1.count(b"\n") % 4 == 0 and r1.count(b"\n") == r2.count(b"\n")
            firsts.append((first, r1.count(b"\n") // 4))
            got1 += r1
            got2 += r2
        assert got1 == fq1 and got2 == fq2
        assert [f for f, _ in firsts] == list(np.cumsum([0] + [n for _, n in firsts[:-1]]))
    single = b"".join(r1 for r1, _, _ in fastq.iter_chunks(str(p1), None, 777))
    assert single == fq1


@pytest.mark.timeout(60)
@pytest.mark.parametrize("short_first", [True, False])
def test_chunker_unequal_pair_files_raise_like_the_reference(tmp_path, short_first):
    """GenomixDriver.java:684-687: paired fastq files of different length are an IOException, also when the shorter one
    ends many chunks before the longer one (this used to spin forever)."""
    from genomix_b200 import fastq
    from genomix_b200.graphbuild import GenomixError
    rng = np.random.default_rng(3)
    a, b = tmp_path / "a.fq", tmp_path / "b.fq"
    a.write_bytes(make_fastq(rng, 2, 40))
    b.write_bytes(make_fastq(rng, 200, 40))
    p1, p2 = (a, b) if short_first else (b, a)
    with pytest.raises(GenomixError) as ei:
        for _ in fastq.iter_chunks(str(p1), str(p2), chunk_bytes=256):
            pass
    assert ei.value.status == -4 and "same number of lines" in str(ei.value)   # GX_ERR_FORMAT
    # a line longer than the chunk size is read through, not waited for
    c = tmp_path / "long.fq"
    c.write_bytes(b"@r\n" + b"A" * 1000 + b"\n+\n" + b"I" * 1000 + b"\n")
    assert sum(x[0].count(b"\n") for x in fastq.iter_chunks(str(c), None, chunk_bytes=256)) == 4
    assert sum(x[0].count(b"\n") for x in fastq.iter_chunks(str(c), str(c), chunk_bytes=256)) == 4


@pytest.mark.gpu
@pytest.mark.parametrize("paired", [False, True])
def test_push_fastq_matches_oracle(paired, tmp_path):
    import genomix_b200 as gx
    from genomix_b200 import fastq
    rng = np.random.default_rng(5 + paired)
    k = 21
    fq1 = make_fastq(rng, 400, 30, messy=True)
    fq2 = make_fastq(rng, 400, 30, messy=True) if paired else None
    text = O.fastq_to_readids(fq1, fq2)
    want = {key: gx.types.Node.read(v, 0)[0].canonical_bytes() for key, v in O.graph_records(k, O.build_graph(k, text)).items()}
    with gx.GraphBuilder(k) as gb:
        gb.push_fastq(fq1, fq2)
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == want
    # through files and the chunker (ids continue across chunks)
    p1 = tmp_path / "r1.fq"
    p1.write_bytes(fq1)
    p2 = None
    if paired:
        p2 = tmp_path / "r2.fq.gz"
        with gzip.open(p2, "wb") as f:
            f.write(fq2)
    with gx.GraphBuilder(k) as gb:
        for r1, r2, first in fastq.iter_chunks(str(p1), str(p2) if p2 else None, 5000):
            gb.push_fastq(r1, r2, first)
        gb.finish()
        assert gx.types.canonical_records(gb.records()) == want


@pytest.mark.gpu
def test_push_fastq_errors():
    import genomix_b200 as gx
    fq = b"@a\nACGTACGT\n+\nIIIIIIII\n"
    with pytest.raises(gx.GenomixError) as ei:      # different line counts: the reference's IOException
        with gx.GraphBuilder(3) as gb:
            gb.push_fastq(fq, fq + b"@b\nACGT\n")
    assert ei.value.status == -4
    with pytest.raises(gx.GenomixError) as ei:      # empty sequence -> line "2\t" has one field
        with gx.GraphBuilder(3) as gb:
            gb.push_fastq(b"@a\n\n+\n\n")
            gb.finish()
    assert ei.value.status == -4
    with pytest.raises(gx.GenomixError) as ei:      # k >= read length
        with gx.GraphBuilder(9) as gb:
            gb.push_fastq(fq)
            gb.finish()
    assert ei.value.status == -6
