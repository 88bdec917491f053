"""Seeded synthetic inputs for the configs of BASELINE.md §4: random genome, uniform read starts and strands,
i.i.d. substitution errors, emitted as the `id\\tseq[\\tmate]\\n` text ReadsKeyValueParserFactory.parse sees
(read ids are the reference's fastq-conversion ids 4*i+2, GenomixDriver.java:689-690,702-703) and, optionally,
as fastq. Vectorised numpy; no reference code involved.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


@dataclass(frozen=True)
class Workload:
    name: str
    genome_bp: int
    read_len: int
    coverage: float
    error: float
    k: int
    paired: bool = False
    seed: int = 1
    outer_mean: float = 500.0
    outer_std: float = 50.0

    @property
    def n_reads(self) -> int:
        """single reads, or pairs when paired"""
        per = self.read_len * (2 if self.paired else 1)
        return int(np.ceil(self.genome_bp * self.coverage / per))


# BASELINE.md §4 (cfg3-5 also have /100 twins for oracle-sized parity runs)
CONFIGS = {
    "cfg1": Workload("cfg1", 10_000, 100, 50, 0.0, 21, seed=1),
    "cfg2": Workload("cfg2", 4_600_000, 150, 50, 0.01, 31, seed=2),
    "cfg3": Workload("cfg3", 100_000_000, 150, 60, 0.0, 55, paired=True, seed=3),
    "cfg4": Workload("cfg4", 107_000_000, 150, 100, 0.0, 55, seed=4),
    "cfg5": Workload("cfg5", 500_000_000, 250, 20, 0.02, 91, seed=5),
}


# single-GPU sized twins of the multi-GPU configs (same coverage / read length / error / k, smaller genome)
CONFIGS["cfg3s"] = Workload("cfg3s", 6_250_000, 150, 60, 0.0, 55, paired=True, seed=3)
CONFIGS["cfg4s"] = Workload("cfg4s", 6_687_500, 150, 100, 0.0, 55, seed=4)
CONFIGS["cfg5s"] = Workload("cfg5s", 7_812_500, 250, 20, 0.02, 91, seed=5)


def scaled(w: Workload, factor: float, name: str | None = None) -> Workload:
    return Workload(name or f"{w.name}/{factor:g}", max(int(w.genome_bp / factor), w.read_len * 4), w.read_len, w.coverage,
                    w.error, w.k, w.paired, w.seed, w.outer_mean, w.outer_std)


def _sample_reads(rng, genome, starts, read_len, strand, error):
    idx = starts[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
    codes = genome[idx]
    rc = strand.astype(bool)
    codes[rc] = _COMP[codes[rc][:, ::-1]]
    if error > 0:
        err = rng.random(codes.shape) < error
        n_err = int(err.sum())
        if n_err:
            codes[err] = (codes[err] + rng.integers(1, 4, size=n_err, dtype=np.uint8)) & 3
    return codes


def generate_codes(w: Workload, n_reads: int | None = None, batch: int = 1 << 20):
    """Yield (first_record_index, mate0 codes [n, L], mate1 codes or None) batches."""
    rng = np.random.Generator(np.random.PCG64(w.seed))
    genome = rng.integers(0, 4, size=w.genome_bp, dtype=np.uint8)
    total = w.n_reads if n_reads is None else n_reads
    done = 0
    L = w.read_len
    while done < total:
        n = min(batch, total - done)
        if not w.paired:
            starts = rng.integers(0, w.genome_bp - L + 1, size=n)
            strand = rng.integers(0, 2, size=n, dtype=np.uint8)
            yield done, _sample_reads(rng, genome, starts, L, strand, w.error), None
        else:
            outer = np.clip(np.rint(rng.normal(w.outer_mean, w.outer_std, size=n)).astype(np.int64), L, w.genome_bp)
            starts = rng.integers(0, w.genome_bp - outer + 1)
            strand = rng.integers(0, 2, size=n, dtype=np.uint8)
            # fragment [s, s+outer): mate 0 reads its left end forward, mate 1 its right end reverse-complemented;
            # a fragment drawn from the reverse strand swaps the roles
            left = _sample_reads(rng, genome, starts, L, np.zeros(n, np.uint8), w.error)
            right = _sample_reads(rng, genome, starts + outer - L, L, np.ones(n, np.uint8), w.error)
            sw = strand.astype(bool)
            m0 = np.where(sw[:, None], right, left)
            m1 = np.where(sw[:, None], left, right)
            yield done, m0, m1
        done += n


def shard_text(w: Workload, rank: int, n_reads: int, batch: int = 1 << 20, first_record: int | None = None) -> np.ndarray:
    """Readid text of one rank's shard of a multi-GPU workload: the genome comes from w.seed (identical on every rank), the
    rank's `n_reads` reads from an independent stream, so a rank generates only what it will push. Record ids continue
    across ranks (rank r holds records [r*n_reads, (r+1)*n_reads), or [first_record, first_record + n_reads) if given)."""
    genome = np.random.Generator(np.random.PCG64(w.seed)).integers(0, 4, size=w.genome_bp, dtype=np.uint8)
    rng = np.random.Generator(np.random.PCG64([w.seed, 7919 + rank]))
    L = w.read_len
    parts = []
    done = 0
    while done < n_reads:
        n = min(batch, n_reads - done)
        first = (rank * n_reads if first_record is None else first_record) + done
        if not w.paired:
            starts = rng.integers(0, w.genome_bp - L + 1, size=n)
            strand = rng.integers(0, 2, size=n, dtype=np.uint8)
            parts.append(lines_from_codes(first, _sample_reads(rng, genome, starts, L, strand, w.error), None))
        else:
            outer = np.clip(np.rint(rng.normal(w.outer_mean, w.outer_std, size=n)).astype(np.int64), L, w.genome_bp)
            starts = rng.integers(0, w.genome_bp - outer + 1)
            sw = rng.integers(0, 2, size=n, dtype=np.uint8).astype(bool)
            left = _sample_reads(rng, genome, starts, L, np.zeros(n, np.uint8), w.error)
            right = _sample_reads(rng, genome, starts + outer - L, L, np.ones(n, np.uint8), w.error)
            parts.append(lines_from_codes(first, np.where(sw[:, None], right, left), np.where(sw[:, None], left, right)))
        done += n
    if not parts:
        return np.zeros(0, dtype=np.uint8)
    return parts[0] if len(parts) == 1 else np.concatenate(parts)


def _id_digits(ids: np.ndarray):
    """ASCII decimal digits of each id, right-aligned in a [n, 20] matrix, plus digit counts."""
    n = ids.shape[0]
    mat = np.zeros((n, 20), dtype=np.uint8)
    v = ids.astype(np.uint64).copy()
    for col in range(19, -1, -1):
        mat[:, col] = (v % 10).astype(np.uint8) + ord("0")
        v //= 10
    nd = np.maximum(1, np.floor(np.log10(np.maximum(ids, 1).astype(np.float64))).astype(np.int64) + 1)
    # guard against float rounding at powers of ten
    pw = np.power(10.0, nd - 1)
    nd = np.where(ids < pw, nd - 1, nd)
    nd = np.where(ids >= np.power(10.0, nd), nd + 1, nd)
    return mat, np.maximum(nd, 1)


def lines_from_codes(first_record: int, m0: np.ndarray, m1: np.ndarray | None) -> np.ndarray:
    """`<4i+2>\\t<seq>[\\t<mate>]\\n` for each record, as one uint8 array."""
    n, L = m0.shape
    ids = 4 * (np.arange(n, dtype=np.int64) + first_record) + 2
    digits, nd = _id_digits(ids)
    body = L + 1 + (0 if m1 is None else m1.shape[1] + 1)  # \t seq [\t mate] -- the final \n counted with the id
    line_len = nd + 1 + body
    ends = np.cumsum(line_len)
    starts = ends - line_len
    out = np.empty(int(ends[-1]), dtype=np.uint8)
    # ids: write the 20-wide right-aligned digits ending at start+nd, most significant first
    for d in range(1, int(nd.max()) + 1):
        sel = nd >= d
        out[starts[sel] + nd[sel] - d] = digits[sel, 20 - d]
    seq_start = starts + nd + 1
    out[seq_start - 1] = ord("\t")
    out[seq_start[:, None] + np.arange(L)[None, :]] = _ACGT[m0]
    if m1 is None:
        out[seq_start + L] = ord("\n")
    else:
        out[seq_start + L] = ord("\t")
        ms = seq_start + L + 1
        out[ms[:, None] + np.arange(m1.shape[1])[None, :]] = _ACGT[m1]
        out[ms + m1.shape[1]] = ord("\n")
    return out


def readid_text(w: Workload, n_reads: int | None = None) -> np.ndarray:
    """The whole workload as readid text (uint8 array)."""
    parts = [lines_from_codes(first, m0, m1) for first, m0, m1 in generate_codes(w, n_reads)]
    return parts[0] if len(parts) == 1 else np.concatenate(parts)


def fastq_from_codes(first_record: int, codes: np.ndarray) -> bytes:
    """Four-line fastq records (@r<i> / seq / + / IIII...) for one mate."""
    n, L = codes.shape
    seqs = _ACGT[codes]
    qual = b"I" * L
    out = []
    for i in range(n):
        out.append(b"@r%d\n%s\n+\n%s\n" % (first_record + i, seqs[i].tobytes(), qual))
    return b"".join(out)


def occurrences(w: Workload, n_reads: int | None = None) -> int:
    n = w.n_reads if n_reads is None else n_reads
    return n * (2 if w.paired else 1) * (w.read_len - w.k + 1)
