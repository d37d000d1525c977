"""Seeded synthetic Illumina-like FASTQ for the BASELINE.json workloads
(labels C2-C5).  Pure numpy, vectorised, so that a
million-pair block is generated in seconds.

Generator constraints that the reference would otherwise abort on
(BASELINE.md section 3): Q <= 41 (fastq.h:15,31-33); a quality char < 59
(ASCII-33 sets) or > 74 (ASCII-64 set) within the first 32768 reads
(trim.cpp:599-617); headers do not start with ``@NS`` (trim.cpp:619-626);
R1/R2 ids equal up to the first space (trim.cpp:188-222); LF line endings,
bare ``+`` line, final newline, no blank lines (fastq.cpp:34-66).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .api import BUILTIN_ADAPTERS

SEEDS = {"C2": 20261017, "C3": 20261018, "C4": 20261019, "C5": 5}
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class Workload:
    """A named workload: reads plus the FaQCs flags it is meant to be run with."""
    name: str
    r1: np.ndarray                      # uint8 FASTQ bytes
    r2: Optional[np.ndarray]            # None for single-end
    cli_flags: List[str]                # reference command-line flags (without -1/-2/-u/-d)
    artifacts: Optional[List[Tuple[str, str]]] = None   # extra adapters (--artifactFile contents)


def _headers(n: int, tag: str, mate: int, start: int = 0) -> np.ndarray:
    """Fixed-width 34-byte headers '@SYN:<tag>:1:TTTT:XXXXX:YYYYY <mate>:N:0:1'."""
    idx = np.arange(start, start + n, dtype=np.int64)
    tile = (idx // 100000) % 10000
    x = (idx * 7919 + 13) % 100000
    y = idx % 100000
    prefix = f"@SYN:{tag}:1:".encode()
    suffix = f" {mate}:N:0:1".encode()
    w = len(prefix) + 4 + 1 + 5 + 1 + 5 + len(suffix)
    h = np.empty((n, w), dtype=np.uint8)
    c = 0
    h[:, c:c + len(prefix)] = np.frombuffer(prefix, dtype=np.uint8); c += len(prefix)

    def digits(v, k):
        out = np.empty((n, k), dtype=np.uint8)
        for d in range(k):
            out[:, k - 1 - d] = 48 + (v // (10 ** d)) % 10
        return out

    h[:, c:c + 4] = digits(tile, 4); c += 4
    h[:, c] = ord(":"); c += 1
    h[:, c:c + 5] = digits(x, 5); c += 5
    h[:, c] = ord(":"); c += 1
    h[:, c:c + 5] = digits(y, 5); c += 5
    h[:, c:c + len(suffix)] = np.frombuffer(suffix, dtype=np.uint8)
    return h


def _qualities(rng, n, L, base_lo, base_hi, slope_hi, sigma, qmin, qmax, tail_frac, tail_max):
    base = rng.integers(base_lo, base_hi + 1, size=(n, 1)).astype(np.float32)
    slope = rng.uniform(0.0, slope_hi, size=(n, 1)).astype(np.float32)
    pos = np.arange(L, dtype=np.float32)[None, :]
    q = base - slope * pos + rng.normal(0.0, sigma, size=(n, L)).astype(np.float32)
    q = np.clip(np.rint(q), qmin, qmax).astype(np.uint8)
    if tail_frac > 0:
        has_tail = rng.random(n) < tail_frac
        tail_len = rng.integers(1, tail_max + 1, size=n)
        tail_len = np.where(has_tail, tail_len, 0)
        cut = (L - tail_len)[:, None]
        q = np.where(np.arange(L)[None, :] >= cut, np.uint8(2), q)
    return q


def _bases(rng, n, L, p_n, lowcomp_frac):
    s = _BASES[rng.integers(0, 4, size=(n, L))]
    if p_n > 0:
        s = np.where(rng.random((n, L)) < p_n, np.uint8(ord("N")), s)
    if lowcomp_frac > 0:
        lc = np.flatnonzero(rng.random(n) < lowcomp_frac)
        kind = rng.integers(0, 3, size=lc.size)
        for i, k in zip(lc, kind):
            if k == 0:                                    # poly-G (NextSeq style)
                s[i, :] = ord("G")
            elif k == 1:                                  # dinucleotide repeat
                a, b = rng.choice(4, size=2, replace=False)
                s[i, 0::2] = _BASES[a]
                s[i, 1::2] = _BASES[b]
            else:                                         # 90 % mono + noise
                s[i, :] = np.where(rng.random(L) < 0.9, _BASES[rng.integers(0, 4)], s[i, :])
    return s


def _assemble(hdr: np.ndarray, seq: np.ndarray, qual_chars: np.ndarray, lengths: Optional[np.ndarray] = None) -> np.ndarray:
    """Lay records out as header LF seq LF + LF qual LF; ragged lengths via a keep-mask."""
    n, hw = hdr.shape
    L = seq.shape[1]
    w = hw + 1 + L + 1 + 2 + L + 1
    rec = np.empty((n, w), dtype=np.uint8)
    c = 0
    rec[:, c:c + hw] = hdr; c += hw
    rec[:, c] = 10; c += 1
    s0 = c
    rec[:, c:c + L] = seq; c += L
    rec[:, c] = 10; c += 1
    rec[:, c] = ord("+"); c += 1
    rec[:, c] = 10; c += 1
    q0 = c
    rec[:, c:c + L] = qual_chars; c += L
    rec[:, c] = 10
    if lengths is None:
        return rec.reshape(-1)
    keep = np.ones((n, w), dtype=bool)
    col = np.arange(L)[None, :]
    drop = col >= lengths[:, None]
    keep[:, s0:s0 + L] = ~drop
    keep[:, q0:q0 + L] = ~drop
    return rec[keep]


def _mutate(rng, s: np.ndarray, k: int) -> np.ndarray:
    s = s.copy()
    if k and s.size:
        for p in rng.integers(0, s.size, size=k):
            s[p] = _BASES[rng.integers(0, 4)]
    return s


def make_primers(seed: int = SEEDS["C3"], n: int = 64) -> List[Tuple[str, str]]:
    rng = np.random.default_rng(seed ^ 0x5EED)
    out = []
    for i in range(n):
        L = int(rng.integers(20, 31))
        out.append((f"P{i:02d}", _BASES[rng.integers(0, 4, size=L)].tobytes().decode()))
    return out


def _inject_adapters(rng, s: np.ndarray, primers: List[Tuple[str, str]], mate: int):
    n, L = s.shape
    nextera = np.frombuffer(dict(BUILTIN_ADAPTERS)[f"Nextera-primer-adapter-{mate}"].encode(), dtype=np.uint8)
    r = rng.random(n)
    short = np.flatnonzero(r < 0.05)
    for i in short:                      # short insert: 3' end runs into the Nextera adapter
        k = int(rng.integers(20, L - 10))
        ad = _mutate(rng, nextera, int(rng.integers(0, 4)))
        m = min(L - k, ad.size)
        s[i, k:k + m] = ad[:m]
    polya = np.flatnonzero((r >= 0.05) & (r < 0.06))
    for i in polya:
        k = int(rng.integers(18, 41))
        s[i, L - k:] = ord("A")
    prim = np.flatnonzero((r >= 0.06) & (r < 0.08))
    for i in prim:
        p = np.frombuffer(primers[int(rng.integers(0, len(primers)))][1].encode(), dtype=np.uint8)
        s[i, :p.size] = p


def c2(n_pairs: int, seed: int = SEEDS["C2"], start: int = 0, L: int = 150) -> Workload:
    """configs[1]: 2x150 PE, ASCII-33, default trim + filters + full stats (other L: the same recipe at another read length)."""
    rng = np.random.default_rng(seed + start)
    mates = []
    for mate in (1, 2):
        q = _qualities(rng, n_pairs, L, 30, 40, 0.12, 3.0, 2, 41, 0.10, 59)
        s = _bases(rng, n_pairs, L, 0.002, 0.01)
        mates.append(_assemble(_headers(n_pairs, "C2", mate, start), s, q + 33))
    return Workload("C2", mates[0], mates[1], [])


def c3(n_pairs: int, seed: int = SEEDS["C3"], start: int = 0) -> Workload:
    """configs[2]: C2-style reads + adapters / poly-A / 64 primers (seq_overlap path)."""
    rng = np.random.default_rng(seed + start)
    L = 150
    primers = make_primers(seed)
    mates = []
    for mate in (1, 2):
        q = _qualities(rng, n_pairs, L, 30, 40, 0.12, 3.0, 2, 41, 0.10, 59)
        s = _bases(rng, n_pairs, L, 0.002, 0.01)
        _inject_adapters(rng, s, primers, mate)
        mates.append(_assemble(_headers(n_pairs, "C3", mate, start), s, q + 33))
    return Workload("C3", mates[0], mates[1], ["--adapter", "--rate", "0.2", "--polyA"], artifacts=primers)


def c4(n_reads: int, seed: int = SEEDS["C4"], start: int = 0) -> Workload:
    """configs[3]: --qc_only statistics pass over 1x100 SE reads."""
    rng = np.random.default_rng(seed + start)
    L = 100
    q = _qualities(rng, n_reads, L, 30, 40, 0.12, 3.0, 2, 41, 0.0, 0)
    s = _bases(rng, n_reads, L, 0.002, 0.01)
    return Workload("C4", _assemble(_headers(n_reads, "C4", 1, start), s, q + 33), None, ["--qc_only"])


def c5(n_reads: int, seed: int = SEEDS["C5"], start: int = 0) -> Workload:
    """configs[4]: mixed 50-300 bp SE, ASCII-64 autodetect, HARD -q 20 --avg_q 25 --replace_to_N_q 10 --discard."""
    rng = np.random.default_rng(seed + start)
    Lmax = 300
    lengths = rng.integers(50, Lmax + 1, size=n_reads)
    q = _qualities(rng, n_reads, Lmax, 22, 40, 0.08, 4.0, 0, 41, 0.0, 0)
    s = _bases(rng, n_reads, Lmax, 0.003, 0.01)
    if n_reads:
        q[0, 0] = 40          # 'h' (104) > 74 right away so autodetect says 64
    return Workload("C5", _assemble(_headers(n_reads, "C5", 1, start), s, q + 64, lengths), None,
                    ["--mode", "HARD", "-q", "20", "--avg_q", "25", "--replace_to_N_q", "10", "--discard"])


def shotgun(n_pairs: int, genome_len: int = 60000, seed: int = 77, start: int = 0, paired: bool = True, L: int = 150) -> Workload:
    """Reads drawn from a small random genome (both strands, 0.5 % substitutions, a few N): the workload of
    --kmer_rarefaction, where k-mers must repeat for the curve and the frequency histogram to mean anything."""
    g = _BASES[np.random.default_rng(seed).integers(0, 4, size=genome_len + 2 * L)]
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"ACGTN")] = list(b"TGCAN")
    rng = np.random.default_rng(seed + 1 + start)
    mates = []
    pos = rng.integers(0, genome_len, size=n_pairs)
    for mate in (1, 2) if paired else (1,):
        off = pos if mate == 1 else (pos + rng.integers(0, L, size=n_pairs)) % genome_len
        s = g[off[:, None] + np.arange(L)[None, :]]
        rev = rng.random(n_pairs) < 0.5
        s = np.where(rev[:, None], comp[s][:, ::-1], s)
        err = rng.random((n_pairs, L)) < 0.005
        s = np.where(err, _BASES[rng.integers(0, 4, size=(n_pairs, L))], s)
        s = np.where(rng.random((n_pairs, L)) < 0.001, np.uint8(ord("N")), s)
        lower = rng.random(n_pairs) < 0.01                 # soft-masked reads: update_kmer folds case (trim.cpp:903-918)
        s = np.where(lower[:, None], s | 0x20, s)
        q = _qualities(rng, n_pairs, L, 30, 40, 0.12, 3.0, 2, 41, 0.10, 59)
        mates.append(_assemble(_headers(n_pairs, "SG", mate, start), s, q + 33))
    return Workload("shotgun", mates[0], mates[1] if paired else None, ["--kmer_rarefaction"])


def fastq_bytes(records: List[Tuple[str, str, str]], eol: str = "\n") -> bytes:
    """Hand-written records -> FASTQ bytes (micro-cases)."""
    return "".join(f"{h}{eol}{s}{eol}+{eol}{q}{eol}" for h, s, q in records).encode()


def algorithmic_bytes(r1: np.ndarray, r2: Optional[np.ndarray], out_bytes: List[int]) -> int:
    """B_in + B_out, summed over the batch."""
    return int(r1.size + (0 if r2 is None else r2.size) + sum(out_bytes))
