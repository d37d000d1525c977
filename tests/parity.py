"""Shared parity helpers: drive an Engine (oracle or CUDA) over a workload and
compare with the reference binary's files or with another Engine."""
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from faqcs_b200.api import STAT, Engine, Options, Stats

MATRIX_FIELDS = ("pre_quality_matrix", "post_quality_matrix", "pre_base_matrix", "post_base_matrix",
                 "pre_read_quality_hist", "pre_base_quality_hist", "post_read_quality_hist",
                 "post_base_quality_hist", "pre_composition", "post_composition",
                 "pre_length_hist", "post_length_hist")


def run_engine(engine: Engine, r1, r2=None, batch_records: Optional[int] = None):
    """Autodetect on the first batch, then process (optionally in several batches of whole records)."""
    r1 = np.asarray(r1, dtype=np.uint8) if not isinstance(r1, (bytes, bytearray)) else np.frombuffer(r1, np.uint8)
    if r2 is not None:
        r2 = np.asarray(r2, dtype=np.uint8) if not isinstance(r2, (bytes, bytearray)) else np.frombuffer(r2, np.uint8)
    engine.autodetect(r1, r2)
    if batch_records is None:
        res = engine.process(r1, r2, 0, True)
        return [res.streams[i] for i in range(4)], [res]
    cuts1 = record_cuts(r1, batch_records)
    cuts2 = record_cuts(r2, batch_records) if r2 is not None else None
    streams = [b"", b"", b"", b""]
    results = []
    first = 0
    for k in range(len(cuts1) - 1):
        a1 = r1[cuts1[k]:cuts1[k + 1]]
        a2 = r2[cuts2[k]:cuts2[k + 1]] if r2 is not None else None
        res = engine.process(a1, a2, first, k == len(cuts1) - 2)
        first += res.n_records
        for i in range(4):
            streams[i] += res.streams[i]
        results.append(res)
    return streams, results


def record_cuts(buf: np.ndarray, batch_records: int) -> List[int]:
    nl = np.flatnonzero(buf == 10)
    assert nl.size % 4 == 0
    ends = nl[3::4] + 1
    cuts = [0] + [int(ends[i]) for i in range(batch_records - 1, ends.size, batch_records)]
    if cuts[-1] != buf.size:
        cuts.append(int(buf.size))
    if len(cuts) == 1:
        cuts.append(0)
    return cuts


_STAT_PATTERNS_NORMAL = [
    ("TOTAL_NUMBER", r"Before Trimming\nReads #: (\d+)"),
    ("TOTAL_LENGTH", r"Before Trimming\nReads #: \d+\nTotal bases: (\d+)"),
    ("TOTAL_TRIMMED_NUMBER", r"After Trimming\nReads #: (\d+)"),
    ("TOTAL_TRIMMED_LENGTH", r"After Trimming\nReads #: \d+ \(.*\)\nTotal bases: (\d+)"),
    ("PAIRED_READ_NUMBER", r"  Paired Reads #: (\d+)"),
    ("PAIRED_BASE_LENGTH", r"  Paired total bases: (\d+)"),
    ("READ_LENGTH", r"  Reads Filtered by length cutoff \(\d+ bp\): (\d+)"),
    ("BASE_LENGTH", r"  Bases Filtered by length cutoff: (\d+)"),
    ("READ_NN", r"  Reads Filtered by continuous base \"N\" \(\d+\): (\d+)"),
    ("BASE_NN", r"  Bases Filtered by continuous base \"N\": (\d+)"),
    ("READ_LOW_COMPLEXITY", r"  Reads Filtered by low complexity ratio \([\d.]+\): (\d+)"),
    ("BASE_LOW_COMPLEXITY", r"  Bases Filtered by low complexity ratio: (\d+)"),
    ("READ_AVG_Q", r"  Reads Filtered by avg quality \([\d.]+\): (\d+)"),
    ("BASE_AVG_Q", r"  Bases Filtered by avg quality: (\d+)"),
    ("READ_QUAL_TRIM", r"  Reads Trimmed by quality \([\d.]+\): (\d+)"),
    ("BASE_QUAL_TRIM", r"  Bases Trimmed by quality: (\d+)"),
    ("READ_ADAPTER", r"  Reads Trimmed with Adapters/Primers: (\d+)"),
    ("BASE_ADAPTER", r"  Bases Trimmed with Adapters/Primers: (\d+)"),
]
_STAT_PATTERNS_QC = [
    ("TOTAL_COUNT", r"\nReads #: (\d+)"),
    ("TOTAL_LENGTH", r"Total bases: (\d+)"),
    ("TOTAL_NUMBER", r"Processed (\d+) reads"),
    ("READ_LENGTH", r"  Reads length < \d+ bp: (\d+)"),
    ("READ_NN", r"  Reads have \d+ continuous base \"N\": (\d+)"),
    ("READ_LOW_COMPLEXITY", r"  Low complexity Reads .*: (\d+) \("),
    ("READ_AVG_Q", r"  Reads < average quality [\d.]+: (\d+)"),
    ("READ_ADAPTER", r"  Reads with Adapters/Primers: (\d+)"),
]


def parse_stats_txt(txt: str, qc_only: bool) -> Dict[str, int]:
    out = {}
    for name, pat in (_STAT_PATTERNS_QC if qc_only else _STAT_PATTERNS_NORMAL):
        m = re.search(pat, txt)
        if m:
            out[name] = int(m.group(1))
    return out


def parse_adapter_lines(txt: str) -> Dict[str, Tuple[int, int]]:
    out = {}
    for m in re.finditer(r"^    (\S+) (\d+) reads \(.*?\) (\d+) bases", txt, flags=re.M):
        out[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return out


def assert_matches_reference(ref: Dict, streams: Sequence[bytes], stats: Stats, opt: Options,
                             adapters: Sequence[Tuple[str, str]] = (), check_streams: bool = True):
    assert ref["returncode"] == 0, ref["stderr"]
    if check_streams and not opt.qc_only and "stream_hashes" in ref:
        import hashlib
        for i, name in enumerate(("R1", "R2", "unpaired", "discard")):
            a = bytes(streams[i])
            n, digest = ref["stream_hashes"][i]
            assert len(a) == n, f"stream {name}: {len(a)} bytes, reference wrote {n}"
            assert hashlib.sha256(a).digest() == digest, f"stream {name}: sha256 differs from the reference's file"
    elif check_streams and not opt.qc_only:
        for i, name in enumerate(("R1", "R2", "unpaired", "discard")):
            a, b = bytes(streams[i]), ref["streams"][i]
            if a != b:
                k = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
                raise AssertionError(f"stream {name} differs at byte {k} (len {len(a)} vs ref {len(b)}): "
                                     f"{a[max(0,k-60):k+60]!r} vs {b[max(0,k-60):k+60]!r}")
    parsed = parse_stats_txt(ref["stats_txt"], opt.qc_only)
    for name, val in parsed.items():
        got = int(stats.filter_stats[STAT[name]])
        assert got == val, f"{name}: {got} != reference {val}"
    if adapters:
        want = parse_adapter_lines(ref["stats_txt"])
        got: Dict[str, Tuple[int, int]] = {}
        for (name, _), r, b in zip(adapters, stats.adapter_reads, stats.adapter_bases):
            if r or b or name in got:
                pr, pb = got.get(name, (0, 0))
                got[name] = (pr + int(r), pb + int(b))
        got = {k: v for k, v in got.items() if v[0] or v[1]}
        assert got == want, f"adapter stats differ: {got} != {want}"
    if "pre_quality_matrix" in ref:
        for f in MATRIX_FIELDS:
            a, b = getattr(stats, f), ref[f]
            if f.endswith("length_hist") and a.size <= 1 and b.size <= 1:
                continue          # the file format cannot tell an empty histogram from one with only bin 0
            assert a.shape == b.shape, f"{f}: shape {a.shape} != reference {b.shape}"
            if not np.array_equal(a, b):
                idx = np.argwhere(a != b)[0]
                raise AssertionError(f"{f}: differs at {tuple(idx)}: {a[tuple(idx)]} != reference {b[tuple(idx)]}")


def assert_engines_equal(a_streams, a_stats: Stats, b_streams, b_stats: Stats, a_results=None, b_results=None):
    for i, name in enumerate(("R1", "R2", "unpaired", "discard")):
        x, y = bytes(a_streams[i]), bytes(b_streams[i])
        if x != y:
            k = next((j for j in range(min(len(x), len(y))) if x[j] != y[j]), min(len(x), len(y)))
            raise AssertionError(f"stream {name} differs at byte {k} (len {len(x)} vs {len(y)}): "
                                 f"{x[max(0,k-80):k+80]!r} vs {y[max(0,k-80):k+80]!r}")
    d = a_stats.diff(b_stats)
    assert not d, "stats differ: " + "; ".join(d)
    if a_results is not None and b_results is not None:
        for ra, rb in zip(a_results, b_results):
            for m in range(2):
                if ra.results[m] is None or rb.results[m] is None:
                    continue
                x, y = ra.results[m], rb.results[m]
                assert x.shape == y.shape
                for f in ("offset_5", "length", "flags", "adapter"):
                    if not np.array_equal(x[f], y[f]):
                        k = int(np.flatnonzero(x[f] != y[f])[0])
                        raise AssertionError(f"read result {f} differs at mate {m} read {k}: {x[k]} vs {y[k]}")
                assert np.array_equal(x["avg_q"].view(np.uint32), y["avg_q"].view(np.uint32)), "avg_q bits differ"


def kmer_files(rare: np.ndarray, freq: np.ndarray) -> Tuple[bytes, bytes]:
    """PlotInfo::kmer_rarefaction / kmer_frequency_histogram as plot.cpp:683-733 prints them
    (prefix.Kmercount.txt, prefix.kmerH.txt)."""
    last = 0
    kc = []
    for num_seq, distinct, total in rare.tolist():
        kc.append(f"{num_seq - last}\t{distinct}\t{total}\n")
        last = num_seq
    kh = [f"{c} {n}\n" for c, n in freq.tolist()]
    return "".join(kc).encode(), "".join(kh).encode()


def run_kmer(engine: Engine, passes, k: int, split_size: int, subset: int, batch_records: int = 32768):
    """Drive the k-mer rarefaction API over `passes` = [(r1, r2 or None), ...] the way FaQCs.cpp does (the paired pass,
    then the unpaired pass); --subset is doubled as options.cpp:506-523 does for every accepted command line."""
    engine.kmer_enable(k, split_size, 2 * subset)
    first = True
    for r1, r2 in passes:
        r1 = np.frombuffer(bytes(r1), np.uint8)
        r2 = np.frombuffer(bytes(r2), np.uint8) if r2 is not None else None
        if first:
            engine.autodetect(r1, r2)
            first = False
        cuts1 = record_cuts(r1, batch_records)
        cuts2 = record_cuts(r2, batch_records) if r2 is not None else None
        done = 0
        for j in range(len(cuts1) - 1):
            a1 = r1[cuts1[j]:cuts1[j + 1]]
            a2 = r2[cuts2[j]:cuts2[j + 1]] if r2 is not None else None
            res = engine.process(a1, a2, done, j == len(cuts1) - 2)
            done += res.n_records
        engine.kmer_end_pass()
    return kmer_files(*engine.kmer_results())
