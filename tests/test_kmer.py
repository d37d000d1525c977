"""k-mer rarefaction (--kmer_rarefaction): the reference's QC.Kmercount.txt / QC.kmerH.txt, committed under
tests/golden/kmer/ by tests/golden/make_golden.py, replayed against the CPU oracle (CPU suite) and the CUDA path (-m gpu)."""
import ast
import glob
import os

import numpy as np
import pytest

from faqcs_b200 import synth
from faqcs_b200.api import Engine, FaqcsError, Options
from oracle_binding import OracleEngine
from parity import run_kmer

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "kmer", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]


def load(path):
    z = np.load(path)
    w = eval(bytes(z["paired"]).decode(), {"synth": synth})
    u = eval(bytes(z["unpaired"]).decode(), {"synth": synth})
    opt = Options(**ast.literal_eval(bytes(z["options"]).decode()))
    k, split, subset = (int(x) for x in z["params"])
    return [(w.r1, w.r2), (u.r1, None)], opt, k, split, subset, bytes(z["kmercount"]), bytes(z["kmerh"])


def test_fixtures_exist():
    assert len(FIXTURES) >= 5 and len(CPU_CASES) == 2


# two of the five on the CPU (the oracle's std::unordered_map needs ~6 s per case)
CPU_CASES = [p for p in FIXTURES if os.path.basename(p)[:-4] in ("qc_only_early_stop", "trimmed_replace_to_n")]


@pytest.mark.parametrize("path", CPU_CASES, ids=[os.path.basename(p)[:-4] for p in CPU_CASES])
def test_oracle_reproduces_reference_kmer_files(path):
    passes, opt, k, split, subset, kc, kh = load(path)
    with OracleEngine(opt) as eng:
        got = run_kmer(eng, passes, k, split, subset)
    assert got[0] == kc
    assert got[1] == kh


def naive_kmers(seqs, k):
    """Independent restatement of update_kmer: canonical = min of the 2-bit words of a k-mer and its reverse complement."""
    code = {"A": 0, "T": 1, "C": 2, "G": 3}
    table = {}
    for s in seqs:
        s = s.upper()
        for i in range(len(s) - k + 1):
            win = s[i:i + k]
            if any(c not in code for c in win):
                continue
            f = 0
            for c in win:
                f = (f << 2) | code[c]
            r = 0
            for c in reversed(win):
                r = (r << 2) | (code[c] ^ 1)
            key = min(f, r)
            table[key] = table.get(key, 0) + 1
    return table


def small_workload():
    rng = np.random.default_rng(5)
    seqs = []
    for i in range(300):
        L = int(rng.integers(1, 80))
        s = "".join(rng.choice(list("ACGTacgtN"), p=[.2, .2, .2, .2, .04, .04, .04, .04, .04], size=L))
        seqs.append(s)
    seqs += ["ACGT" * 10, "A" * 40, "acgtn" * 8, "N" * 12, "T"]
    recs = [(f"@r{i}", s, "I" * len(s)) for i, s in enumerate(seqs)]
    return seqs, np.frombuffer(synth.fastq_bytes(recs), dtype=np.uint8)


def check_small(engine_cls, k):
    seqs, fq = small_workload()
    opt = Options(qc_only=True, input_quality_offset=33)
    with engine_cls(opt) as eng:
        eng.kmer_enable(k, 1000000, 20)
        eng.process(fq, None, 0, True)
        eng.kmer_end_pass()
        rare, freq = eng.kmer_results()
    table = naive_kmers(seqs, k)
    want = {}
    for c in table.values():
        want[c] = want.get(c, 0) + 1
    assert {int(c): int(n) for c, n in freq} == want
    assert rare.tolist() == [[len(seqs), len(table), sum(table.values())]]


@pytest.mark.parametrize("k", [2, 3, 7, 16, 31])
def test_oracle_kmers_against_naive_restatement(k):
    check_small(OracleEngine, k)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_reproduces_reference_kmer_files(path):
    passes, opt, k, split, subset, kc, kh = load(path)
    with Engine(opt) as eng:
        got = run_kmer(eng, passes, k, split, subset)
    assert got[0] == kc
    assert got[1] == kh


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 3, 7, 16, 31])
def test_cuda_kmers_against_naive_restatement(k):
    check_small(Engine, k)


@pytest.mark.gpu
@pytest.mark.parametrize("batch_records", [32768, 65536, 10 * 32768])
def test_cuda_kmer_curve_does_not_depend_on_batch_size(batch_records):
    """The points are taken where the reference's 32768-read trim() calls end, whatever the size of the caller's batches."""
    passes, opt, k, split, subset, kc, kh = load(FIXTURES[0])
    with Engine(opt) as eng:
        got = run_kmer(eng, passes, k, split, subset, batch_records=batch_records)
    assert got == (kc, kh)


@pytest.mark.gpu
def test_cuda_kmer_table_grows():
    """More distinct k-mers than the first table holds: the table is rehashed into a larger one between batches."""
    w = synth.c2(3 * 32768)                           # random reads: ~120 distinct 31-mers per read
    opt = Options(qc_only=True)
    with Engine(opt) as eng, OracleEngine(opt) as ora:
        a = run_kmer(eng, [(w.r1, w.r2)], 31, 40000, 4)
        b = run_kmer(ora, [(w.r1, w.r2)], 31, 40000, 4)
    assert a == b


@pytest.mark.gpu
def test_cuda_kmer_rejects_unaligned_batches():
    w = synth.c4(5000)
    with Engine(Options(qc_only=True)) as eng:
        eng.kmer_enable(31, 1000, 2)
        eng.autodetect(w.r1, None)
        with pytest.raises(FaqcsError):
            eng.process(w.r1, None, 100, True)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(8))
def test_cuda_kmer_fuzz_against_oracle(seed):
    """Random option sets (trimming modes, clips, --replace_to_N_q, --qc_only, adapters) over random reads (IUPAC letters, lower
    case, N runs, CRLF): the k-mer curve and histogram must equal the oracle's, whose rule is the reference's -- count what
    trim_read left in the read (trim.cpp:260-262, 389-403, 545-547)."""
    from fuzz import fuzz_bytes, fuzz_options, fuzz_reads
    rng = np.random.default_rng(4000 + seed)
    in_off = 64 if seed % 4 == 3 else 33
    paired = seed % 2 == 0
    eol = "\r\n" if seed % 5 == 4 else "\n"
    r1 = fuzz_bytes(fuzz_reads(rng, 500, in_off, "1" if paired else None), rng, eol)
    r2 = fuzz_bytes(fuzz_reads(rng, 500, in_off, "2"), rng, eol) if paired else None
    kw = fuzz_options(rng, in_off, adapters=seed % 3 == 1)
    if seed % 2:
        kw["replace_to_N_q"] = 20
        kw["qc_only"] = False
    k, split = int(rng.choice([2, 5, 11, 21, 31])), int(rng.choice([100, 400, 100000]))
    res = []
    for cls in (Engine, OracleEngine):
        with cls(Options(**kw)) as e:
            e.kmer_enable(k, split, 4)
            e.process(r1, r2, 0, True)
            e.kmer_end_pass()
            rare, freq = e.kmer_results()
            res.append((rare.tolist(), freq.tolist()))
    assert res[0] == res[1]
