"""Micro-cases from SURVEY.md Appendix E / section 9, oracle vs the reference binary."""
import numpy as np
import pytest

import refcli
from faqcs_b200 import synth
from faqcs_b200.api import BUILTIN_ADAPTERS, MODE_BWA, MODE_HARD, Options
from oracle_binding import OracleEngine
from parity import assert_matches_reference, run_engine
from faqcs_b200.synth import Workload, fastq_bytes

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not built")]
AD = dict(BUILTIN_ADAPTERS)


def rnd(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n))


def check_records(recs, opt, threads=1, polyA=False, artifacts=None, paired_recs=None, eol="\n", expect_rc=0):
    r1 = np.frombuffer(fastq_bytes(recs, eol), dtype=np.uint8)
    r2 = np.frombuffer(fastq_bytes(paired_recs, eol), dtype=np.uint8) if paired_recs is not None else None
    flags = refcli.flags_for(opt, polyA=polyA)
    if r2 is not None:
        ref = refcli.run_reference(r1, r2, flags=flags, threads=threads, artifacts=artifacts)
    else:
        ref = refcli.run_reference(unpaired=r1, flags=flags, threads=threads, artifacts=artifacts)
    opt.adapters = refcli.adapters_for(opt.filter_adapter, polyA, artifacts)
    if artifacts:
        opt.filter_adapter = True
    with OracleEngine(opt) as eng:
        streams, res = run_engine(eng, r1, r2)
        assert_matches_reference(ref, streams, eng.stats(), opt, opt.adapters)
    return ref, streams


def test_q1_hard_quirks():
    recs = [("@r0", "A" * 30 + "C" * 30, "#" * 60),                      # all low: kept whole
            ("@r1", "ACGT" * 15, "I" * 30 + "#" * 30),                   # high then low
            ("@r2", "ACGT" * 15, "I" + "#" * 59),                        # only pos 0 high
            ("@r3", "ACGT" * 15, "#" * 20 + "I" * 20 + "#" * 20)]
    ref, streams = check_records(recs, Options(mode=MODE_HARD, quality=10, min_read_length=1, input_quality_offset=33,
                                               low_complexity_cutoff_ratio=1.0))
    lens = [len(l) for l in ref["streams"][2].decode().split("\n")[1::4]]
    assert lens == [60, 30, 60, 20]


def test_q2_find_mask_range_glue():
    rng = np.random.default_rng(1)
    s = rnd(rng, 20) + AD["Nextera-primer-adapter-1"] + rnd(rng, 10) + AD["Nextera-primer-adapter-2"] + rnd(rng, 30)
    recs = [("@q2", s, "I" * (len(s) - 1) + "#")]
    check_records(recs, Options(filter_adapter=True, min_read_length=1, num_thread=1, input_quality_offset=33))


@pytest.mark.parametrize("threads", [1, 2, 4])
def test_q3_thread_dependent_threshold(threads):
    rng = np.random.default_rng(3)
    recs = []
    for i in range(37):
        L = int(rng.integers(26, 90))
        k = int(rng.integers(18, 30))
        s = rnd(rng, L - k) + AD["Nextera-primer-adapter-1"][:k] if L > k else rnd(rng, L)
        recs.append((f"@q3_{i}", s, "I" * (len(s) - 1) + "5"))
    check_records(recs, Options(filter_adapter=True, min_read_length=1, num_thread=threads, input_quality_offset=33),
                  threads=threads)


def test_q3_nine_identical_reads():
    rng = np.random.default_rng(4)
    s = rnd(rng, 5) + AD["Nextera-primer-adapter-1"][:25]
    recs = [(f"@n{i}", s, "I" * 29 + "5") for i in range(9)]
    for t in (1, 2, 4):
        check_records(recs, Options(filter_adapter=True, min_read_length=1, num_thread=t, input_quality_offset=33),
                      threads=t)


def test_q5_stale_range_and_tie_rule():
    rng = np.random.default_rng(5)
    noA = lambda n: rnd(rng, n, "CGT")
    tie = noA(30) + "A" * 20 + noA(30) + "A" * 20 + noA(10)             # Appendix E-17
    polyg = "G" * 80                                                     # shares no base with polyA: stale range
    ct = "CT" * 50
    recs = [("@tie", tie, "I" * (len(tie) - 1) + "5"), ("@pg", polyg, "I" * 79 + "5"), ("@ct", ct, "I" * 99 + "5"),
            ("@mix", noA(40) + AD["Nextera-junction-adapter-1"] + noA(40), "I" * 117 + "5")]
    check_records(recs, Options(filter_adapter=True, min_read_length=1, low_complexity_cutoff_ratio=1.0,
                                num_thread=1, input_quality_offset=33), polyA=True,
                  artifacts=[("polyT", "T" * 25), ("polyC", "C" * 22)])


def test_q6_polya_alone_is_noop():
    rng = np.random.default_rng(6)
    s = rnd(rng, 56) + "A" * 20
    recs = [("@pa", s, "I" * 75 + "5")]
    check_records(recs, Options(min_read_length=1, input_quality_offset=33), polyA=True)


def test_crlf_ascii64_replace_to_n():
    recs = [("@c0 x", "GGACGTACGTNN" * 6, "h" * 30 + "D" * 42), ("@c1", "NNAGGT" * 12, "hE" * 36)]
    check_records(recs, Options(replace_to_N_q=10, min_read_length=10), eol="\r\n")


def test_terminal_n_masking_and_pairs():
    rng = np.random.default_rng(7)
    r1, r2 = [], []
    for i in range(40):
        L = int(rng.integers(40, 120))
        s1 = "NN" + rnd(rng, L) + "TNN" if i % 3 == 0 else rnd(rng, L, "ACGTN" if i % 5 == 0 else "ACGT")
        s2 = rnd(rng, L) if i % 4 else "A" * L
        q1 = "".join(chr(int(x)) for x in rng.integers(35, 74, size=len(s1)))
        q2 = "".join(chr(int(x)) for x in rng.integers(33, 74, size=len(s2)))
        r1.append((f"@p{i}/1", s1, q1))
        r2.append((f"@p{i}/2", s2, q2))
    check_records(r1, Options(discard_output=True, quality=12, input_quality_offset=33), paired_recs=r2)


def test_bwa_plus_short_reads():
    rng = np.random.default_rng(8)
    recs = []
    for i in range(400):
        L = int(rng.integers(1, 12))
        q = "".join(chr(33 + int(x)) for x in rng.choice([2, 2, 2, 8, 20, 30, 40], size=L))
        recs.append((f"@s{i}", rnd(rng, L), q))
    for protect in (False, True):
        check_records(recs, Options(quality=10, min_read_length=1, low_complexity_cutoff_ratio=1.0, protect_5=protect,
                                    input_quality_offset=33, discard_output=True))
    check_records(recs, Options(mode=MODE_BWA, quality=10, min_read_length=1, low_complexity_cutoff_ratio=1.0,
                                input_quality_offset=33, discard_output=True))
    check_records(recs, Options(mode=MODE_HARD, quality=10, min_read_length=1, low_complexity_cutoff_ratio=1.0,
                                input_quality_offset=33, discard_output=True))


def test_clip_longer_than_read_and_min_len():
    rng = np.random.default_rng(9)
    recs = [(f"@k{i}", rnd(rng, L), "I" * (L - 1) + "#") for i, L in enumerate([5, 10, 30, 49, 50, 51, 60, 80, 100])]
    check_records(recs, Options(trim_5=20, trim_3=35, min_read_length=20, input_quality_offset=33, discard_output=True))


def test_low_complexity_edges():
    recs = [("@mono86", "A" * 86 + "CGTCGTCGTCGTCG", "I" * 99 + "5"),
            ("@mono85", "A" * 85 + "CGTCGTCGTCGTCGT", "I" * 99 + "5"),
            ("@di", "AC" * 50, "I" * 99 + "5"), ("@dilow", "ac" * 50, "I" * 99 + "5"),
            ("@dibreak", "AC" * 20 + "N" + "AC" * 29 + "G", "I" * 99 + "5"),
            ("@ok", "ACGT" * 25, "I" * 99 + "5")]
    check_records(recs, Options(input_quality_offset=33, discard_output=True, max_num_poly_N=5))


def test_avgq_boundary():
    recs = [("@e%d" % i, "ACGT" * 25, chr(33 + 30) * (100 - i) + chr(33 + 29) * i) for i in range(0, 6)]
    check_records(recs, Options(average_quality=30.0, quality=0, input_quality_offset=33, discard_output=True))
    check_records(recs, Options(average_quality=29.97, quality=0, input_quality_offset=33, discard_output=True))


def lone_cr_input():
    """Lines that hold a '\\r' which is not part of a CRLF line end: the reference keeps what precedes the FIRST '\\r' of a line
    (strpbrk, fastq.cpp:44,70,100) and drops the rest of the line."""
    rng = np.random.default_rng(77)
    buf = b""
    for i in range(300):
        L = int(rng.integers(60, 140))
        s, q = rnd(rng, L), "".join(chr(int(x)) for x in rng.integers(40, 74, size=L - 1)) + "5"
        k = i % 6
        if k == 0:          # junk behind a CR on the base and quality lines (equal content lengths)
            buf += f"@cr{i}\n{s}\rJUNK\n+\n{q}\rMOREJUNK!\n".encode()
        elif k == 1:        # CR inside the header
            buf += f"@cr{i} left\rright side\n{s}\n+\n{q}\n".encode()
        elif k == 2:        # CRLF record among them
            buf += f"@cr{i}\r\n{s}\r\n+\r\n{q}\r\n".encode()
        elif k == 3:        # CR on the '+' line (content is dropped anyway)
            buf += f"@cr{i}\n{s}\n+x\ry\n{q}\n".encode()
        elif k == 4:        # double CR before the line end
            buf += f"@cr{i}\n{s}\r\r\n+\n{q}\r\r\n".encode()
        else:
            buf += f"@cr{i}\n{s}\n+\n{q}\n".encode()
    return np.frombuffer(buf, dtype=np.uint8)


def test_lone_carriage_returns_cut_the_line():
    r1 = lone_cr_input()
    opt = Options(discard_output=True, min_read_length=100, input_quality_offset=33)
    ref = refcli.run_reference(unpaired=r1, flags=refcli.flags_for(opt), threads=1)
    with OracleEngine(opt) as eng:
        streams, _ = run_engine(eng, r1, None)
        assert_matches_reference(ref, streams, eng.stats(), opt)
    assert b"JUNK" not in ref["streams"][2] and b"right side" not in ref["streams"][2]
