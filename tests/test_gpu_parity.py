"""Parity of the CUDA path (through the C ABI) with the CPU oracle on seeded inputs:
emitted FASTQ bytes, every statistic, per-read verdicts.  Bit-exact; the only floats
(average quality) are compared bit for bit as well."""
import numpy as np
import pytest

import refcli
from faqcs_b200 import synth
from faqcs_b200.api import (BUILTIN_ADAPTERS, MODE_BWA, MODE_HARD, OFFSET_AUTO, POLYA_ADAPTER, Engine, FaqcsError, Stats,
                            Options)
from faqcs_b200.synth import fastq_bytes
from oracle_binding import OracleEngine
from fuzz import fuzz_bytes as _fuzz_bytes, fuzz_options, fuzz_reads as _fuzz_reads
from parity import assert_engines_equal, assert_matches_reference, run_engine

pytestmark = pytest.mark.gpu
AD = dict(BUILTIN_ADAPTERS)


def both(r1, r2, opt_factory, batch_records=None, check_results=True):
    with OracleEngine(opt_factory()) as ora, Engine(opt_factory()) as gpu:
        ora.set_debug_results(True)
        gpu.set_debug_results(True)
        so, ro = run_engine(ora, r1, r2, batch_records)
        sg, rg = run_engine(gpu, r1, r2, batch_records)
        # avg_q of invalid reads is not defined on the GPU side (never computed)
        for a, b in zip(ro, rg):
            for m in range(2):
                if a.results[m] is not None and b.results[m] is not None:
                    inv = (a.results[m]["flags"] & 1) == 0
                    a.results[m]["avg_q"][inv] = 0
                    b.results[m]["avg_q"][inv] = 0
        assert_engines_equal(sg, gpu.stats(), so, ora.stats(), rg if check_results else None, ro if check_results else None)
        return sg, gpu.stats()


def test_c2_defaults():
    w = synth.c2(30000)
    both(w.r1, w.r2, lambda: Options())


def test_c2_multibatch_equals_single():
    w = synth.c2(20000)
    s1, st1 = both(w.r1, w.r2, lambda: Options(discard_output=True))
    s2, st2 = both(w.r1, w.r2, lambda: Options(discard_output=True), batch_records=3000)
    assert [bytes(x) for x in s1] == [bytes(x) for x in s2]
    assert not st1.diff(st2)


def test_c2_modes_and_clips():
    w = synth.c2(8000)
    both(w.r1, w.r2, lambda: Options(trim_5=7, trim_3=11, discard_output=True))
    both(w.r1, w.r2, lambda: Options(mode=MODE_BWA, quality=15, discard_output=True))
    both(w.r1, w.r2, lambda: Options(mode=MODE_HARD, quality=25, discard_output=True))
    both(w.r1, w.r2, lambda: Options(quality=20, protect_5=True, average_quality=30.0))
    both(w.r1, w.r2, lambda: Options(quality=20, min_read_length=100, max_num_poly_N=1, low_complexity_cutoff_ratio=0.4))


def test_c4_qc_only():
    w = synth.c4(40000)
    both(w.r1, None, lambda: Options(qc_only=True))


def test_c5_mixed_ascii64_hard():
    w = synth.c5(30000)
    both(w.r1, None, lambda: Options(mode=MODE_HARD, quality=20, average_quality=25.0, replace_to_N_q=10,
                                     discard_output=True))


def test_c3_adapters():
    w = synth.c3(3000)
    ads = BUILTIN_ADAPTERS + [POLYA_ADAPTER] + w.artifacts
    both(w.r1, w.r2, lambda: Options(filter_adapter=True, adapters=list(ads)))
    both(w.r1, w.r2, lambda: Options(filter_adapter=True, adapters=list(ads), qc_only=True))


def test_short_reads_thread_emulation():
    rng = np.random.default_rng(3)
    recs = []
    for i in range(2000):
        L = int(rng.integers(26, 120))
        k = int(rng.integers(18, 30))
        s = "".join(rng.choice(list("ACGT"), size=max(L - k, 1))) + AD["Nextera-primer-adapter-1"][:k]
        recs.append((f"@q3_{i}", s, "I" * (len(s) - 1) + "5"))
    r1 = np.frombuffer(fastq_bytes(recs), dtype=np.uint8)
    for t in (1, 2, 4, 7):
        both(r1, None, lambda: Options(filter_adapter=True, adapters=list(BUILTIN_ADAPTERS), min_read_length=1,
                                       num_thread=t, input_quality_offset=33))


def test_micro_cases_vs_oracle():
    rng = np.random.default_rng(11)
    rnd = lambda n, al="ACGT": "".join(rng.choice(list(al), size=n))
    recs = [("@r0", "A" * 30 + "C" * 30, "#" * 60), ("@r1", "ACGT" * 15, "I" * 30 + "#" * 30),
            ("@r2", "ACGT" * 15, "I" + "#" * 59), ("@n0", "NN" + rnd(70) + "TNN", "I" * 75),
            ("@n1", "N" * 64, "I" * 64), ("@n2", "N" * 33 + "ACGT" * 10 + "N", "I" * 74),
            ("@lc0", "A" * 86 + "CGTCGTCGTCGTCG", "I" * 100), ("@lc1", "AC" * 50, "I" * 100),
            ("@lc2", "ac" * 50, "I" * 100), ("@lc3", "AC" * 20 + "N" + "AC" * 29 + "G", "I" * 100),
            ("@nn", rnd(40) + "NN" + rnd(40), "I" * 82), ("@nnn", rnd(31) + "NNN" + rnd(40), "I" * 74),
            ("@cross", rnd(30) + "NNNN" + rnd(40), "I" * 74), ("@one", "A", "I"), ("@two", "AC", "I#")]
    for L in range(1, 40):
        q = "".join(chr(33 + int(x)) for x in rng.choice([2, 2, 2, 8, 20, 30, 40], size=L))
        recs.append((f"@s{L}", rnd(L, "ACGTN"), q))
    r1 = np.frombuffer(fastq_bytes(recs), dtype=np.uint8)
    for mode in (MODE_HARD, MODE_BWA, 2):
        for protect in (False, True):
            both(r1, None, lambda: Options(mode=mode, quality=10, min_read_length=1, protect_5=protect,
                                           input_quality_offset=33, discard_output=True, max_num_poly_N=3))
    both(r1, None, lambda: Options(input_quality_offset=33, discard_output=True, max_num_poly_N=4, min_read_length=30))


def test_crlf_and_ragged_headers():
    recs = [("@c0 x", "GGACGTACGTNA" * 6, "h" * 30 + "D" * 42), ("@c1", "NAAGGT" * 12, "hE" * 36),
            ("@a_very_long_header_" + "z" * 150, "ACGT" * 20, "h" * 80), ("@", "ACGTTGCA" * 9, "g" * 72)]
    r1 = np.frombuffer(fastq_bytes(recs, "\r\n"), dtype=np.uint8)
    both(r1, None, lambda: Options(replace_to_N_q=10, min_read_length=10, discard_output=True))


def test_third_line_variants():
    """fastq.cpp:79-91: the content of the '+' line is never looked at and never written back; the output line is always a bare
    '+'.  One-character third lines that are not '+', named '+' lines and empty third lines must all be re-emitted as "+"."""
    rng = np.random.default_rng(5)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), size=n))
    thirds = ["+", "x", "-", "+name 1", "", "+", "@", "+"]
    buf = b""
    for i in range(400):
        L = int(rng.integers(60, 150))
        q = "".join(chr(int(x)) for x in rng.integers(40, 74, size=L - 1)) + "5"
        buf += f"@t{i}\n{rnd(L)}\n{thirds[i % len(thirds)]}\n{q}\n".encode()
    r1 = np.frombuffer(buf, dtype=np.uint8)
    streams, _ = both(r1, None, lambda: Options(discard_output=True, input_quality_offset=33))
    assert b"\nx\n" not in bytes(streams[2]) and b"+name" not in bytes(streams[2])
    both(r1, None, lambda: Options(discard_output=True, input_quality_offset=33, min_read_length=100))      # raw copies to discard


def test_lone_carriage_returns_cut_the_line():
    """A '\\r' that is not part of a CRLF line end: the content of the line ends there (strpbrk, fastq.cpp:44); pinned against the
    reference binary in tests/test_oracle_micro.py, here CUDA vs oracle in byte mode, pieces mode and several batches."""
    from test_oracle_micro import lone_cr_input
    r1 = lone_cr_input()
    both(r1, None, lambda: Options(discard_output=True, min_read_length=100, input_quality_offset=33))
    both(r1, None, lambda: Options(discard_output=True, min_read_length=100, input_quality_offset=33), batch_records=70)
    with Engine(Options(discard_output=True, min_read_length=100, input_quality_offset=33)) as a, \
            Engine(Options(discard_output=True, min_read_length=100, input_quality_offset=33)) as b:
        b.set_output_pieces(True)
        x, y = a.process(r1), b.process(r1)
        assert [bytes(s_) for s_ in y.expand(r1)] == [bytes(s_) for s_ in x.streams]
    bad = b"@a\nACGT\rXX\n+\nIII5#\n"             # content lengths 4 vs 5 after the cut
    with Engine(Options(input_quality_offset=33)) as e:
        with pytest.raises(FaqcsError) as ei:
            e.process(bad)
        assert "|Sequence| != |Quality|" in str(ei.value)


def test_pairs_routing_with_discard():
    rng = np.random.default_rng(7)
    rnd = lambda n, al="ACGT": "".join(rng.choice(list(al), size=n))
    r1, r2 = [], []
    for i in range(3000):
        L = int(rng.integers(30, 130))
        s1 = "NN" + rnd(L) + "TNN" if i % 3 == 0 else rnd(L, "ACGTN" if i % 5 == 0 else "ACGT")
        s2 = rnd(L) if i % 4 else "A" * L
        q1 = "".join(chr(int(x)) for x in rng.integers(35, 74, size=len(s1)))
        q2 = "".join(chr(int(x)) for x in rng.integers(33, 74, size=len(s2)))
        r1.append((f"@p{i}/1", s1, q1))
        r2.append((f"@p{i}/2", s2, q2))
    a = np.frombuffer(fastq_bytes(r1), dtype=np.uint8)
    b = np.frombuffer(fastq_bytes(r2), dtype=np.uint8)
    both(a, b, lambda: Options(discard_output=True, quality=12, input_quality_offset=33))
    both(a, b, lambda: Options(discard_output=True, quality=12, input_quality_offset=33), batch_records=700)


def test_anticorrelated_mates_fill_the_unpaired_stream():
    """Pairs alternate between (long valid R1, short invalid R2) and the reverse: the unpaired stream receives the LONG mate of
    every pair and the discard stream the short one, so the unpaired output approaches the larger input, not max(n1, n2) / 2
    (the unpaired device buffer must be sized for sum_i max(rec1_i, rec2_i))."""
    rng = np.random.default_rng(21)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), size=n))
    r1, r2 = [], []
    for i in range(4000):
        long_s, short_s = rnd(int(rng.integers(200, 300))), rnd(int(rng.integers(5, 20)))
        a, b = (long_s, short_s) if i % 2 == 0 else (short_s, long_s)
        r1.append((f"@ac{i}/1", a, "I" * (len(a) - 1) + "5"))
        r2.append((f"@ac{i}/2", b, "I" * (len(b) - 1) + "5"))
    a = np.frombuffer(fastq_bytes(r1), dtype=np.uint8)
    b = np.frombuffer(fastq_bytes(r2), dtype=np.uint8)
    streams, _ = both(a, b, lambda: Options(discard_output=True, input_quality_offset=33))
    assert len(streams[2]) > 0.9 * max(a.size, b.size) and len(streams[0]) == 0
    both(a, b, lambda: Options(discard_output=True, input_quality_offset=33), batch_records=900)


@pytest.mark.parametrize("case", ["c2", "c2_single", "c5", "crlf", "routing", "thirds"])
def test_pieces_mode_expands_to_the_byte_streams(case):
    """fq_set_output_pieces: every stream comes back as pieces of the caller's input plus literal bytes; expanding them must give
    exactly the bytes byte mode emits (and the oracle), with the same statistics.  Most C2 records travel as pieces."""
    rng = np.random.default_rng(31)
    rnd = lambda n, al="ACGT": "".join(rng.choice(list(al), size=n))
    r2 = None
    kw = dict(discard_output=True)
    if case == "c2":
        w = synth.c2(20000); r1, r2 = w.r1, w.r2
    elif case == "c2_single":
        w = synth.c2(20000); r1 = w.r1; kw = dict(quality=20)
    elif case == "c5":
        w = synth.c5(15000); r1 = w.r1
        kw = dict(mode=MODE_HARD, quality=20, average_quality=25.0, replace_to_N_q=10, discard_output=True)
    elif case == "crlf":
        recs = [(f"@c{i} x", rnd(int(rng.integers(60, 140))), None) for i in range(500)]
        recs = [(h, s, "h" * (len(s) - 1) + "D") for h, s, _ in recs]
        r1 = np.frombuffer(fastq_bytes(recs, "\r\n"), dtype=np.uint8); kw = dict(min_read_length=100, discard_output=True)
    elif case == "routing":
        a, b = [], []
        for i in range(3000):
            L = int(rng.integers(30, 130))
            a.append((f"@p{i}/1", "NN" + rnd(L) + "TNN" if i % 3 == 0 else rnd(L, "ACGTN" if i % 5 == 0 else "ACGT"), None))
            b.append((f"@p{i}/2", rnd(L) if i % 4 else "A" * L, None))
        q = lambda s_: "".join(chr(int(x)) for x in rng.integers(35, 74, size=len(s_)))
        r1 = np.frombuffer(fastq_bytes([(h, s_, q(s_)) for h, s_, _ in a]), dtype=np.uint8)
        r2 = np.frombuffer(fastq_bytes([(h, s_, q(s_)) for h, s_, _ in b]), dtype=np.uint8)
        kw = dict(discard_output=True, quality=12, input_quality_offset=33)
    else:
        thirds = ["+", "x", "+name 1", "", "+"]
        buf = b""
        for i in range(600):
            L = int(rng.integers(60, 150))
            buf += f"@t{i}\n{rnd(L)}\n{thirds[i % 5]}\n{'I' * (L - 1)}5\n".encode()
        r1 = np.frombuffer(buf, dtype=np.uint8); kw = dict(discard_output=True, input_quality_offset=33, min_read_length=100)
    with Engine(Options(**kw)) as plain, Engine(Options(**kw)) as pc, OracleEngine(Options(**kw)) as ora:
        pc.set_output_pieces(True)
        for e in (plain, pc, ora):
            e.autodetect(r1, r2)
        a = plain.process(r1, r2)
        b = pc.process(r1, r2)
        c = ora.process(r1, r2)
        assert b.pieces is not None and b.stream_bytes == a.stream_bytes
        got = b.expand(r1, r2)
        assert [bytes(x) for x in got] == [bytes(x) for x in a.streams] == [bytes(x) for x in c.streams]
        assert not plain.stats().diff(pc.stats()) and not pc.stats().diff(ora.stats())
        if case == "c2":        # almost everything is a piece of the input: literal bytes are a small fraction
            assert sum(len(x) for x in b.streams) < 0.2 * sum(a.stream_bytes)
            assert sum(len(p) for p in b.pieces) < 0.4 * 2 * b.n_records
        # several batches: pieces refer to the buffers of THEIR batch
        cuts1 = [0] + [int(x) for x in np.flatnonzero(r1 == 10)[3::4][999::1000] + 1]
        if cuts1[-1] != r1.size:
            cuts1.append(int(r1.size))
        if r2 is not None:
            cuts2 = [0] + [int(x) for x in np.flatnonzero(r2 == 10)[3::4][999::1000] + 1]
            if cuts2[-1] != r2.size:
                cuts2.append(int(r2.size))
        acc = [b"", b"", b"", b""]
        with Engine(Options(**kw)) as many:
            many.set_output_pieces(True)
            many.autodetect(r1, r2)
            for k in range(len(cuts1) - 1):
                x1 = r1[cuts1[k]:cuts1[k + 1]]
                x2 = r2[cuts2[k]:cuts2[k + 1]] if r2 is not None else None
                res = many.process(x1, x2, 1000 * k, k == len(cuts1) - 2)
                for s_, part in enumerate(res.expand(x1, x2)):
                    acc[s_] += part
        assert [bytes(x) for x in acc] == [bytes(x) for x in a.streams]


def test_pipelined_submit_run_wait_equals_synchronous():
    from faqcs_b200 import shard
    w = synth.c2(12000)
    opt = lambda: Options(discard_output=True, quality=12)
    with Engine(opt()) as a, Engine(opt()) as b:
        a.autodetect(w.r1, w.r2)
        b.autodetect(w.r1, w.r2)
        sync = a.process(w.r1, w.r2)
        b1, b2 = shard.record_batches(w.r1, 2500), shard.record_batches(w.r2, 2500)
        n = len(b1)
        streams = [b"", b"", b"", b""]
        tk = [None] * n
        tk[0] = b.submit(w.r1[b1[0][0]:b1[0][1]], w.r2[b2[0][0]:b2[0][1]], 0, n == 1)
        done = []
        for i in range(n):
            if i + 1 < n:
                tk[i + 1] = b.submit(w.r1[b1[i + 1][0]:b1[i + 1][1]], w.r2[b2[i + 1][0]:b2[i + 1][1]], (i + 1) * 2500, i + 2 == n)
            b.run(tk[i])
            if i > 0:
                done.append(b.wait(tk[i - 1]))
        done.append(b.wait(tk[n - 1]))
        for res in done:
            for k in range(4):
                streams[k] += res.streams[k]
        assert [bytes(x) for x in streams] == [bytes(x) for x in sync.streams]
        assert not a.stats().diff(b.stats())
        assert sum(r.n_records for r in done) == sync.n_records


def test_long_reads_beyond_shared_rows():
    rng = np.random.default_rng(12)
    recs = []
    for i in range(60):
        L = int(rng.integers(500, 2500))
        s = "".join(rng.choice(list("ACGTN"), p=[.24, .25, .25, .25, .01], size=L))
        q = "".join(chr(33 + int(x)) for x in np.clip(rng.normal(25, 10, size=L), 0, 41).astype(int))
        recs.append((f"@long{i}", s, q))
    r1 = np.frombuffer(fastq_bytes(recs), dtype=np.uint8)
    both(r1, None, lambda: Options(quality=15, max_num_poly_N=3, discard_output=True))


def test_errors_match_reference_text():
    ok = fastq_bytes([("@a", "ACGT" * 20, "I" * 79 + "#")])
    with Engine(Options(input_quality_offset=33)) as e:              # Q > 41 (with autodetect 'K' would mean ASCII-64)
        bad = fastq_bytes([("@a", "ACGT" * 20, "K" * 79 + "#")])
        with pytest.raises(FaqcsError) as ei:
            e.process(bad)
        assert "greater than the maximum allowed quality score" in str(ei.value)
    with Engine(Options()) as e:                                     # Q10: nothing decisive
        with pytest.raises(FaqcsError) as ei:
            e.autodetect(fastq_bytes([("@a", "ACGT" * 20, "I" * 80)]))
        assert "Unknown quality format!" in str(ei.value)
    with Engine(Options()) as e:                                     # |seq| != |qual|
        with pytest.raises(FaqcsError) as ei:
            e.autodetect(ok + b"@b\nACGT\n+\nII#\n")
        assert "|Sequence| != |Quality|" in str(ei.value)
    with Engine(Options(input_quality_offset=33)) as e:              # truncated record
        with pytest.raises(FaqcsError) as ei:
            e.process(ok + b"@b\nACGT\n")
        assert "Unable to read '+'" in str(ei.value)
    with Engine(Options(input_quality_offset=33)) as e:              # pair id mismatch
        with pytest.raises(FaqcsError) as ei:
            e.process(ok, fastq_bytes([("@zzz", "ACGT" * 20, "I" * 79 + "#")]))
        assert "FaQCs.cpp:trim: I/O error" in str(ei.value)
    with Engine(Options()) as e:                                     # NextSeq bumps -q to 20
        off, q = e.autodetect(fastq_bytes([("@NS500:1", "ACGT" * 20, "I" * 79 + "#")]))
        assert (off, q) == (33, 20)


@pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not present")
def test_c2_against_reference_binary_directly():
    w = synth.c2(20000)
    opt = Options(discard_output=True)
    ref = refcli.run_reference(w.r1, w.r2, flags=refcli.flags_for(opt), threads=4)
    with Engine(opt) as gpu:
        streams, _ = run_engine(gpu, w.r1, w.r2, batch_records=7000)
        assert_matches_reference(ref, streams, gpu.stats(), opt)


@pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not present")
def test_c5_and_c3_against_reference_binary_directly():
    w = synth.c5(15000)
    opt = Options(mode=MODE_HARD, quality=20, average_quality=25.0, replace_to_N_q=10, discard_output=True)
    ref = refcli.run_reference(unpaired=w.r1, flags=refcli.flags_for(opt), threads=2)
    with Engine(opt) as gpu:
        streams, _ = run_engine(gpu, w.r1, None)
        assert_matches_reference(ref, streams, gpu.stats(), opt)
    w = synth.c3(1500)
    opt = Options(filter_adapter=True, adapters=refcli.adapters_for(True, True, w.artifacts))
    ref = refcli.run_reference(w.r1, w.r2, flags=refcli.flags_for(opt, polyA=True), threads=3, artifacts=w.artifacts)
    with Engine(opt) as gpu:
        streams, _ = run_engine(gpu, w.r1, w.r2)
        assert_matches_reference(ref, streams, gpu.stats(), opt, opt.adapters)


@pytest.mark.parametrize("seed", range(12))
def test_randomized_options_and_reads(seed):
    rng = np.random.default_rng(1000 + seed)
    in_off = 64 if seed % 4 == 3 else 33
    paired = seed % 2 == 0
    n = 700
    eol = "\r\n" if seed % 5 == 4 else "\n"
    r1 = _fuzz_bytes(_fuzz_reads(rng, n, in_off, "1" if paired else None), rng, eol)
    r2 = _fuzz_bytes(_fuzz_reads(rng, n, in_off, "2"), rng, eol) if paired else None
    kw = fuzz_options(rng, in_off, adapters=seed % 3 == 1)
    # (thread-count emulation, SURVEY Q3, needs 32768-record batch boundaries: adapter runs stay single-batch)
    batch = None if kw.get("filter_adapter") else (int(rng.choice([0, 257])) or None)
    both(r1, r2, lambda: Options(**kw), batch_records=batch)


@pytest.mark.parametrize("seed", range(8))
def test_adapter_sweep_fuzz(seed):
    """The segment sweep is a filter in front of the exact alignment: whatever it does, results must equal the oracle's
    (which aligns every adapter against every read): adapter sets and reads built to reach every branch of it."""
    from fuzz import adapter_fuzz_case
    recs, kw = adapter_fuzz_case(seed)
    r1 = np.frombuffer(fastq_bytes(recs), dtype=np.uint8)
    both(r1, None, lambda: Options(**kw))


def test_full_batch_size_properties():
    """BASELINE-size batch (2 M pairs = 4 M reads in one call, the bench's step): properties that do not need the oracle.
    Linearity: a batch made of 8 copies of a 250 k-pair block must give exactly 8 x the block's statistics and 8 x its
    output bytes (checksum of checksums); conservation: every base is either kept or accounted to exactly one filter /
    trimming counter, and the position matrices add up to the base counters."""
    import hashlib
    w = synth.c2(250_000)
    reps = 8
    with Engine(Options(discard_output=True)) as one, Engine(Options(discard_output=True)) as many:
        one.autodetect(w.r1, w.r2)
        many.autodetect(w.r1, w.r2)
        a = one.process(w.r1, w.r2)
        b = many.process(np.tile(w.r1, reps), np.tile(w.r2, reps))
        sa, sb = one.stats(), many.stats()
    assert b.n_records == reps * a.n_records
    for s in range(4):
        assert len(b.streams[s]) == reps * len(a.streams[s])
        block = hashlib.sha256(a.streams[s]).digest()
        n = len(a.streams[s])
        for k in range(reps):            # every copy emits the same bytes at k x the block's offset
            assert hashlib.sha256(b.streams[s][k * n:(k + 1) * n]).digest() == block, (s, k)
    for f in Stats.FIELDS:
        x, y = getattr(sa, f), getattr(sb, f)
        assert x.shape == y.shape and np.array_equal(reps * x.astype(np.int64), y.astype(np.int64)), f
    fs = sb.filter_stats.astype(np.int64)
    T = {n: i for i, n in enumerate(["TOTAL_COUNT", "TOTAL_NUMBER", "TOTAL_LENGTH", "TRIMMED_NUMBER", "TRIMMED_LENGTH", "PAIRED_NUMBER",
                                    "PAIRED_LENGTH", "READ_LENGTH", "BASE_LENGTH", "READ_NN", "BASE_NN", "READ_PHIX", "BASE_PHIX",
                                    "READ_ADAPTER", "BASE_ADAPTER", "READ_AVG_Q", "BASE_AVG_Q", "READ_QUAL_TRIM", "BASE_QUAL_TRIM",
                                    "READ_LC", "BASE_LC"])}
    assert fs[T["TOTAL_NUMBER"]] == 2 * b.n_records and fs[T["TOTAL_LENGTH"]] == 150 * fs[T["TOTAL_NUMBER"]]
    # reads: kept + one counter per discarded read; bases: kept + quality-trimmed + the bases of discarded reads
    assert fs[T["TRIMMED_NUMBER"]] + fs[T["READ_LENGTH"]] + fs[T["READ_NN"]] + fs[T["READ_AVG_Q"]] + fs[T["READ_LC"]] == fs[T["TOTAL_NUMBER"]]
    assert (fs[T["TRIMMED_LENGTH"]] + fs[T["BASE_QUAL_TRIM"]] + fs[T["BASE_LENGTH"]] + fs[T["BASE_NN"]] + fs[T["BASE_AVG_Q"]]
            + fs[T["BASE_LC"]] == fs[T["TOTAL_LENGTH"]])
    assert int(sb.pre_quality_matrix.sum()) == fs[T["TOTAL_LENGTH"]] == int(sb.pre_base_matrix.sum())
    assert int(sb.post_quality_matrix.sum()) == fs[T["TRIMMED_LENGTH"]] == int(sb.post_base_matrix.sum())
    assert int(sb.pre_length_hist.sum()) == fs[T["TOTAL_NUMBER"]] and int(sb.post_length_hist.sum()) == fs[T["TRIMMED_NUMBER"]]
    assert int(sb.pre_read_quality_hist.sum()) == fs[T["TOTAL_NUMBER"]] and int(sb.post_base_quality_hist.sum()) == fs[T["TRIMMED_LENGTH"]]
    assert sum(b.n_valid) == fs[T["TRIMMED_NUMBER"]] and b.paired_read_number == fs[T["PAIRED_NUMBER"]]


def test_quality_change_between_batches():
    """fq_set_quality (the drivers' NextSeq adjustment, FaQCs.cpp:272-277): the second batch is trimmed at Q20."""
    w1, w2 = synth.c2(3000), synth.c2(2000, start=50_000)
    outs = []
    for cls in (OracleEngine, Engine):
        with cls(Options(input_quality_offset=33, discard_output=True)) as eng:
            a = eng.process(w1.r1, w1.r2, 0, False)
            eng.set_quality(20)
            b = eng.process(w2.r1, w2.r2, a.n_records, True)
            outs.append(([a.streams, b.streams], eng.stats()))
    (so, sto), (sg, stg) = outs
    assert so == sg
    assert not stg.diff(sto), stg.diff(sto)
    with Engine(Options(input_quality_offset=33, discard_output=True)) as eng:       # and it does change the outcome
        eng.process(w1.r1, w1.r2, 0, False)
        c = eng.process(w2.r1, w2.r2, 3000, True)
    assert c.streams != sg[1]


@pytest.mark.gpu
def test_two_contexts_on_one_device_overlap_batches_and_merge():
    """Two contexts on the same device, each driven by its own host thread (one batch in flight per context, so that one
    context's k_trim overlaps the other's framing / emit): streams in batch order and the merged statistics
    (fq_merge_stats: dst += src, src = 0) equal one context over the whole input, reads of 600 bases on one context only."""
    import threading
    from faqcs_b200 import shard
    w = synth.c2(24000)
    long_reads = synth.fastq_bytes([(f"@long{i}", "ACGT" * 150, "I" * 599 + "5") for i in range(64)])
    r1 = np.concatenate([np.asarray(w.r1), np.frombuffer(long_reads, dtype=np.uint8)])
    r2 = np.concatenate([np.asarray(w.r2), np.frombuffer(long_reads, dtype=np.uint8)])
    okw = dict(discard_output=True, quality=15)
    batch_records = 3000
    b1, b2 = shard.record_batches(r1, batch_records), shard.record_batches(r2, batch_records)
    with Engine(Options(**okw)) as one:
        one.autodetect(r1, r2)
        single = one.process(r1, r2)
        single_stats = one.stats()
    engines = [Engine(Options(**okw)), Engine(Options(**okw))]
    try:
        for e in engines:
            e.autodetect(r1, r2)
        results = [None] * len(b1)

        def work(j):
            for k in range(j, len(b1), 2):
                results[k] = engines[j].process(r1[b1[k][0]:b1[k][1]], r2[b2[k][0]:b2[k][1]], k * batch_records, k == len(b1) - 1)

        th = [threading.Thread(target=work, args=(j,)) for j in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for i in range(4):
            assert b"".join(bytes(r.streams[i]) for r in results) == bytes(single.streams[i])
        engines[0].merge_stats_from(engines[1])
        assert not engines[0].stats().diff(single_stats), engines[0].stats().diff(single_stats)
        assert int(engines[1].stats().filter_stats.sum()) == 0
    finally:
        for e in engines:
            e.close()
