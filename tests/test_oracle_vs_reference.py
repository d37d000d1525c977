"""Pin the CPU oracle against the UNMODIFIED reference binary (oracle/_ref/FaQCs).

These tests are what makes the oracle trustworthy: emitted FASTQ bytes, every
integer in QC.stats.txt and the ten --debug matrix / histogram files must agree.
They run only where oracle/_ref/FaQCs exists (this container and, because the
binary travels with the snapshot, the GPU box)."""
import numpy as np
import pytest

import refcli
from faqcs_b200 import synth
from faqcs_b200.api import MODE_BWA, MODE_HARD, Options
from oracle_binding import OracleEngine
from parity import assert_matches_reference, run_engine

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not built")]


def check(w, opt, threads=1, polyA=False, artifacts=None, extra_flags=(), batch_records=None):
    flags = refcli.flags_for(opt, polyA=polyA) + list(extra_flags)
    if w.r2 is not None:
        ref = refcli.run_reference(w.r1, w.r2, flags=flags, threads=threads, artifacts=artifacts)
    else:
        ref = refcli.run_reference(unpaired=w.r1, flags=flags, threads=threads, artifacts=artifacts)
    opt.adapters = refcli.adapters_for(opt.filter_adapter, polyA, artifacts)
    if artifacts:
        opt.filter_adapter = True
    with OracleEngine(opt) as eng:
        streams, _ = run_engine(eng, w.r1, w.r2, batch_records)
        assert_matches_reference(ref, streams, eng.stats(), opt, opt.adapters)


def test_c2_defaults():
    check(synth.c2(20000), Options())


def test_c2_defaults_multibatch_t4():
    check(synth.c2(40000), Options(), threads=4, batch_records=32768)


def test_c2_discard_5end_3end():
    check(synth.c2(5000), Options(trim_5=7, trim_3=11, discard_output=True))


def test_c2_bwa_and_hard_modes():
    check(synth.c2(5000), Options(mode=MODE_BWA, quality=15, discard_output=True))
    check(synth.c2(5000), Options(mode=MODE_HARD, quality=25, discard_output=True))
    check(synth.c2(5000), Options(quality=20, protect_5=True, average_quality=30.0))


def test_c4_qc_only():
    check(synth.c4(30000), Options(qc_only=True))


def test_c5_mixed_ascii64_hard():
    check(synth.c5(20000), Options(mode=MODE_HARD, quality=20, average_quality=25.0, replace_to_N_q=10,
                                   discard_output=True))


def test_c3_adapters_polya_artifacts_t1():
    w = synth.c3(1500)
    check(w, Options(filter_adapter=True, num_thread=1), threads=1, polyA=True, artifacts=w.artifacts)


def test_c3_adapters_qc_only():
    w = synth.c3(1000)
    check(w, Options(filter_adapter=True, qc_only=True, num_thread=2), threads=2, polyA=True)


@pytest.mark.parametrize("seed", range(8))
def test_fuzz_reads_and_options(seed):
    """Random option sets over records with lower case, IUPAC letters, N runs, Q2 tails, lengths 1..420, '+name'
    third lines and (one seed in five) CRLF line ends: the oracle must reproduce the reference byte for byte."""
    from fuzz import fuzz_bytes, fuzz_options, fuzz_reads
    rng = np.random.default_rng(2000 + seed)
    in_off = 64 if seed % 4 == 3 else 33
    paired = seed % 2 == 0
    eol = "\r\n" if seed % 5 == 4 else "\n"
    r1 = fuzz_bytes(fuzz_reads(rng, 600, in_off, "1" if paired else None), rng, eol)
    r2 = fuzz_bytes(fuzz_reads(rng, 600, in_off, "2"), rng, eol) if paired else None
    kw = fuzz_options(rng, in_off, adapters=seed % 3 == 1)
    polyA = bool(kw.pop("adapters", None))
    threads = kw.get("num_thread", 0) or 2
    check(synth.Workload("fuzz", r1, r2, []), Options(**kw), threads=threads, polyA=polyA)


@pytest.mark.parametrize("qc_only", [True, False], ids=["qc_only", "trimmed"])
def test_kmer_rarefaction_files(qc_only):
    """--kmer_rarefaction: QC.Kmercount.txt / QC.kmerH.txt of a paired pass followed by an unpaired pass, curve cut short
    by --subset (the reference is always at k = 31: -m is an ambiguous abbreviation under its getopt_long_only)."""
    from parity import run_kmer
    w = synth.shotgun(40000, genome_len=30000)
    u = synth.shotgun(35000, genome_len=30000, seed=78, paired=False, L=100)
    opt = Options(qc_only=qc_only)
    flags = refcli.flags_for(opt) + ["--kmer_rarefaction", "--split_size", "25000", "--subset", "2"]
    ref = refcli.run_reference(w.r1, w.r2, unpaired=u.r1, flags=flags, threads=3)
    assert ref["returncode"] == 0, ref["stderr"][:500]
    with OracleEngine(opt) as eng:
        kc, kh = run_kmer(eng, [(w.r1, w.r2), (u.r1, None)], 31, 25000, 2)
    assert kc == ref["files"]["QC.Kmercount.txt"]
    assert kh == ref["files"]["QC.kmerH.txt"]
