"""Drop-in check at the command line: the same flags through the reference binary (oracle/_ref/FaQCs)
and through faqcs_b200/host/faqcs_b200 must leave byte-identical files: the four trimmed FASTQ files,
QC.stats.txt and the ten --debug matrix / histogram files."""
import gzip
import os
import shutil
import signal
import subprocess
import tempfile

import numpy as np
import pytest

import refcli
from faqcs_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "faqcs_b200", "host", "faqcs_b200")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refcli.have_ref(), reason="reference binary not present"),
              pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built")]


def bgzf_bytes(data: bytes, block: int = 0xff00, level: int = 1, eof_marker: bool = True) -> bytes:
    """Blocked gzip as bgzip writes it (SAM specification 4.1): members of <= 64 KiB with their size in a 'BC' extra field."""
    import struct
    import zlib
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + ([b""] if eof_marker else [])
    for c in chunks:
        z = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = z.compress(c) + z.flush()
        out += struct.pack("<4BI2BH2BHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, 66, 67, 2, len(payload) + 25)
        out += payload + struct.pack("<II", zlib.crc32(c), len(c))
    return bytes(out)


def run_both(inputs, flags, threads=2, extra_cli=(), env=None):
    tmp = tempfile.mkdtemp(prefix="faqcs_cli_")
    try:
        args = []
        for flag, (name, data) in inputs.items():
            path = os.path.join(tmp, name)
            if name.endswith(".bgz.gz"):
                open(path, "wb").write(bgzf_bytes(bytes(data)))
            elif name.endswith(".mixed.gz"):       # blocked members first, ordinary gzip members appended
                d = bytes(data)
                cut = d.index(b"\n", len(d) // 2) + 1
                open(path, "wb").write(bgzf_bytes(d[:cut], eof_marker=False) + gzip.compress(d[cut:], 1))
            elif name.endswith(".gz"):
                with gzip.open(path, "wb", compresslevel=1) as fh:
                    fh.write(bytes(data))
            else:
                open(path, "wb").write(bytes(data))
            args += [flag, path]
        outs = {}
        for tag, exe, more in (("ref", refcli.REF_BIN, []), ("gpu", CLI, list(extra_cli))):
            out = os.path.join(tmp, tag)
            p = subprocess.run([exe, "-d", out, "-t", str(threads), "--debug"] + args + list(flags) + more, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, preexec_fn=lambda: signal.signal(signal.SIGPIPE, signal.SIG_IGN),
                               env=dict(os.environ, **(env or {})))
            assert p.returncode == 0, (tag, p.stderr.decode(errors="replace")[-600:])
            outs[tag] = {n: open(os.path.join(out, n), "rb").read() for n in sorted(os.listdir(out)) if not n.endswith(".pdf")}
        return outs
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def assert_same_files(outs):
    ref, gpu = outs["ref"], outs["gpu"]
    assert sorted(ref) == sorted(gpu), (sorted(ref), sorted(gpu))
    for name in ref:
        if ref[name] != gpu[name]:
            a, b = gpu[name], ref[name]
            k = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
            raise AssertionError(f"{name} differs at byte {k}: {a[max(0,k-80):k+80]!r} vs reference {b[max(0,k-80):k+80]!r}")


def test_paired_defaults_discard_small_batches():
    w = synth.c2(40000)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--discard"], threads=4, extra_cli=["--batch_mb", "3"])
    assert_same_files(outs)
    assert len(outs["ref"]) == 15          # 4 fastq + stats + 10 data files


def test_paired_mates_of_different_length_many_batches():
    """Mate 2 records are shorter than mate 1 records, so every batch cut leaves tails of different sizes in the
    two reader buffers (host pipeline); gzip for one mate, plain for the other."""
    w = synth.c2(30000)
    rec = bytes(w.r2).split(b"\n")
    short = bytearray()
    for i in range(0, len(rec) - 1, 4):
        short += rec[i] + b"\n" + rec[i + 1][:90] + b"\n+\n" + rec[i + 3][:90] + b"\n"
    outs = run_both({"-1": ("r1.fq.gz", w.r1), "-2": ("r2.fq", np.frombuffer(bytes(short), dtype=np.uint8))}, ["--discard"], threads=3,
                    extra_cli=["--batch_mb", "1"])
    assert_same_files(outs)


def test_parallel_io_paths_small_input():
    """The multi-threaded pread reader and the mapped multi-threaded writer normally engage above 8 MiB per transfer;
    here they are forced on for a small input (many batches, ragged slices) and must leave the same files."""
    w = synth.c2(30000)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--discard", "-q", "20"], threads=2,
                    extra_cli=["--batch_mb", "2"], env={"FAQCS_B200_IO_SLICE_MIN": "1000"})
    assert_same_files(outs)


def test_nextseq_recheck_on_the_final_partial_batch():
    """FaQCs.cpp:272-277 / 613-618: at the end of the input the reference tests the first header of its final partial
    32768-read batch for "@NS" again and, if it matches, trims that batch with -q 20.  Record 32768 of this input starts
    with "@NS" while the file does not: only the last 432 pairs are trimmed at Q20."""
    n = 33200
    w = synth.c2(n)
    def mark(buf):
        b = bytearray(bytes(buf))
        rec = len(b) // n                       # fixed-width synthetic records
        assert b[32768 * rec:32768 * rec + 4] == b"@SYN"
        b[32768 * rec + 1:32768 * rec + 3] = b"NS"
        return np.frombuffer(bytes(b), dtype=np.uint8)
    r1, r2 = mark(w.r1), mark(w.r2)
    outs = run_both({"-1": ("r1.fq", r1), "-2": ("r2.fq", r2)}, ["--discard"], threads=2)
    assert_same_files(outs)
    assert b"-q 20" in outs["ref"]["QC.stats.txt"] or b"20" in outs["ref"]["QC.stats.txt"]
    # the same through the single-end driver, small batches: the cut at record 32768 falls inside the last device batch
    assert_same_files(run_both({"-u": ("u.fq", r1)}, [], threads=2, extra_cli=["--batch_mb", "8"]))


def test_unpaired_gz_ascii64_hard():
    w = synth.c5(20000)
    outs = run_both({"-u": ("u.fq.gz", w.r1)}, ["--mode", "HARD", "-q", "20", "--avg_q", "25", "--replace_to_N_q", "10", "--discard"])
    assert_same_files(outs)


def test_qc_only_single_end():
    w = synth.c4(30000)
    assert_same_files(run_both({"-u": ("u.fq", w.r1)}, ["--qc_only"]))


def test_adapters_polya_artifacts():
    w = synth.c3(2000)
    fa = "".join(f">{n}\n{s}\n" for n, s in w.artifacts).encode()
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2), "--artifactFile": ("primers.fa", fa)},
                    ["--adapter", "--polyA", "--rate", "0.2", "--5end", "3", "--min_L", "40"], threads=3)
    assert_same_files(outs)


def test_paired_then_unpaired_truncation_quirk():
    w = synth.c2(3000)
    u = synth.c4(2000)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2), "-u": ("u.fq", u.r1)}, ["--discard", "-q", "20"])
    assert_same_files(outs)            # Q11: the -u pass truncates QC.unpaired / QC.discard written by the paired pass


def test_empty_input_prints_nan_layout():
    # an empty file with --ascii 33 yields "Reads Length: -nan" etc.; both binaries must agree byte for byte
    outs = run_both({"-u": ("empty.fq", b"")}, ["--ascii", "33"])
    assert_same_files(outs)
    assert b"-nan" in outs["ref"]["QC.stats.txt"]


def _run(exe, args, tmp, tag):
    out = os.path.join(tmp, tag)
    return subprocess.run([exe, "-d", out, "-t", "2"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          preexec_fn=lambda: signal.signal(signal.SIGPIPE, signal.SIG_IGN))


def test_error_exits_match():
    """Uneven pair files, a Q42 character and an undetectable offset: both binaries must fail (non-zero exit)."""
    w = synth.c2(200)
    tmp = tempfile.mkdtemp(prefix="faqcs_cli_err_")
    try:
        def put(name, data):
            path = os.path.join(tmp, name)
            open(path, "wb").write(bytes(data))
            return path
        r1, r2 = put("r1.fq", w.r1), put("r2.fq", w.r2)
        cut = put("r2_short.fq", bytes(w.r2)[: w.r2.size // 2 - (w.r2.size // 2) % 339])
        q42 = put("q42.fq", b"@a\n" + b"ACGT" * 20 + b"\n+\n" + b"K" * 79 + b"#\n")
        allI = put("allI.fq", b"@a\n" + b"ACGT" * 20 + b"\n+\n" + b"I" * 80 + b"\n")
        cases = [(["-1", r1, "-2", cut], "uneven"), (["-u", q42, "--ascii", "33"], "q42"), (["-u", allI], "offset")]
        for args, tag in cases:
            ref = _run(refcli.REF_BIN, args, tmp, "ref_" + tag)
            gpu = _run(CLI, args, tmp, "gpu_" + tag)
            assert ref.returncode != 0, (tag, "reference unexpectedly succeeded")
            assert gpu.returncode != 0, (tag, gpu.stderr.decode()[-300:])
        assert b"Unknown quality format!" in _run(CLI, ["-u", allI], tmp, "gpu_offset2").stderr
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_phix_filter_with_thread_emulation():
    """--phiX adds two 5386-nt targets; full SIMD groups and tail groups use different thresholds,
    so this only matches because the -t dependent grouping is reproduced."""
    rng = np.random.default_rng(44)
    phix = open(os.path.join(ROOT, "faqcs_b200", "host", "phix174.inc")).read()
    phix = "".join(l.strip().strip('"') for l in phix.splitlines() if l.startswith('"'))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    recs = []
    for i in range(1203):
        L = int(rng.integers(60, 151))
        kind = i % 4
        if kind == 0:                                   # phiX read, forward strand, a few substitutions
            p0 = int(rng.integers(0, len(phix) - L))
            s = list(phix[p0:p0 + L])
            for k in rng.integers(0, L, size=int(rng.integers(0, 6))):
                s[k] = "ACGT"[int(rng.integers(0, 4))]
            s = "".join(s)
        elif kind == 1:                                 # reverse complement strand
            p0 = int(rng.integers(0, len(phix) - L))
            s = "".join(comp[c] for c in reversed(phix[p0:p0 + L]))
        else:
            s = "".join(rng.choice(list("ACGT"), size=L))
        q = "".join(chr(33 + int(x)) for x in np.clip(rng.normal(32, 6, size=L), 2, 41).astype(int))
        recs.append((f"@px{i}", s, q))
    data = synth.fastq_bytes(recs)
    for t in (1, 3):
        assert_same_files(run_both({"-u": ("u.fq", data)}, ["--phiX", "--discard", "--min_L", "30"], threads=t))


def test_substitute_flag_is_a_noop_with_its_stats_line():
    """--substitute is parsed, warned about and never applied; QC.stats.txt still prints the N_TO_* line."""
    w = synth.c2(3000)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--substitute", "--discard"])
    assert_same_files(outs)
    plain = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--discard"])
    assert outs["ref"]["QC.1.trimmed.fastq"] == plain["ref"]["QC.1.trimmed.fastq"]
    assert outs["ref"]["QC.stats.txt"] != plain["ref"]["QC.stats.txt"]


def test_polya_alone_does_nothing():
    """--polyA without --adapter / --artifactFile / --phiX adds the poly-A target but the adapter pass never runs."""
    w = synth.c3(1500)
    alone = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--polyA", "--min_L", "30"])
    assert_same_files(alone)
    both_flags = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--polyA", "--adapter", "--min_L", "30"])
    assert_same_files(both_flags)
    assert alone["ref"]["QC.1.trimmed.fastq"] != both_flags["ref"]["QC.1.trimmed.fastq"]


def test_c1_example_reads_through_the_cli():
    """BASELINE configs[0]: the example reads (committed inside tests/golden/c1_example_*.npz), defaults, -t 2."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "c1_example_paired.npz"))
    assert_same_files(run_both({"-1": ("r1.fq", z["r1"]), "-2": ("r2.fq", z["r2"])}, [], threads=2))
    u = np.load(os.path.join(ROOT, "tests", "golden", "c1_example_unpaired.npz"))
    assert_same_files(run_both({"-u": ("u.fq", u["r1"])}, ["--discard"], threads=2))
    assert_same_files(run_both({"-1": ("r1.fq", z["r1"]), "-2": ("r2.fq", z["r2"]), "-u": ("u.fq", u["r1"])}, [], threads=2,
                               extra_cli=["--batch_mb", "1"]))


@pytest.mark.parametrize("flags", [["--qc_only", "--split_size", "30000", "--subset", "2"],
                                   ["--split_size", "10000", "--subset", "3", "-q", "20"]], ids=["qc_only", "trimmed"])
def test_kmer_rarefaction_files(flags):
    """--kmer_rarefaction: QC.Kmercount.txt and QC.kmerH.txt next to the other --debug files, paired pass
    then unpaired pass, batches cut at multiples of 32768 records."""
    w = synth.shotgun(70000)
    u = synth.shotgun(40000, seed=78, paired=False, L=100)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2), "-u": ("u.fq", u.r1)}, ["--kmer_rarefaction"] + flags, threads=3,
                    extra_cli=["--batch_mb", "16"])
    assert "QC.Kmercount.txt" in outs["ref"] and "QC.kmerH.txt" in outs["ref"]
    assert_same_files(outs)


def test_blocked_gzip_inputs_are_inflated_in_parallel():
    """BGZF input: members are inflated by several threads straight into the batch buffer; a file that
    continues with ordinary gzip members falls through to zlib.  Same files as the reference reading the same .gz."""
    w = synth.c2(60000)
    outs = run_both({"-1": ("r1.bgz.gz", w.r1), "-2": ("r2.mixed.gz", w.r2)}, ["--discard"], threads=3, extra_cli=["--batch_mb", "7"])
    assert_same_files(outs)
    u = synth.c5(30000)
    outs = run_both({"-u": ("u.bgz.gz", u.r1)}, ["--mode", "HARD", "-q", "20", "--avg_q", "25"], threads=2, extra_cli=["--batch_mb", "2"])
    assert_same_files(outs)


def test_two_contexts_on_one_device():
    """--devices 0,0: two contexts on the same GPU take alternate batches (their kernels overlap on the device); the
    statistics merge locally (fq_merge_stats), the files are the reference's."""
    w = synth.c2(50000)
    outs = run_both({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, ["--discard", "-q", "15"], threads=4,
                    extra_cli=["--batch_mb", "2", "--devices", "0,0"])
    assert_same_files(outs)


def run_cli_only(inputs, flags, extra_cli=()):
    tmp = tempfile.mkdtemp(prefix="faqcs_cli_")
    try:
        args = []
        for flag, (name, data) in inputs.items():
            path = os.path.join(tmp, name)
            open(path, "wb").write(bytes(data))
            args += [flag, path]
        out = os.path.join(tmp, "gpu")
        p = subprocess.run([CLI, "-d", out, "-t", "2", "--debug"] + args + list(flags) + list(extra_cli), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr.decode(errors="replace")[-600:]
        return {n: open(os.path.join(out, n), "rb").read() for n in sorted(os.listdir(out)) if not n.endswith(".pdf")}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def test_gz_out_holds_the_reference_files():
    """--gz_out: the four streams as blocked gzip, deflated by several threads per batch; gunzip gives the
    reference's files byte for byte, every other file is unchanged."""
    w = synth.c2(60000)
    inputs = {"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}
    outs = run_both(inputs, ["--discard"], threads=2, extra_cli=["--batch_mb", "8", "--gz_out"])
    gpu = {}
    for name, data in outs["gpu"].items():
        if name.endswith(".fastq.gz"):
            assert data[:4] == b"\x1f\x8b\x08\x04" and data.endswith(b"\x1b\x00\x03\x00" + b"\x00" * 8)     # BGZF members, end-of-file member
            gpu[name[:-3]] = gzip.decompress(data)
        else:
            gpu[name] = data
    assert_same_files({"ref": outs["ref"], "gpu": gpu})


def test_keep_unpaired_appends_the_single_end_pass():
    """--keep_unpaired: the opt-in fix of the reference's Q11 quirk (the -u pass truncates the orphans the paired pass wrote to
    prefix.unpaired.trimmed.fastq): orphans of the paired pass, then the reads of the -u pass; plain and --gz_out."""
    w = synth.c2(20000)
    u = synth.c4(8000)
    paired_only = run_cli_only({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2)}, [])
    single_only = run_cli_only({"-u": ("u.fq", u.r1)}, [])
    want = paired_only["QC.unpaired.trimmed.fastq"] + single_only["QC.unpaired.trimmed.fastq"]
    assert paired_only["QC.unpaired.trimmed.fastq"] and single_only["QC.unpaired.trimmed.fastq"]
    both = run_cli_only({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2), "-u": ("u.fq", u.r1)}, ["--keep_unpaired"])
    assert both["QC.unpaired.trimmed.fastq"] == want
    quirk = run_cli_only({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2), "-u": ("u.fq", u.r1)}, [])
    assert quirk["QC.unpaired.trimmed.fastq"] == single_only["QC.unpaired.trimmed.fastq"]          # the reference's behaviour (Q11)
    gz = run_cli_only({"-1": ("r1.fq", w.r1), "-2": ("r2.fq", w.r2), "-u": ("u.fq", u.r1)}, ["--keep_unpaired", "--gz_out"])
    assert gzip.decompress(gz["QC.unpaired.trimmed.fastq.gz"]) == want
    assert gzip.decompress(gz["QC.1.trimmed.fastq.gz"]) == both["QC.1.trimmed.fastq"]


def test_ordinary_gzip_inputs_are_inflated_in_parallel():
    """Ordinary gzip input of some size: several threads enter the deflate stream at block boundaries
    (faqcs_b200/host/pgzip.hpp); mate 2 as two concatenated members.  Same files as the reference reading the same .gz."""
    w = synth.c2(70000)
    tmp = tempfile.mkdtemp(prefix="faqcs_pgz_")
    try:
        d2 = bytes(w.r2)
        cut = d2.index(b"\n@", len(d2) // 3) + 1
        z1, z2 = gzip.compress(bytes(w.r1), 6), gzip.compress(d2[:cut], 9) + gzip.compress(d2[cut:], 1)
        assert len(z1) > (4 << 20) and len(z2) > (4 << 20)
        p1, p2 = os.path.join(tmp, "r1.fastq.gz"), os.path.join(tmp, "r2.fastq.gz")
        open(p1, "wb").write(z1)
        open(p2, "wb").write(z2)
        outs = {}
        for tag, exe, more in (("ref", refcli.REF_BIN, []), ("gpu", CLI, ["--batch_mb", "9"])):
            out = os.path.join(tmp, tag)
            p = subprocess.run([exe, "-d", out, "-t", "2", "--debug", "-1", p1, "-2", p2, "--discard"] + more, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                               preexec_fn=lambda: signal.signal(signal.SIGPIPE, signal.SIG_IGN))
            assert p.returncode == 0, (tag, p.stderr.decode(errors="replace")[-600:])
            outs[tag] = {n: open(os.path.join(out, n), "rb").read() for n in sorted(os.listdir(out)) if not n.endswith(".pdf")}
        assert_same_files(outs)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
