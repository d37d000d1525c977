#!/usr/bin/env python
"""Regenerate the committed golden fixtures by running the UNMODIFIED reference binary
(oracle/_ref/FaQCs, built from /root/reference by `make -C oracle ref`) on small seeded inputs.

    python tests/golden/make_golden.py

Each fixture is one .npz holding the inputs, the reference's emitted FASTQ streams, its
QC.stats.txt and the integers of its ten --debug matrix / histogram files.  The tests in
tests/test_golden.py replay them against the CPU oracle (always) and the CUDA path (-m gpu),
so parity stays pinned on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refcli  # noqa: E402
from faqcs_b200 import synth  # noqa: E402
from faqcs_b200.api import MODE_HARD, Options  # noqa: E402
from parity import MATRIX_FIELDS  # noqa: E402

AD = dict(__import__("faqcs_b200.api", fromlist=["BUILTIN_ADAPTERS"]).BUILTIN_ADAPTERS)


def micro_records():
    rng = np.random.default_rng(2026)
    rnd = lambda n, al="ACGT": "".join(rng.choice(list(al), size=n))
    noA = lambda n: rnd(n, "CGT")
    recs = [("@hard_all_low", "A" * 30 + "C" * 30, "#" * 60), ("@hard_high_low", "ACGT" * 15, "I" * 30 + "#" * 30),
            ("@termN", "NN" + rnd(70) + "TNN", "I" * 75), ("@allN", "N" * 64, "I" * 64),
            ("@mono86", "A" * 86 + "CGTCGTCGTCGTCG", "I" * 99 + "5"), ("@di", "AC" * 50, "I" * 99 + "5"),
            ("@q2", rnd(20) + AD["Nextera-primer-adapter-1"] + rnd(10) + AD["Nextera-primer-adapter-2"] + rnd(30), "I" * 124 + "5"),
            ("@tie", noA(30) + "A" * 20 + noA(30) + "A" * 20 + noA(10), "I" * 109 + "5"),
            ("@polyG", "G" * 80, "I" * 79 + "5")]
    for L in range(1, 30):
        q = "".join(chr(33 + int(x)) for x in rng.choice([2, 2, 2, 8, 20, 30, 40], size=L))
        recs.append((f"@short{L}", rnd(L, "ACGTN"), q))
    return recs


def example_reads(*names):
    """BASELINE configs[0] (C1): the reference tree ships only the OUTPUT of its example run
    (example/output/QC.{1,2}.trimmed.fastq, 17 607 real HiSeq pairs, 50-100 bp; QC.unpaired.trimmed.fastq),
    so those files are the inputs here."""
    d = "/root/reference/example/output"
    bufs = [np.fromfile(os.path.join(d, n), dtype=np.uint8) for n in names]
    return synth.Workload("c1", bufs[0], bufs[1] if len(bufs) > 1 else None, [])


# cases whose emitted streams are committed as (length, sha256) instead of bytes (fixture size)
HASHED = {"c1_example_paired", "c1_example_unpaired", "c2_2m_pairs"}
# cases whose INPUT is regenerated from the seeded generator at test time instead of being stored
REGENERATED = {"c2_2m_pairs": "synth.c2(2_000_000)"}

CASES = {
    # name: (workload factory, Options kwargs, reference extras)
    "c2_defaults": (lambda: synth.c2(1500), dict(discard_output=True), dict(threads=2)),
    "c4_qc_only": (lambda: synth.c4(3000), dict(qc_only=True), dict(threads=1)),
    "c5_hard_ascii64": (lambda: synth.c5(2500), dict(mode=MODE_HARD, quality=20, average_quality=25.0, replace_to_N_q=10,
                                                      discard_output=True), dict(threads=2)),
    "c3_adapters": (lambda: synth.c3(400), dict(filter_adapter=True, num_thread=1), dict(threads=1, polyA=True, artifacts=True)),
    "c1_example_paired": (lambda: example_reads("QC.1.trimmed.fastq", "QC.2.trimmed.fastq"), dict(discard_output=True), dict(threads=2)),
    "c1_example_unpaired": (lambda: example_reads("QC.unpaired.trimmed.fastq"), dict(), dict(threads=2)),
    # BASELINE batch size: 2 M pairs = one bench step, compared through hashes of the four streams + every statistic
    "c2_2m_pairs": (lambda: synth.c2(2_000_000), dict(discard_output=True), dict(threads=os.cpu_count() or 1)),
    # composition-bin detector lengths: an all-one-base read lands in bin 10000 except for L = 549, 579, 587
    "composition_detector_lengths": (lambda: synth.Workload("e16", np.frombuffer(synth.fastq_bytes(
        [(f"@L{L}_{b}", b * L, "I" * (L - 1) + "5") for L in (3, 150, 548, 549, 550, 579, 587, 588, 1000) for b in "ACGTN"]), dtype=np.uint8), None, []),
        dict(min_read_length=1, low_complexity_cutoff_ratio=1.0, max_num_poly_N=2000, input_quality_offset=33), dict(threads=1)),
    "micro_adapter_polya": (lambda: synth.Workload("micro", np.frombuffer(synth.fastq_bytes(micro_records()), dtype=np.uint8), None, []),
                            dict(filter_adapter=True, num_thread=1, min_read_length=1, low_complexity_cutoff_ratio=1.0, quality=10,
                                 input_quality_offset=33, discard_output=True), dict(threads=1, polyA=True)),
}


# k-mer rarefaction (--kmer_rarefaction): paired pass + unpaired pass, inputs regenerated from the seeded generator;
# the fixture holds the reference's QC.Kmercount.txt and QC.kmerH.txt.  (-m is not reachable through the reference's
# getopt_long_only -- it is an ambiguous abbreviation of --mode / --min_L -- so the reference always runs k = 31.)
KMER_CASES = {
    # name: (paired generator, unpaired generator, Options kwargs, split_size, subset, threads)
    "qc_only": ("synth.shotgun(70000)", "synth.shotgun(40000, seed=78, paired=False, L=100)", dict(qc_only=True), 30000, 2, 3),
    "qc_only_early_stop": ("synth.shotgun(70000)", "synth.shotgun(40000, seed=78, paired=False, L=100)", dict(qc_only=True), 10000, 3, 2),
    "qc_only_one_point": ("synth.shotgun(40000)", "synth.shotgun(10000, seed=78, paired=False, L=100)", dict(qc_only=True), 1000000, 10, 2),
    "trimmed": ("synth.shotgun(70000)", "synth.shotgun(40000, seed=78, paired=False, L=100)", dict(trim_5=3, quality=20), 50000, 1, 3),
    # G -> N below --replace_to_N_q happens before the k-mers of a surviving read are counted (trim.cpp:389-403, 545-547)
    "trimmed_replace_to_n": ("synth.shotgun(40000)", "synth.shotgun(35000, seed=78, paired=False, L=100)", dict(replace_to_N_q=30, quality=10), 30000, 2, 2),
}


def make_kmer(only):
    os.makedirs(os.path.join(HERE, "kmer"), exist_ok=True)
    for name, (gp, gu, okw, split, subset, threads) in KMER_CASES.items():
        if only and "kmer_" + name not in only:
            continue
        w, u = eval(gp, {"synth": synth}), eval(gu, {"synth": synth})
        flags = refcli.flags_for(Options(**okw)) + ["--kmer_rarefaction", "--split_size", str(split), "--subset", str(subset)]
        ref = refcli.run_reference(w.r1, w.r2, unpaired=u.r1, flags=flags, threads=threads)
        assert ref["returncode"] == 0, ref["stderr"]
        b = lambda x: np.frombuffer(x if isinstance(x, bytes) else x.encode(), dtype=np.uint8)
        path = os.path.join(HERE, "kmer", name + ".npz")
        np.savez_compressed(path, paired=b(gp), unpaired=b(gu), options=b(repr(okw)), params=np.array([31, split, subset], dtype=np.int64),
                            kmercount=b(ref["files"]["QC.Kmercount.txt"]), kmerh=b(ref["files"]["QC.kmerH.txt"]),
                            cmd=b(" ".join(["FaQCs"] + flags + ["-t", str(threads)])))
        print(f"kmer/{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    import hashlib
    assert refcli.have_ref(), "build the reference first: make -C oracle ref"
    only = set(sys.argv[1:])
    make_kmer(only)
    for name, (factory, okw, extra) in CASES.items():
        if only and name not in only:
            continue
        w = factory()
        opt = Options(**okw)
        polyA = extra.get("polyA", False)
        artifacts = w.artifacts if extra.get("artifacts") else None
        flags = refcli.flags_for(opt, polyA=polyA)
        if w.r2 is not None:
            ref = refcli.run_reference(w.r1, w.r2, flags=flags, threads=extra["threads"], artifacts=artifacts)
        else:
            ref = refcli.run_reference(unpaired=w.r1, flags=flags, threads=extra["threads"], artifacts=artifacts)
        assert ref["returncode"] == 0, ref["stderr"]
        adapters = refcli.adapters_for(opt.filter_adapter, polyA, artifacts)
        regen = name in REGENERATED
        out = dict(r1=np.zeros(0, np.uint8) if regen else np.asarray(w.r1),
                   r2=np.asarray(w.r2) if (w.r2 is not None and not regen) else np.zeros(0, np.uint8),
                   generator=np.frombuffer(REGENERATED.get(name, "").encode(), dtype=np.uint8),
                   paired=np.array([w.r2 is not None]), stats_txt=np.frombuffer(ref["stats_txt"].encode(), dtype=np.uint8),
                   options=np.frombuffer(repr(okw).encode(), dtype=np.uint8),
                   adapters=np.frombuffer(repr(adapters).encode(), dtype=np.uint8),
                   cmd=np.frombuffer(" ".join(["FaQCs"] + flags + ["-t", str(extra["threads"])]).encode(), dtype=np.uint8))
        for i, s in enumerate(ref["streams"]):
            if name in HASHED:
                out[f"stream{i}_sha256"] = np.frombuffer(hashlib.sha256(s).digest(), dtype=np.uint8)
                out[f"stream{i}_len"] = np.array([len(s)], dtype=np.int64)
            else:
                out[f"stream{i}"] = np.frombuffer(s, dtype=np.uint8)
        for f in MATRIX_FIELDS:
            out[f] = ref[f]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
