"""Long campaign on the CPU: the oracle against the UNMODIFIED reference binary on the fuzz generator's inputs (emitted bytes,
every integer of QC.stats.txt, the ten --debug files), with --kmer_rarefaction on every third seed.
python scratch/oracle_fuzz.py FIRST LAST"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import refcli
from faqcs_b200 import synth
from faqcs_b200.api import Options
from fuzz import fuzz_bytes, fuzz_options, fuzz_reads
from oracle_binding import OracleEngine
from parity import assert_matches_reference, kmer_files, run_engine

bad = refused = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(2000 + seed)
    in_off = 64 if seed % 4 == 3 else 33
    paired = seed % 2 == 0
    eol = "\r\n" if seed % 5 == 4 else "\n"
    r1 = fuzz_bytes(fuzz_reads(rng, 600, in_off, "1" if paired else None), rng, eol)
    r2 = fuzz_bytes(fuzz_reads(rng, 600, in_off, "2"), rng, eol) if paired else None
    kw = fuzz_options(rng, in_off, adapters=seed % 3 == 1)
    polyA = bool(kw.pop("adapters", None))
    threads = kw.get("num_thread", 0) or 2
    opt = Options(**kw)
    kmer = seed % 3 == 2
    split = int(rng.choice([100, 400, 100000]))
    flags = refcli.flags_for(opt, polyA=polyA) + (["--kmer_rarefaction", "--split_size", str(split), "--subset", "2"] if kmer else [])
    ref = refcli.run_reference(r1, r2, flags=flags, threads=threads) if paired else refcli.run_reference(unpaired=r1, flags=flags, threads=threads)
    if ref["returncode"] != 0:
        refused += 1
        continue
    opt.adapters = refcli.adapters_for(opt.filter_adapter, polyA, None)
    try:
        with OracleEngine(opt) as eng:
            if kmer:
                eng.kmer_enable(31, split, 4)
            streams, _ = run_engine(eng, r1, r2)
            assert_matches_reference(ref, streams, eng.stats(), opt, opt.adapters)
            if kmer:
                eng.kmer_end_pass()
                kc, kh = kmer_files(*eng.kmer_results())
                # plot() writes the two files only when some k-mer was counted (plot.cpp:85-91)
                if "QC.kmerH.txt" in ref["files"]:
                    assert kc == ref["files"]["QC.Kmercount.txt"] and kh == ref["files"]["QC.kmerH.txt"], "k-mer files differ"
                else:
                    assert kh == b"", "k-mers counted where the reference counted none"
    except AssertionError as e:
        bad += 1
        print("seed", seed, "FAILED", str(e)[:300], flags, flush=True)
print("seeds", sys.argv[1], "..", sys.argv[2], "failures", bad, "refused by the reference", refused)
