"""k_kmer on one 2 M-pair batch: device time with and without --kmer_rarefaction (random reads = every k-mer distinct,
shotgun reads = 60 kb genome, heavy reuse)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from faqcs_b200 import synth
from faqcs_b200.api import Engine, Options

N = 32768 * 32
for name, w in (("random", synth.c2(N)), ("shotgun", synth.shotgun(N, genome_len=5_000_000))):
    for kmer in (False, True):
        with Engine(Options(qc_only=True)) as eng:
            if kmer:
                eng.kmer_enable(31, 1000000, 10)
            eng.autodetect(w.r1, w.r2)
            ts = []
            for it in range(3):
                torch.cuda.synchronize()
                t = time.perf_counter()
                eng.process(w.r1, w.r2, it * N, False)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t)
            if kmer:
                t = time.perf_counter()
                eng.kmer_end_pass()
                rare, freq = eng.kmer_results()
                print(name, "end_pass s", round(time.perf_counter() - t, 3), rare.tolist()[:3], len(freq))
            print(name, "kmer" if kmer else "plain", [round(x, 3) for x in ts], "s per", 2 * N, "reads (host call, incl. H2D)")
