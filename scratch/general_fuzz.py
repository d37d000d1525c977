"""Long general fuzz campaign (CUDA vs oracle: options x reads x line ends x batch cuts, then k-mer rarefaction on the same
kind of reads): python scratch/general_fuzz.py FIRST LAST"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from faqcs_b200.api import Engine, Options
from fuzz import fuzz_bytes, fuzz_options, fuzz_reads
from oracle_binding import OracleEngine
from test_gpu_parity import both
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(1000 + seed)
    in_off = 64 if seed % 4 == 3 else 33
    paired = seed % 2 == 0
    n = 500
    eol = "\r\n" if seed % 5 == 4 else "\n"
    r1 = fuzz_bytes(fuzz_reads(rng, n, in_off, "1" if paired else None), rng, eol)
    r2 = fuzz_bytes(fuzz_reads(rng, n, in_off, "2"), rng, eol) if paired else None
    kw = fuzz_options(rng, in_off, adapters=seed % 3 == 1)
    batch = None if kw.get("filter_adapter") else (int(rng.choice([0, 257])) or None)
    try:
        both(r1, r2, lambda: Options(**kw), batch_records=batch)
        # pieces mode and the pipelined entry points (submit / run / wait) on the same input: same streams, same statistics
        sg, st_g = both(r1, r2, lambda: Options(**kw), batch_records=None, check_results=False)
        with Engine(Options(**kw)) as pc:
            pc.set_output_pieces(True)
            pc.autodetect(r1, r2)
            t = pc.submit(r1, r2, 0, True)
            pc.run(t)
            res = pc.wait(t)
            ex = res.expand(r1, r2)
            assert [bytes(x) for x in ex] == [bytes(x) for x in sg], "pieces (pipelined) differ from byte mode"
            assert not pc.stats().diff(st_g), "statistics of the pieces run differ"
        # k-mer rarefaction over the same reads (one batch; k and the sampling cadence vary)
        k = int(rng.choice([2, 5, 11, 21, 31]))
        split = int(rng.choice([100, 400, 100000]))
        res = []
        for cls in (Engine, OracleEngine):
            with cls(Options(**kw)) as e:
                e.kmer_enable(k, split, 4)
                e.process(r1, r2, 0, True)
                e.kmer_end_pass()
                rare, freq = e.kmer_results()
                res.append((rare.tolist(), freq.tolist()))
        assert res[0] == res[1], f"k-mer results differ (k={k}, split={split}) paired={paired} eol={eol!r} opts={ {x: kw[x] for x in ('qc_only','replace_to_N_q','trim_5','trim_3','mode','output_quality_offset','filter_adapter') if x in kw} } gpu={res[0][0][:2]} {res[0][1][:3]} ora={res[1][0][:2]} {res[1][1][:3]}"
    except AssertionError as e:
        bad += 1
        print("seed", seed, "FAILED", str(e)[:700], flush=True)
print("seeds", sys.argv[1], "..", sys.argv[2], "failures", bad)
