"""CPU campaign: the oracle against the reference binary on the adapter-centred fuzz cases (tests/fuzz.py::adapter_fuzz_case),
the case's adapters handed to the reference as its --artifactFile.   python scratch/oracle_adapter_fuzz.py FIRST LAST"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import refcli
from faqcs_b200.api import Options
from faqcs_b200.synth import fastq_bytes
from fuzz import adapter_fuzz_case
from oracle_binding import OracleEngine
from parity import assert_matches_reference, run_engine

bad = refused = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    recs, kw = adapter_fuzz_case(seed)
    artifacts = [a for a in kw.pop("adapters") if a[0].startswith("A") and a[0][1:].isdigit()]
    threads = kw.get("num_thread", 0) or 2
    kw["num_thread"] = threads
    opt = Options(**kw)
    r1 = np.frombuffer(fastq_bytes(recs), dtype=np.uint8)
    ref = refcli.run_reference(unpaired=r1, flags=refcli.flags_for(opt), threads=threads, artifacts=artifacts)
    if ref["returncode"] != 0:
        refused += 1
        continue
    opt.adapters = refcli.adapters_for(True, False, artifacts)
    opt.filter_adapter = True
    try:
        with OracleEngine(opt) as eng:
            streams, _ = run_engine(eng, r1, None)
            assert_matches_reference(ref, streams, eng.stats(), opt, opt.adapters)
    except AssertionError as e:
        bad += 1
        print("seed", seed, "FAILED", str(e)[:300], flush=True)
print("seeds", sys.argv[1], "..", sys.argv[2], "failures", bad, "refused by the reference", refused)
