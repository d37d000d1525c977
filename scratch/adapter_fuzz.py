"""Long adapter fuzz campaign (CUDA vs oracle): python scratch/adapter_fuzz.py FIRST LAST"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from faqcs_b200.api import Options
from faqcs_b200.synth import fastq_bytes
from fuzz import adapter_fuzz_case
from test_gpu_parity import both
bad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    recs, kw = adapter_fuzz_case(seed)
    r1 = np.frombuffer(fastq_bytes(recs), dtype=np.uint8)
    try:
        both(r1, None, lambda: Options(**kw))
    except AssertionError as e:
        bad += 1
        print("seed", seed, "FAILED", str(e)[:300], flush=True)
print("seeds", sys.argv[1], "..", sys.argv[2], "failures", bad)
