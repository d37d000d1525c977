"""gpurun_out/other_{c3,c4,c5}.json (scratch/other_configs.sh) -> profiles/r2_other_workloads.json (what bench.py surfaces)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for wl in ("c3", "c4", "c5"):
    d = json.load(open(os.path.join(ROOT, "gpurun_out", f"other_{wl}.json")))
    out[wl] = {"reads_per_s": d["value"], "ms_per_batch": d["roofline"]["ms_per_batch"], "whole_path_hbm_frac": d["roofline"]["frac"],
               "kernels_ms": {k: v["ms"] for k, v in d["roofline"]["kernels"].items()},
               "pairs_or_reads_per_batch": d["config"]["device_batch_pairs"], "read_length": d["config"]["read_length"],
               "contexts_per_gpu": d["config"].get("contexts_per_gpu", 1), "result_check": d["config"]["result_check"]}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_other_workloads.json"), "w"), indent=1)
print(json.dumps({k: round(v["reads_per_s"] / 1e6, 1) for k, v in out.items()}))
