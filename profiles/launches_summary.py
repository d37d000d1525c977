"""Summarise an ncu launch list (scratch/launches.sh: gpu__time_duration + DRAM bytes per launch): the launches of the LAST
device batch of the run (10 kernels per paired batch), shares per segment, DRAM bytes against the algorithmic bytes.
usage: python profiles/launches_summary.py gpurun_out/r2_launches_final.csv [launches_per_batch] [algorithmic_GB]"""
import csv, sys
path = sys.argv[1]
per_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 10
alg = float(sys.argv[3]) if len(sys.argv) > 3 else 1.337
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 10]
hdr = rows[0]
iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iU, iID = hdr.index("Metric Unit"), hdr.index("ID")
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[iID]), {"name": r[iK].split("(")[0]})
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    d[r[iM]] = v * scale
ids = sorted(launches)[-per_batch:]
tot = sum(launches[i]["gpu__time_duration.sum"] for i in ids)
seg = {"frame": 0.0, "trim": 0.0, "emit": 0.0}
dram = 0.0
print(f"launches of the last 1 M-pair batch ({path}; ncu: cold cache, serialised -- compare shares):")
for i in ids:
    L = launches[i]
    t = L["gpu__time_duration.sum"]
    rd, wr = L.get("dram__bytes_read.sum", 0.0), L.get("dram__bytes_write.sum", 0.0)
    dram += rd + wr
    key = "trim" if "k_trim" in L["name"] else "emit" if any(k in L["name"] for k in ("k_route", "k_scan_tiles", "k_emit")) else "frame"
    seg[key] += t
    print(f"  {L['name']:<30} {t:8.1f} us  {100 * t / tot:5.1f} %   DRAM read {rd:8.1f} MB  written {wr:8.1f} MB")
print(f"total {tot:.1f} us; shares: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in seg.items()) +
      f"; DRAM bytes moved {dram / 1e3:.3f} GB for {alg:.3f} GB algorithmic = {dram / 1e3 / alg:.2f} x")
