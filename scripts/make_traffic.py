"""Builds profiles/r02_traffic.json from the ncu --set full captures of scripts/final_r02.sh: DRAM bytes read / written and the
duration of one launch of each hot kernel, stamped with bench.py's fingerprint of the kernel sources (run it on the same
tree the captures were taken from).
usage: python scripts/make_traffic.py"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

REPORTS = {"tpch_q1_sf100": "gpurun_out/r02_q1_sf100_sq_agg_small.ncu-rep", "tpch_q3_sf100": "gpurun_out/r02_q3_sf100_kernels.ncu-rep",
           "tpch_q3_sf10": "gpurun_out/r02_q3_sf10_kernels.ncu-rep"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}

out = {"note": "DRAM bytes of ONE launch from ncu --set full --clock-control none (scripts/final_r02.sh); bench.py reports them as "
               "roofline.traffic while source_fingerprint matches the kernel sources",
       "source_fingerprints": {"q1": bench.source_fingerprint("q1"), "q3": bench.source_fingerprint("q3")}, "workloads": {}}
only = sys.argv[1:]  # e.g. "q3": refresh the Q3' captures only, keep the other workloads (and their fingerprint) as committed
if only:
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
        old = json.load(f)
    for q in ("q1", "q3"):
        if q not in only:
            out["source_fingerprints"][q] = old["source_fingerprints"][q]
    out["workloads"] = {w: v for w, v in old["workloads"].items() if not any(f"_{q}_" in w for q in only)}
for workload, rep in REPORTS.items():
    if only and not any(f"_{q}_" in workload for q in only):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    w = out["workloads"].setdefault(workload, {})
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        def val(metric):
            return float(r[col[metric]].replace(",", "")) * SCALE[units[col[metric]]]
        w[name] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                   "ncu_ms": round(val("gpu__time_duration.sum"), 6)}
with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
