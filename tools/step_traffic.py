"""Whole-step DRAM traffic from an ncu launch list of `bench.py --profile` (one warm-up + one profiled step: the second
half of the launches) -> profiles/step_traffic.json {workload: {dram_bytes_per_step, launches, serialised_us, source}},
read by bench.py's `hbm` block.   usage: python tools/step_traffic.py launches.csv workload [out.json]"""
import csv
import json
import os
import sys
from collections import defaultdict

path, wl = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "step_traffic.json")
lines = [l for l in open(path) if not l.startswith("==")]
L = defaultdict(dict)
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0}
for r in csv.DictReader(lines):
    L[int(r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SCALE.get(r.get("Metric Unit", ""), 1.0)
ids = sorted(L)
step = ids[len(ids) // 2:]
tot_b = sum(L[i].get("dram__bytes_read.sum", 0.0) + L[i].get("dram__bytes_write.sum", 0.0) for i in step)
tot_us = sum(L[i].get("gpu__time_duration.sum", 0.0) for i in step) / 1e3
cur = json.load(open(out)) if os.path.exists(out) else {}
cur[wl] = {"dram_bytes_per_step": tot_b, "launches": len(step), "serialised_us": tot_us,
           "source": "profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step)" % os.path.basename(path)}
json.dump(cur, open(out, "w"), indent=1)
print(json.dumps(cur[wl], indent=1))
