"""Per-kernel DRAM traffic from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list -> JSON {kernel: {launches, dram_bytes_per_launch, avg_us}} (bench.py reads the conv_tc entry for
`roofline.traffic`).  Usage: python tools/traffic_from_ncu.py launches.csv out.json [kernel-substring ...]"""
import csv
import json
import re
import sys
from collections import defaultdict

path, out = sys.argv[1], sys.argv[2]
want = sys.argv[3:] or ["conv_tc_kernel", "conv_tc_wgrad_kernel", "conv_tc_s2_kernel"]
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3}
per_id = defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", "")) * SCALE.get(r.get("Metric Unit", ""), 1.0)
    per_id[r["ID"]]["name"] = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))
    per_id[r["ID"]][r["Metric Name"]] = v
agg = defaultdict(lambda: [0, 0.0, 0.0])
for d in per_id.values():
    for w in want:
        if w in d["name"]:
            a = agg[w]
            a[0] += 1
            a[1] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            a[2] += d.get("gpu__time_duration.sum", 0.0)
res = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / max(v[0], 1), "avg_us": v[2] / max(v[0], 1)} for k, v in agg.items()}
res["source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, python bench.py --profile"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
