"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total ms, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip_first = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # launches of the warm-up step to drop
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(unit, 1)
    rows.append((r["Kernel Name"], ns))
rows = rows[skip_first:]
agg = defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    agg[short][0] += 1
    agg[short][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {tot / 1e6:.3f} ms (serialised, cold-cache: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6:9.3f} ms  {100 * v[1] / tot:5.1f}%  x{v[0]:4d}  {k[:110]}")
