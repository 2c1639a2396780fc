"""Per-kernel HBM table from an ncu metrics CSV (gpu__time_duration, dram__bytes_read/write): avg duration, DRAM MB per launch,
achieved TB/s and % of the measured copy bandwidth (MEASURED_PEAKS.json: hbm_gbps, else 6570.6).
usage: python tools/bandwidth_table.py gpurun_out/bandwidth_kernels_r02.csv > profiles/bandwidth_kernels_r02.txt"""
import csv, json, os, re, sys
from collections import OrderedDict, defaultdict
path = sys.argv[1]
peak = 6570.6
try:
    mp = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
    for k in ("hbm_gbs", "hbm_gbps", "hbm_copy_gbps"):
        if k in mp:
            peak = float(mp[k]); break
except Exception:
    pass
L = defaultdict(dict)
for r in csv.DictReader(l for l in open(path) if not l.startswith("==")):
    i = int(r["ID"])
    L[i]["name"] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    L[i]["grid"] = r["Grid Size"]
    L[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", "") or 0)
    L[i]["unit:" + r["Metric Name"]] = r["Metric Unit"]
def to_bytes(d, m):
    v, u = d.get(m, 0.0), d.get("unit:" + m, "byte").lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
def to_us(d):
    v, u = d.get("gpu__time_duration.sum", 0.0), d.get("unit:gpu__time_duration.sum", "us").lower()
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3}.get(u, 1)
ids = sorted(L)
half = len(ids) // 2
agg = OrderedDict()
for i in ids[half:]:
    d = L[i]
    k = (d["name"][:46], d["grid"])
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += to_us(d); a[2] += to_bytes(d, "dram__bytes_read.sum") + to_bytes(d, "dram__bytes_write.sum")
print("measured copy bandwidth used as 100 %%: %.1f GB/s; second (profiled) eager step of bench.py --profile, cold caches" % peak)
print("%-46s %-16s %4s %9s %10s %8s %7s" % ("kernel", "grid", "n", "avg us", "MB/launch", "TB/s", "% peak"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    us, mb = a[1] / a[0], a[2] / a[0] / 1e6
    tbs = a[2] / a[1] / 1e6 if a[1] else 0.0
    print("%-46s %-16s %4d %9.1f %10.1f %8.2f %7.1f" % (k[0], k[1], a[0], us, mb, tbs, 100.0 * tbs * 1e3 / peak))
