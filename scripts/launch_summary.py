"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file list.csv ...`) per kernel: launches, total,
share, mean, max.  Usage: python scripts/launch_summary.py list.csv "<comment line>" ["<live stage shares line>"] > summary.csv"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki]).strip()
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print("# " + c)
print("kernel,launches,total_us,share_pct,mean_us,max_us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{a[0]},{a[1]:.1f},{100 * a[1] / tot:.1f},{a[1] / a[0]:.1f},{a[2]:.1f}")
print(f"# total,{sum(a[0] for a in agg.values())},{tot:.1f}")
