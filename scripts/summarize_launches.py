"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::", "", name)
        if "gemm" in name:   # GEMMs: one line per tile configuration and grid (decode-step vs encoder problems)
            name += " grid" + r.get("Grid Size", "").replace(" ", "")
        rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"launches {len(rows)}  total {tot / 1e6:.3f} ms (serialised, cold-cache; compare shares)")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * ns / tot:6.2f}%  {ns / 1e6:10.3f} ms  {c:6d}x  avg {ns / c / 1e3:9.1f} us  {n[:110]}")
