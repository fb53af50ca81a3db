"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v))
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = k.replace("<unnamed>::", "").replace("void ", "")
    k = re.sub(r"\(.*$", "", k)            # drop the argument list
    k = re.sub(r"^at::native::", "torch:", k)
    k = re.sub(r"<at::native.*", "<...>", k)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'n':>7s} {'ms':>10s} {'share':>7s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k[:60]:60s} {n:7d} {v/1e3:10.3f} {100*v/tot:6.1f}%")
