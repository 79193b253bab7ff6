"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python scripts/ncu_launch_summary.py launches.csv "note" > profiles/rNN_ncu_launch_summary.json"""
import csv
import json
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
ks = [{"kernel": k[:160], "launches": a[0], "total_us": round(a[1], 1), "avg_us": round(a[1] / a[0], 2), "share": round(a[1] / tot, 4)}
      for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
print(json.dumps({"note": sys.argv[2] if len(sys.argv) > 2 else "", "total_us": tot, "launches": sum(k["launches"] for k in ks),
                  "kernels": ks}, indent=1))
