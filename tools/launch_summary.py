"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/launch_summary.py FILE [steps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
kn, mv = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[kn])[:72]
    agg[name][0] += 1
    agg[name][1] += v
total = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values()) / steps:.0f} launches, {total / 1e3 / steps:.1f} us per step (cold-cache, serialised)")
for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print(f"{t / 1e3 / steps:9.1f} us {c / steps:6.1f}  {name}")
