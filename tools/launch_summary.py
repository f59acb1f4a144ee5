"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total per kernel, share of the run, and the
launches of the headline kernel.   python tools/launch_summary.py profiles/r1_launches.csv [headline-kernel-regex]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
idx = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    val = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    rows.append((r[idx["Kernel Name"]], ms))
tot = sum(ms for _, ms in rows)
agg = defaultdict(lambda: [0.0, 0])
for k, ms in rows:
    agg[k][0] += ms
    agg[k][1] += 1
print(f"# total {tot:.3f} ms over {len(rows)} launches")
for k, (ms, n) in sorted(agg.items(), key=lambda t: -t[1][0]):
    print(f"  {ms:10.3f} ms  n={n:4d}  avg {ms / n:8.4f} ms  {100 * ms / tot:5.1f}%  {k[:120]}")
if pat:
    print(f"\n# launches matching /{pat.pattern}/:")
    for k, ms in rows:
        if pat.search(k):
            print(f"  {ms:9.4f} ms  {k[:100]}")
