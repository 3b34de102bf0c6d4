#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals/shares.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    name = r.get("Kernel Name") or r.get("Kernel") or ""
    if (r.get("Metric Name") or "") != "gpu__time_duration.sum":
        continue
    v = float((r.get("Metric Value") or "0").replace(",", ""))
    unit = (r.get("Metric Unit") or "ns").strip()
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9, "s": 1e9}.get(unit, 1)
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short).replace("wlk::", "").replace("<unnamed>::", "")
    tot[short][0] += 1
    tot[short][1] += ns
total = sum(v[1] for v in tot.values()) or 1.0
print(f"# kernel launch list summary ({path})\n")
print("ncu serialises launches and runs them cold-cache: compare SHARES, not absolute times.\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |")
print(f"\ntotal {total / 1e6:.2f} ms over {sum(v[0] for v in tot.values())} launches")
