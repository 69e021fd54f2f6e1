#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in r:
    name = row[ki].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*", "", name).replace("__nv_bfloat16", "bf16")
    name = re.sub(r"at::native::.*?<", "at::<", name)[:70]
    tot[name][0] += 1; tot[name][1] += float(row[vi].replace(",", "")); n += 1
s = sum(v[1] for v in tot.values())
print(f"{n} launches, {s / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k:70s} n={v[0]:4d} {v[1] / 1e6:8.3f} ms {100 * v[1] / s:5.1f}%  avg {v[1] / v[0] / 1e3:8.1f} us")
