#!/usr/bin/env python
"""ncu launch-list CSV (`--metrics gpu__time_duration.sum --csv`) -> markdown table of shares per kernel.

    python tools/launch_list.py gpurun_out/launches.csv "title" > profiles/r2_launch_list.md"""
import collections
import csv
import io
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
lines = [l for l in open(path, errors="replace") if l.startswith('"')]
rows = list(csv.reader(io.StringIO("".join(lines))))
hdr = rows[0]
ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "ns ": 1e-6}
by = collections.defaultdict(lambda: [0, 0.0])
n = 0
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    ms = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
    name = r[ki].split("(")[0].replace("void ", "")
    by[name][0] += 1
    by[name][1] += ms
    n += 1
tot = sum(v[1] for v in by.values())
print(f"# {title}\n")
print(f"{n} launches, {tot:.2f} ms of device time (cold-cache, serialised by ncu, clocks not locked: compare SHARES with the "
      "`kernel_families` / timeline of the bench line, not absolutes).\n")
print("| share | ms | launches | kernel |\n|---|---|---|---|")
for name, (c, ms) in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {100 * ms / tot:.1f}% | {ms:.3f} | {c} | `{name[:100]}` |")
