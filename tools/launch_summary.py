#!/usr/bin/env python
"""summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py --steps K --warmup W`: the complete
resident steps only (a step starts at a category_counts launch; data generation and warm-up before the first one are cut)."""
import collections
import csv
import re
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
ev = []
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(r[ui], 1e-6)
    ev.append((r[ki], v))
starts = [i for i, (k, _) in enumerate(ev) if "category_counts" in k]
first = starts[-(steps + 1)] if len(starts) > steps else starts[0]
last = starts[-1] if len(starts) > steps else len(ev)
n_steps = max(1, len([s for s in starts if first <= s < last]))
sel = ev[first:last]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in sel:
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"\(.*$", "", k)
    k = k.replace("at::native::", "at::").replace("(anonymous namespace)::", "")
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in agg.values())
ours = [(k, c, v) for k, (c, v) in agg.items() if "b200flow" in k]
print("# launch list of `python bench.py --steps %d --warmup 1 --no-cpu-baseline --no-e2e` (KDD99-full): %d complete resident step(s), data generation excluded" % (steps, n_steps))
print("# source: ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 (%s); serialised and cold-cache, so compare SHARES, not absolutes" % path)
print("# per step: %d launches, %.2f ms of kernel time; b200flow kernels: %d launches, %.2f ms (%.1f %%); torch helper kernels (fills, copies, index ops): %d launches, %.2f ms"
      % (len(sel) / n_steps, tot / n_steps, sum(c for _, c, _ in ours) / n_steps, sum(v for _, _, v in ours) / n_steps, 100 * sum(v for _, _, v in ours) / tot,
         (len(sel) - sum(c for _, c, _ in ours)) / n_steps, (tot - sum(v for _, _, v in ours)) / n_steps))
print("%-72s %9s %10s %7s" % ("kernel", "per step", "ms / step", "share"))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %9.1f %10.3f %6.1f%%" % (k[:72], c / n_steps, v / n_steps, 100 * v / tot))
