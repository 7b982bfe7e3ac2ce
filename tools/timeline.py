#!/usr/bin/env python
"""Kernel timeline of one resident step (torch.profiler / CUPTI): device busy time, idle gaps and what surrounds them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spark-network-traffic-classifier_b200"))
import torch, bench
from torch.profiler import profile, ProfilerActivity
E2E = "--shim" in sys.argv
if E2E:
    sys.argv.remove("--shim")
a = bench.parse()
wl = bench.Workload(a)
rec, dicts = wl.make(a.rows, "cuda")
if E2E:                                                  # the pyspark.ml-shaped path from pinned host records
    host = rec.cpu().pin_memory(); del rec
    step = lambda: bench.step_e2e(wl, host, dicts, a)
else:
    step = lambda: bench.step_resident(wl, rec, dicts, a, None)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = 0.0; gaps = []; cur_end = ev[0].time_range.start
for i, e in enumerate(ev):
    s, t = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, ev[i - 1].name[:60], e.name[:60], (cur_end - t0) / 1e3))
    busy += max(0.0, t - max(s, cur_end)); cur_end = max(cur_end, t)
print("span %.3f ms, busy %.3f ms, idle %.3f ms, %d device activities" % ((t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(ev)))
gaps.sort(reverse=True)
print("largest gaps (us) : after -> before   @ms")
for g in gaps[:int(os.environ.get("GAPS", "40"))]:
    print("%8.1f  %-60s -> %-60s @%.2f" % g)
import collections
hist = collections.Counter()
for g in gaps:
    hist["<5us" if g[0] < 5 else "<20us" if g[0] < 20 else "<100us" if g[0] < 100 else ">=100us"] += g[0]
print({k: round(v / 1e3, 3) for k, v in hist.items()}, "ms by gap size;", len(gaps), "gaps")

import collections as _c
agg = _c.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e.name.split("(")[0][:70]; agg[k][0] += 1; agg[k][1] += (e.time_range.end - e.time_range.start) / 1e3
print("device activities by total time (ms):")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:int(os.environ.get("TOPK", "25"))]:
    print("%9.3f %5d  %s" % (v[1], v[0], k))
