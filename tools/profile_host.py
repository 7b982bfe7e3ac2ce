#!/usr/bin/env python
"""cProfile of one resident step (host side): where the Python/driver time between kernels goes."""
import cProfile, pstats, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spark-network-traffic-classifier_b200"))
import torch, bench
from b200flow import synth
a = bench.parse()
rec, dicts = synth.make_kdd(a.rows, a.classes, seed=2019, device="cuda")
for _ in range(2):
    bench.step_resident(rec, dicts, a, None)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
bench.step_resident(rec, dicts, a, None)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
