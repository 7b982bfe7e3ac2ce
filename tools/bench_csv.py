"""CSV ingest: KDD-shaped text -> typed device records, device reader (csrc/csv.cu) vs pandas on the host cores.
Prints one JSON line; not part of bench.py's contract (the headline excludes CSV parsing on both sides)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spark-network-traffic-classifier_b200"))
from b200flow import csvio, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    import pandas as pd
    rec, dicts = synth.make_kdd(n, 23, seed=1, device="cuda")
    a = rec.cpu().numpy().view(synth.kdd_schema().numpy_dtype()).reshape(-1)
    pdf = pd.DataFrame({c: (np.asarray(dicts[c], object)[a[c]] if c in dicts else a[c]) for c in synth.KDD_COLUMNS})
    for c in synth.KDD_RATE:
        pdf[c] = pdf[c].map(lambda v: "%.2f" % v)
    for c in synth.KDD_COLUMNS:
        if c not in dicts and c not in synth.KDD_RATE:
            pdf[c] = pdf[c].astype(np.int64)
    d = tempfile.mkdtemp()
    p = os.path.join(d, "kdd.csv")
    pdf.to_csv(p, header=False, index=False)
    size = os.path.getsize(p)
    out = {"rows": n, "csv_bytes": size}
    for _ in range(2):
        csvio.read_csv([p], False, True)                      # warm-up (page cache, kernels)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r, schema, dc = csvio.read_csv([p], False, True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    out["device_s"] = float(np.median(ts)); out["device_GBps"] = size / out["device_s"] / 1e9; out["device_rows_per_s"] = n / out["device_s"]
    phases = {}
    for _ in range(3):
        csvio.read_csv([p], False, True, stats=phases)
    out["phases_ms"] = {k: round(v / 3 * 1e3, 3) for k, v in phases.items()}
    t0 = time.perf_counter()
    pd.read_csv(p, header=None, float_precision="round_trip")
    out["pandas_s"] = time.perf_counter() - t0
    out["pandas_rows_per_s"] = n / out["pandas_s"]
    out["speedup_vs_pandas"] = out["pandas_s"] / out["device_s"]
    out["cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
