#!/usr/bin/env python
"""Write synthetic dataset files in the formats the reference scripts open (no real data in this environment):
  kdd     -> <dir>/kddcup.data.corrected       (no header, 42 columns, labels end with '.')   [kdd99.py:25]
  cicids  -> <dir>/{Monday,Tuesday}-WorkingHours.pcap_ISCX.csv (header with leading blanks, 78 features + ' Label',
             'Fwd Header Length' twice, NaN/Infinity cells, one label containing U+FFFD)     [cicids17.py:19-27]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spark-network-traffic-classifier_b200"))
import numpy as np
from b200flow import synth


def write_kdd(d, n, seed):
    rec, dicts = synth.make_kdd(n, 23, seed=seed, device="cpu")
    a = rec.numpy().view(synth.kdd_schema().numpy_dtype()).reshape(-1)
    with open(os.path.join(d, "kddcup.data.corrected"), "w") as f:
        for r in a:
            cells = []
            for name in synth.KDD_COLUMNS:
                v = r[name]
                if name in dicts:
                    cells.append(dicts[name][int(v)] + ("." if name == "label" else ""))
                elif name in synth.KDD_RATE:
                    cells.append("%.2f" % v)
                else:
                    cells.append("%d" % int(v))
            f.write(",".join(cells) + "\n")


def write_cicids(d, n, seed):
    rec, dicts = synth.make_cicids(n, 15, seed=seed, device="cpu", nan_fraction=0.002)
    a = rec.numpy().view(synth.cicids_schema().numpy_dtype()).reshape(-1)
    names = [" Feature %d" % i for i in range(78)]
    for i, nm in {0: " Flow Duration", 3: " Init_Win_bytes_forward", 6: " Init_Win_bytes_backward", 9: " Flow IAT Min",
                  12: " Fwd IAT Min", 15: " Fwd IAT Max", 34: " Fwd Header Length", 55: " Fwd Header Length"}.items():
        names[i] = nm
    names.append(" Label")
    filt = (0, 3, 6, 9, 12, 15)                           # the six columns cicids17.py:30-35 requires to be > 0
    rng = np.random.default_rng(seed)
    killed = rng.random(len(a)) < 0.15                    # ~15 % of the flows fail a filter (real data: 73 %)
    for part, fn in enumerate(["Monday-WorkingHours.pcap_ISCX.csv", "Tuesday-WorkingHours.pcap_ISCX.csv"]):
        with open(os.path.join(d, fn), "w", encoding="utf-8") as f:
            f.write(",".join(names) + "\n")
            for ridx in range(part, len(a), 2):
                r = a[ridx]
                lab = dicts["Label"][int(r["Label"])]
                if lab.startswith("Web Attack"):
                    lab = lab.replace("Web Attack ", "Web Attack � ")
                cells = []
                for i in range(78):
                    v = float(r["f%02d" % i])
                    if i in filt and v == v:
                        v = -1.0 if (killed[ridx] and i == 3) else abs(np.floor(v)) + 1.0
                    cells.append("NaN" if v != v else ("%d" % v if v == int(v) else repr(v)))
                f.write(",".join(cells) + "," + lab + "\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("kind", choices=["kdd", "cicids"]); ap.add_argument("dir")
    ap.add_argument("--rows", type=int, default=100000); ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    (write_kdd if a.kind == "kdd" else write_cicids)(a.dir, a.rows, a.seed)
