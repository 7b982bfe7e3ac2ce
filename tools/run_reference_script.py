#!/usr/bin/env python
"""Run an UNMODIFIED reference driver script (e.g. /root/reference/code/network_traffic_classifier_kdd99.py) against the
b200flow pyspark shim: puts the shim on sys.path ahead of any real pyspark, stubs matplotlib/seaborn when they are not
installed (the CICIDS script only plots with them), then executes the script in the current working directory (which must
hold the dataset files the script opens by relative name)."""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spark-network-traffic-classifier_b200"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    if len(sys.argv) < 2:
        sys.exit("usage: run_reference_script.py <path/to/reference_script.py> [args...]")
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        noop = lambda *a, **k: None
        plt = _stub("matplotlib.pyplot", figure=noop, title=noop, xlabel=noop, ylabel=noop, draw=noop, tight_layout=noop, show=noop)
        _stub("matplotlib", pyplot=plt)
    try:
        import seaborn  # noqa: F401
    except ImportError:
        _stub("seaborn", set=lambda *a, **k: None, heatmap=lambda *a, **k: None)
    try:                                     # cicids17.py:103 passes `labels` positionally, which scikit-learn >= 1.0 rejects
        import sklearn.metrics as skm
        _cm = skm.confusion_matrix
        def confusion_matrix(y_true, y_pred, *args, **kw):
            if args:
                kw.setdefault("labels", args[0])
            return _cm(y_true, y_pred, **kw)
        skm.confusion_matrix = confusion_matrix
    except ImportError:
        pass
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
