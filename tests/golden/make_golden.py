#!/usr/bin/env python
"""Generates the committed golden fixtures (run from the repo root: `python tests/golden/make_golden.py`).

The reference's arithmetic (Spark MLlib) cannot run here (no JVM — SURVEY.md 8c) and the reference ships no golden vectors, so
these fixtures are NOT reference outputs: they are (a) the oracle's outputs on a small KDD-shaped batch, frozen so that the
oracle itself cannot drift between rounds and so that the CUDA path can be checked against committed numbers without calling
the oracle, and (b) scikit-learn's exact-CART tree on an exhaustively binned data set — an independent implementation that the
DecisionTree path must reproduce node for node (tests/test_oracle_vs_sklearn.py explains why)."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "spark-network-traffic-classifier_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import oracle
from b200flow import synth
from util import kdd_luts_oracle, kdd_plan, oracle_encode


def kdd_fixture():
    n = 3000
    rec, dicts = synth.make_kdd(n, 5, seed=424242, device="cpu")
    rec_np = rec.numpy()
    schema = synth.kdd_schema()
    luts, ordered = kdd_luts_oracle(rec_np, schema, dicts)
    plan = kdd_plan(schema, luts, ordered)
    x, y, valid = oracle_encode(plan, rec_np)
    full = kdd_plan(schema, luts, ordered, onehot=True)
    xf, _, _ = oracle_encode(full, rec_np)
    mean, std = oracle.moments(xf)
    arity = [0] * 38 + [len(ordered[c]) for c in synth.KDD_CATEGORICAL]
    C = len(ordered["label"])
    out = dict(records=rec_np, x=x, y=y, onehot_checksum=np.array([xf.sum(), (xf * xf).sum()]), mean=mean, std=std,
               arity=np.asarray(arity, np.int32), num_classes=np.int32(C))
    for c in synth.KDD_CATEGORICAL + ["label"]:
        out["lut_" + c] = np.asarray(luts[c], np.int32)
    sid = oracle.random_split(2019, n, [0.75, 1.0])
    out["split_id"] = sid
    tr = sid == 0
    for tag, kw in (("dt", dict(num_trees=1, max_bins=70, max_depth=5, seed=11)), ("rf", dict(num_trees=6, max_bins=70, max_depth=6, seed=2019))):
        fo, meta = oracle.fit_forest(x[tr], y[tr], C, arity, **kw)
        ex = fo.export()
        for k, v in ex.items():
            out["%s_%s" % (tag, k)] = v
        out[tag + "_thresholds"] = meta["thresholds"]; out[tag + "_n_thr"] = meta["n_thr"]
        tp, _ = oracle.bin_rows(x[~tr], meta["thresholds"], meta["n_thr"], meta["arity"], meta["max_bins"])
        raw, prob, pred = fo.predict(tp, dt_mode=(tag == "dt"))
        out[tag + "_raw"] = raw; out[tag + "_prob"] = prob; out[tag + "_pred"] = pred
        cm = oracle.confusion(pred, y[~tr].astype(np.float64), C)
        out[tag + "_confusion"] = cm
        mt = oracle.metrics(cm)
        out[tag + "_metrics"] = np.array([mt[k] for k in ("accuracy", "weightedPrecision", "weightedRecall", "f1", "macroF1")])
    np.savez_compressed(os.path.join(HERE, "kdd_small_oracle.npz"), **out)
    print("kdd_small_oracle.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]}, "...")


def cart_fixture():
    from sklearn.tree import DecisionTreeClassifier
    rng = np.random.default_rng(7)
    n, F = 4000, 8
    x = rng.integers(0, 24, size=(n, F)).astype(np.float64)
    score = (x[:, 1] > 11.5) * 1.0 + (x[:, 3] > 5.5) * 0.9 + (x[:, 6] > 17.5) * 0.8 + rng.normal(0, 0.55, n)
    y = np.digitize(score, [0.7, 1.6]).astype(np.int32)
    sk = DecisionTreeClassifier(criterion="gini", max_depth=4, random_state=0).fit(x, y)
    t = sk.tree_
    leaf = sk.apply(x)
    counts = np.zeros((t.node_count, 3)); np.add.at(counts, (leaf, y), 1.0)
    np.savez_compressed(os.path.join(HERE, "cart_exhaustive_sklearn.npz"), x=x.astype(np.int8), y=y.astype(np.int8),
                        node_count=np.int32(t.node_count), feature=t.feature.astype(np.int32), threshold=t.threshold,
                        pred=sk.predict(x).astype(np.int8), proba=sk.predict_proba(x), leaf_counts=counts[leaf])
    print("cart_exhaustive_sklearn.npz: nodes", t.node_count)


if __name__ == "__main__":
    kdd_fixture(); cart_fixture()
