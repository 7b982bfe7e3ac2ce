"""Independent anchor for the oracle's tree path (R5, R7, R8, R9, R10): scikit-learn's exact CART.

The reference's arithmetic is Spark MLlib on a JVM, absent from this image, so the oracle cannot be pinned on
MLlib outputs.  What CAN be pinned: when every feature has fewer distinct values than maxBins, MLlib's candidate thresholds
are ALL midpoints between consecutive distinct values (findSplitsForContinuousFeature, A.2) — the same candidate set exact
CART scans — and with all features per node (numTrees = 1) and Gini impurity both algorithms pick the max-gain split.  Away
from exact gain ties the two trees must then be identical: same node count, same (feature, threshold) at every internal node,
same leaf class counts.  scikit-learn is an independent implementation; it is not the reference and not a parity oracle.
"""
import numpy as np
import pytest

import oracle

sklearn_tree = pytest.importorskip("sklearn.tree")
sklearn_metrics = pytest.importorskip("sklearn.metrics")


def _flows(seed, n=3000, F=6):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 20, size=(n, F)).astype(np.float64)           # 20 distinct values < maxBins 32; exact in float32
    score = (x[:, 0] > 9.5) * 1.0 + (x[:, 2] > 4.5) * 1.0 + (x[:, 4] > 14.5) * 0.7 + rng.normal(0, 0.6, n)
    return x, np.digitize(score, [0.8, 1.7]).astype(np.int32)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("depth", [1, 2, 3, 4])
def test_decision_tree_matches_exact_cart(seed, depth):
    x, y = _flows(seed)
    fo, meta = oracle.fit_forest(x, y, 3, [0] * x.shape[1], num_trees=1, max_bins=32, max_depth=depth, seed=seed)
    sk = sklearn_tree.DecisionTreeClassifier(criterion="gini", max_depth=depth, random_state=0).fit(x, y)
    ex = fo.export()
    assert fo.num_nodes() == sk.tree_.node_count
    # internal nodes: the (feature, threshold) multiset must agree
    ours = sorted((int(f), float(meta["thresholds"][f, b])) for f, b, leaf in zip(ex["feat"], ex["bin_thr"], ex["is_leaf"]) if not leaf)
    theirs = sorted((int(f), float(t)) for f, t in zip(sk.tree_.feature, sk.tree_.threshold) if f >= 0)
    assert ours == theirs
    raw, prob, pred = fo.predict(meta["tp"], dt_mode=True)
    assert np.array_equal(pred, sk.predict(x).astype(np.float64))
    assert np.abs(prob - sk.predict_proba(x)).max() < 1e-15
    # leaf class counts: raw prediction of a DecisionTree is the leaf's count vector (R9)
    leaf = sk.apply(x)
    counts = np.zeros((sk.tree_.node_count, 3))
    np.add.at(counts, (leaf, y), 1.0)
    assert np.array_equal(raw, counts[leaf])


def test_multiclass_metrics_match_sklearn():
    rng = np.random.default_rng(5)
    label = rng.integers(0, 5, 4000); label[label == 3] = 0          # a class with no true rows: excluded from the label set
    pred = np.where(rng.random(4000) < 0.7, label, rng.integers(0, 5, 4000))
    cm = oracle.confusion(pred.astype(np.float64), label.astype(np.float64), 5)
    assert np.array_equal(cm, sklearn_metrics.confusion_matrix(label, pred, labels=list(range(5))))
    mt = oracle.metrics(cm)
    acc, wp, wr, f1, macro = (mt[k] for k in ("accuracy", "weightedPrecision", "weightedRecall", "f1", "macroF1"))
    labs = sorted(set(label.tolist()))
    kw = dict(labels=labs, zero_division=0)
    assert abs(acc - sklearn_metrics.accuracy_score(label, pred)) < 1e-12
    assert abs(wp - sklearn_metrics.precision_score(label, pred, average="weighted", **kw)) < 1e-12
    assert abs(wr - sklearn_metrics.recall_score(label, pred, average="weighted", **kw)) < 1e-12
    assert abs(f1 - sklearn_metrics.f1_score(label, pred, average="weighted", **kw)) < 1e-12
    assert abs(macro - sklearn_metrics.f1_score(label, pred, average="macro", **kw)) < 1e-12
