"""Pins the CPU oracle against the upstream PySpark/MLlib doctest known answers
 — the only golden values available: the reference ships no tests."""
import numpy as np
import pytest

import oracle
from oracle import SLOT_DTYPE


def _records(cols):
    """pack equally long 4-byte columns (float32 / int32 arrays) into AoS records."""
    n = len(cols[0])
    rec = np.zeros((n, 4 * len(cols)), np.uint8)
    for j, c in enumerate(cols):
        rec[:, 4 * j:4 * j + 4] = np.ascontiguousarray(c).view(np.uint8).reshape(n, 4)
    return rec


def test_philox_reference_vectors():
    # Random123 known-answer vectors for philox4x32-10 (counter, key) -> output
    lib = oracle.lib()
    import ctypes as C
    # orc_philox xors the purpose into key0; purpose 0 gives the raw generator
    out = oracle.philox(0, 0, 0, 0, 0, 0)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = oracle.philox(0xffffffffffffffff, 0, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    out = oracle.philox(0x299f31d0a4093822, 0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)
    assert [hex(v) for v in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_string_indexer_doctest():
    labels = ["a", "b", "c"]
    codes = np.array([0, 1, 2, 0, 0, 2], np.int32)        # a b c a a c
    counts = oracle.category_counts(_records([codes]), 4, 0, 3)
    assert counts.tolist() == [3, 1, 2]
    ordered, lut = oracle.string_index_order(counts, labels)
    assert ordered == ["a", "c", "b"]
    plan = np.zeros(1, SLOT_DTYPE); plan["kind"] = 3; plan["lut_len"] = 3; plan["scale"] = 1.0
    out, _, valid = oracle.encode(_records([codes]), 4, plan, lut)
    assert out[:, 0].tolist() == [0.0, 2.0, 1.0, 0.0, 0.0, 1.0] and valid.all()


def test_string_indexer_tie_is_alphabetical_and_unseen_invalid():
    ordered, lut = oracle.string_index_order(np.array([2, 2, 0, 5]), ["b", "a", "zzz", "c"])
    assert ordered == ["c", "a", "b"] and lut.tolist() == [2, 1, -1, 0]
    plan = np.zeros(1, SLOT_DTYPE); plan["kind"] = 3; plan["lut_len"] = 4; plan["scale"] = 1.0
    _, _, valid = oracle.encode(_records([np.array([2, 3, 7], np.int32)]), 4, plan, lut)
    assert valid.tolist() == [0, 1, 0]


def test_vector_assembler_doctest():
    rec = _records([np.array([1], np.int32), np.array([0], np.int32), np.array([3], np.int32)])
    plan = np.zeros(3, SLOT_DTYPE); plan["kind"] = 2; plan["src_off"] = [0, 4, 8]; plan["scale"] = 1.0
    out, _, _ = oracle.encode(rec, 12, plan, None)
    assert out.tolist() == [[1.0, 0.0, 3.0]]


def test_vector_assembler_skip_marks_nan_rows():
    rec = _records([np.array([1.0, np.nan, np.inf], np.float32)])
    plan = np.zeros(1, SLOT_DTYPE); plan["kind"] = 0; plan["scale"] = 1.0
    _, _, valid = oracle.encode(rec, 4, plan, None, check_nan=1)
    assert valid.tolist() == [1, 0, 1]        # NaN dropped, Infinity kept


def test_standard_scaler_doctest():
    mean, std = oracle.moments(np.array([[0.0], [2.0]]))
    assert mean[0] == 1.0 and abs(std[0] - 1.4142135623730951) < 1e-15
    rec = _records([np.array([0.0, 2.0], np.float32)])
    plan = np.zeros(1, SLOT_DTYPE); plan["kind"] = 0; plan["scale"] = 1.0 / std[0]
    out, _, _ = oracle.encode(rec, 4, plan, None)
    assert np.allclose(out[:, 0], [0.0, 1.4142135623730951], rtol=1e-15)


def test_one_hot_encoder_doctest_droplast():
    codes = np.array([0, 1, 2], np.int32)
    lut = np.array([0, 1, 2], np.int32)
    plan = np.zeros(2, SLOT_DTYPE); plan["kind"] = 4; plan["lut_len"] = 3; plan["hot"] = [0, 1]; plan["scale"] = 1.0
    out, _, _ = oracle.encode(_records([codes]), 4, plan, lut)
    assert out.tolist() == [[1.0, 0.0], [0.0, 1.0], [0.0, 0.0]]


def test_find_splits_known_answers():
    assert oracle.find_splits_1d([2.0] * 10 + [3.0, 4.0, 5.0], 2).tolist() == [2.5, 3.5]
    assert oracle.find_splits_1d([0.0, 1.0] + [2.0] * 12, 2).tolist() == [0.5, 1.5]
    assert oracle.find_splits_1d([7.0] * 5, 3).tolist() == []


def test_decision_tree_doctest():
    x = np.array([[1.0], [0.0]]); y = np.array([1, 0], np.int32)
    fo, meta = oracle.fit_forest(x, y, 2, [0], num_trees=1, max_bins=32, max_depth=2)
    ex = fo.export()
    assert fo.num_nodes() == 3 and ex["is_leaf"].tolist() == [0, 1, 1]          # numNodes 3, depth 1
    assert meta["thresholds"][0, 0] == 0.5
    tp, _ = oracle.bin_rows(np.array([[-1.0], [1.0]]), meta["thresholds"], meta["n_thr"], [0], meta["max_bins"])
    raw, prob, pred = fo.predict(tp, dt_mode=True)
    assert raw[0].tolist() == [1.0, 0.0] and pred.tolist() == [0.0, 1.0]


def test_random_forest_doctest_shape():
    # upstream doctest: numTrees=3, maxDepth=2 on the same 2 rows predicts 0 for [-1], 1 for [1].
    # Spark's bagging RNG is irreproducible; with our Philox bagging the same holds for a fixed seed
    # in which the majority of trees see both rows.
    x = np.array([[1.0], [0.0]]); y = np.array([1, 0], np.int32)
    ok = 0
    for seed in range(20):
        fo, meta = oracle.fit_forest(x, y, 2, [0], num_trees=3, max_bins=32, max_depth=2, seed=seed)
        tp, _ = oracle.bin_rows(np.array([[-1.0], [1.0]]), meta["thresholds"], meta["n_thr"], [0], meta["max_bins"])
        _, prob, pred = fo.predict(tp)
        assert np.allclose(prob.sum(1), 1.0) or (meta["w"].sum() == 0)   # all-empty bags: no votes
        ok += pred.tolist() == [0.0, 1.0]
    assert ok >= 8


def test_multiclass_metrics_doctest():
    pairs = [(0, 0), (0, 1), (0, 0), (1, 0), (1, 1), (1, 1), (1, 1), (2, 2), (2, 0)]   # (pred, label)
    cm = oracle.confusion([p for p, _ in pairs], [l for _, l in pairs], 3)
    m = oracle.metrics(cm)
    assert abs(m["accuracy"] - 6 / 9) < 1e-12
    assert abs(m["f1"] - 0.661376) < 1e-6
    assert abs(m["macroF1"] - 0.662698) < 1e-6
    from sklearn.metrics import f1_score
    yt = [l for _, l in pairs]; yp = [p for p, _ in pairs]
    assert abs(m["f1"] - f1_score(yt, yp, average="weighted")) < 1e-12
    assert abs(m["macroF1"] - f1_score(yt, yp, average="macro")) < 1e-12


def test_poisson_table_and_bagging_mean():
    cdf = oracle.poisson_cdf_table(1.0)
    assert cdf[0] == int(np.floor(np.exp(-1.0) * 2 ** 32))
    w = oracle.bag_weights(7, 4, 200000, cdf)
    assert abs(w.mean() - 1.0) < 0.01 and abs((w == 0).mean() - np.exp(-1)) < 0.005
    # sharding independence: rows [1000,2000) drawn alone equal the slice of the full draw
    w2 = oracle.bag_weights(7, 4, 1000, cdf, row_offset=1000)
    assert (w2 == w[:, 1000:2000]).all()
    assert (oracle.bag_weights(7, 1, 10, None) == 1).all()


def test_metadata_unordered_rule():
    # KDD script case: 23 classes, maxBins 70 -> U = 6: arity 3 unordered, 11/70 ordered
    mpb, kind, m = oracle.build_metadata(10 ** 6, 41, 23, [0] * 38 + [3, 70, 11], 70, 20)
    assert kind[38:].tolist() == [2, 1, 1] and m == 7
    mpb, kind, m = oracle.build_metadata(10 ** 6, 41, 2, [0] * 38 + [3, 70, 11], 70, 1)
    assert kind[38:].tolist() == [1, 1, 1] and m == 41
    with pytest.raises(ValueError):
        oracle.build_metadata(10 ** 6, 41, 2, [0] * 38 + [3, 70, 11], 32, 1)
    assert oracle.build_metadata(10 ** 6, 78, 15, [0] * 78, 78, 20)[2] == 9


def test_shared_reciprocal_division_is_correctly_rounded():
    # forest.cu scores splits with q = RN(a*y), RN(q + (a - b*q)*y), y = RN(1/b) instead of a division per class count;
    # Markstein's theorem makes that the correctly rounded a/b for integer operands — checked here against `/`
    assert oracle.check_shared_reciprocal_division(3000, 100_000_000) == 0
