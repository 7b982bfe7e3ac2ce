"""ctypes front-end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY — see the header of oracle.cpp.  PARITY UNPINNED at the MLlib
boundary (no JVM here; the reference has no tests); pinned only against the upstream
doctest known answers.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

SLOT_DTYPE = np.dtype([("kind", "<i4"), ("src_off", "<i4"), ("lut_off", "<i4"), ("lut_len", "<i4"),
                       ("hot", "<i4"), ("reserved", "<i4"), ("mean", "<f8"), ("scale", "<f8")])

PURPOSE_SAMPLE, PURPOSE_BAG, PURPOSE_FEAT, PURPOSE_RSPLIT = 0x53414D50, 0x42414747, 0x46454154, 0x5253504C


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_rf_train.restype = C.c_void_p
        _lib.orc_forest_num_nodes.restype = C.c_int64
        _lib.orc_find_splits_1d.restype = C.c_int32
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n=None):
    """use n OpenMP threads (default: every core this process may run on), whatever OMP_NUM_THREADS says —
    torchrun exports OMP_NUM_THREADS=1 to its workers."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def philox(seed, purpose, c0, c1=0, c2=0, c3=0):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(C.c_uint64(seed), C.c_uint32(purpose), C.c_uint32(c0), C.c_uint32(c1),
                     C.c_uint32(c2), C.c_uint32(c3), _p(out))
    return out


def category_counts(records, row_bytes, src_off, K):
    records = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
    n = records.size // row_bytes
    counts = np.zeros(K, np.int64)
    lib().orc_category_counts(_p(records), C.c_int64(n), C.c_int32(row_bytes), C.c_int32(src_off),
                              C.c_int32(K), _p(counts))
    return counts


def string_index_order(counts, labels):
    """StringIndexer.fit ordering (A.7): frequencyDesc, ties alphabetical ascending.
    Returns (ordered label list, lut code->rank) dropping never-seen codes."""
    idx = [i for i in range(len(labels)) if counts[i] > 0]
    idx.sort(key=lambda i: (-int(counts[i]), labels[i]))
    lut = np.full(len(labels), -1, np.int32)
    for rank, i in enumerate(idx):
        lut[i] = rank
    return [labels[i] for i in idx], lut


def encode(records, row_bytes, plan, lut, label_off=-1, label_lut_off=0, label_lut_len=0, check_nan=0):
    records = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
    n = records.size // row_bytes
    plan = np.ascontiguousarray(plan, dtype=SLOT_DTYPE)
    lut = np.ascontiguousarray(lut if lut is not None else np.zeros(1), dtype=np.int32)
    out = np.empty((n, len(plan)), np.float64)
    lab = np.empty(n, np.int32)
    valid = np.empty(n, np.uint8)
    lib().orc_encode(_p(records), C.c_int64(n), C.c_int32(row_bytes), _p(plan), C.c_int32(len(plan)), _p(lut),
                     C.c_int32(label_off), C.c_int32(label_lut_off), C.c_int32(label_lut_len),
                     C.c_int32(check_nan), _p(out), _p(lab), _p(valid))
    return out, lab, valid


def moments(x):
    x = np.ascontiguousarray(x, np.float64)
    n, D = x.shape
    mean = np.empty(D); std = np.empty(D)
    lib().orc_moments(_p(x), C.c_int64(n), C.c_int32(D), _p(mean), _p(std))
    return mean, std


def find_splits(x, seed, keep_threshold, arity, max_bins, row_offset=0):
    x = np.ascontiguousarray(x, np.float64)
    n, F = x.shape
    arity = np.ascontiguousarray(arity, np.int32)
    thr = np.zeros((F, max_bins - 1), np.float64)
    n_thr = np.zeros(F, np.int32)
    ns = C.c_int32(0)
    lib().orc_find_splits(_p(x), C.c_int64(n), C.c_int32(F), C.c_uint64(seed), C.c_uint64(keep_threshold),
                          C.c_int64(row_offset), _p(arity), C.c_int32(max_bins), _p(thr), _p(n_thr), C.byref(ns))
    return thr, n_thr, ns.value


def find_splits_1d(samples, num_splits):
    s = np.ascontiguousarray(samples, np.float64)
    thr = np.zeros(max(num_splits, 1))
    nt = lib().orc_find_splits_1d(_p(s), C.c_int32(len(s)), C.c_int32(num_splits), _p(thr))
    return thr[:nt].copy()


def tp_stride(F):
    return (F + 1 + 15) // 16 * 16


def bin_rows(x, thresholds, n_thr, arity, max_bins, labels=None):
    x = np.ascontiguousarray(x, np.float64)
    n, F = x.shape
    stride = tp_stride(F)
    tp = np.zeros((n, stride), np.uint8)
    bad = C.c_int32(0)
    lab = None if labels is None else np.ascontiguousarray(labels, np.int32)
    lib().orc_bin_rows(_p(x), C.c_int64(n), C.c_int32(F), _p(np.ascontiguousarray(thresholds, np.float64)),
                       _p(np.ascontiguousarray(n_thr, np.int32)), _p(np.ascontiguousarray(arity, np.int32)),
                       C.c_int32(max_bins), _p(lab), _p(tp), C.c_int32(stride), C.byref(bad))
    return tp, bad.value


def bag_weights(seed, T, n, cdf, row_offset=0):
    w = np.empty((T, n), np.uint8)
    cdf_a = None if cdf is None else np.ascontiguousarray(cdf, np.uint32)
    lib().orc_bag_weights(C.c_uint64(seed), C.c_int32(T), C.c_int64(row_offset), C.c_int64(n), _p(cdf_a), _p(w))
    return w


def feature_subset(seed, tree, nid, F, m):
    out = np.empty(m, np.int32)
    lib().orc_feature_subset(C.c_uint64(seed), C.c_int32(tree), C.c_uint32(nid), C.c_int32(F), C.c_int32(m), _p(out))
    return out


def hist_node(tp, F, rows, w, subset, n_bins, Cc):
    tp = np.ascontiguousarray(tp, np.uint8)
    rows = np.ascontiguousarray(rows, np.int32); w = np.ascontiguousarray(w, np.uint8)
    subset = np.ascontiguousarray(subset, np.int32)
    hist = np.zeros((len(subset), n_bins, Cc), np.int64)
    lib().orc_hist_node(_p(tp), C.c_int32(tp.shape[1]), C.c_int32(F), _p(rows), _p(w), C.c_int64(len(rows)),
                        _p(subset), C.c_int32(len(subset)), C.c_int32(n_bins), C.c_int32(Cc), _p(hist))
    return hist


class Forest:
    """Handle on an oracle-trained forest (R7+R8) with predict (R9) and canonical export."""

    def __init__(self, handle, Cc, T):
        self._h, self.C, self.T = handle, Cc, T

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().orc_forest_free(C.c_void_p(self._h)); self._h = None
        except Exception:          # interpreter shutdown
            pass

    def num_nodes(self):
        return int(lib().orc_forest_num_nodes(C.c_void_p(self._h)))

    def export(self):
        n = self.num_nodes()
        d = dict(tree=np.empty(n, np.int32), nid=np.empty(n, np.uint32), feat=np.empty(n, np.int32),
                 kind=np.empty(n, np.int32), bin_thr=np.empty(n, np.int32), is_leaf=np.empty(n, np.int32),
                 gain=np.empty(n), impurity=np.empty(n), mask=np.empty((n, 4), np.uint64),
                 counts=np.empty((n, self.C), np.int64))
        lib().orc_forest_export(C.c_void_p(self._h), *[_p(d[k]) for k in
                                ("tree", "nid", "feat", "kind", "bin_thr", "is_leaf", "gain", "impurity", "mask", "counts")])
        return d

    def predict(self, tp, dt_mode=False):
        tp = np.ascontiguousarray(tp, np.uint8)
        n = tp.shape[0]
        raw = np.empty((n, self.C)); prob = np.empty((n, self.C)); pred = np.empty(n)
        lib().orc_rf_predict(C.c_void_p(self._h), _p(tp), C.c_int64(n), C.c_int32(tp.shape[1]),
                             C.c_int32(1 if dt_mode else 0), _p(raw), _p(prob), _p(pred))
        return raw, prob, pred


def rf_train(tp, F, Cc, w, feat_bins, feat_kind, n_bins, m, max_depth, min_instances, min_info_gain, seed):
    tp = np.ascontiguousarray(tp, np.uint8)
    w = np.ascontiguousarray(w, np.uint8)
    T, n = w.shape
    h = lib().orc_rf_train(_p(tp), C.c_int64(n), C.c_int32(F), C.c_int32(tp.shape[1]), C.c_int32(Cc), C.c_int32(T),
                           _p(w), _p(np.ascontiguousarray(feat_bins, np.int32)),
                           _p(np.ascontiguousarray(feat_kind, np.int32)), C.c_int32(n_bins), C.c_int32(m),
                           C.c_int32(max_depth), C.c_int32(min_instances), C.c_double(min_info_gain), C.c_uint64(seed))
    return Forest(h, Cc, T)


def confusion(pred, label, Cc):
    pred = np.ascontiguousarray(pred, np.float64); label = np.ascontiguousarray(label, np.float64)
    cm = np.zeros((Cc, Cc), np.int64)
    lib().orc_confusion(_p(pred), _p(label), C.c_int64(len(pred)), C.c_int32(Cc), _p(cm))
    return cm


def metrics(cm):
    cm = np.ascontiguousarray(cm, np.int64)
    out = np.zeros(5)
    lib().orc_metrics(_p(cm), C.c_int32(cm.shape[0]), _p(out))
    return dict(accuracy=out[0], weightedPrecision=out[1], weightedRecall=out[2], f1=out[3], macroF1=out[4])


def random_split(seed, n, cum_bounds, row_offset=0):
    cum = np.ascontiguousarray(cum_bounds, np.float64)
    out = np.empty(n, np.uint8)
    lib().orc_random_split(C.c_uint64(seed), C.c_int64(row_offset), C.c_int64(n), _p(cum), C.c_int32(len(cum)), _p(out))
    return out


# ---- host-side restatement of DecisionTreeMetadata.buildMetadata (A.1) and the fit driver ----
def check_shared_reciprocal_division(small_b=3000, n_random=100_000_000):
    """mismatch count of the shared-reciprocal exact division identity used by the CUDA split scorer (0 expected)."""
    f = lib().orc_check_shared_reciprocal_division
    f.restype = C.c_int64
    return int(f(C.c_int32(small_b), C.c_int64(n_random)))


def poisson_cdf_table(rate=1.0):
    """32 uint32 thresholds floor(CDF(k)·2^32) (saturating) for the bagging inverse-CDF (A.4)."""
    import math
    out = np.empty(32, np.uint32)
    term = math.exp(-rate); cdf = 0.0
    for k in range(32):
        cdf += term
        out[k] = min(int(math.floor(cdf * 4294967296.0)), 0xFFFFFFFF)
        term = term * rate / (k + 1)
    return out


def build_metadata(n_rows, F, num_classes, arity, max_bins, num_trees, strategy="auto"):
    import math
    arity = np.asarray(arity, np.int32)
    mpb = min(max_bins, n_rows)
    if arity.size and arity.max() > mpb:
        raise ValueError("DecisionTree requires maxBins (= %d) to be at least as large as the number of values "
                         "in each categorical feature, but categorical feature has %d values" % (mpb, arity.max()))
    kind = np.zeros(F, np.int32)
    if num_classes > 2:
        U = int(math.floor(math.log(mpb // 2 + 1) / math.log(2.0) + 1))
    else:
        U = 0
    for f in range(F):
        if arity[f] > 1:
            kind[f] = 2 if (num_classes > 2 and arity[f] <= U) else 1
        elif arity[f] == 1:
            kind[f] = 1
    if strategy == "auto":
        strategy = "all" if num_trees == 1 else "sqrt"
    if strategy == "all": m = F
    elif strategy == "sqrt": m = int(math.ceil(math.sqrt(F)))
    elif strategy == "log2": m = max(1, int(math.ceil(math.log(F) / math.log(2))))
    elif strategy == "onethird": m = int(math.ceil(F / 3.0))
    else:
        v = float(strategy)
        m = int(v) if v >= 1 and float(int(v)) == v and "." not in str(strategy) else int(math.ceil(v * F))
    return mpb, kind, max(1, min(m, F))


def fit_forest(x, y, num_classes, arity, num_trees=20, max_bins=32, max_depth=5, min_instances=1,
               min_info_gain=0.0, seed=0, strategy="auto", subsampling_rate=1.0, row_offset=0):
    """RandomForest.run restated end to end on a dense fp64 matrix (kdd99.py:64,79). Returns
    (Forest, dict(thresholds, n_thr, feat_bins, feat_kind, n_bins, m, tp))."""
    x = np.ascontiguousarray(x, np.float64)
    n, F = x.shape
    arity = np.asarray(arity, np.int32)
    mpb, kind, m = build_metadata(n, F, num_classes, arity, max_bins, num_trees, strategy)
    frac = min(1.0, max(mpb * mpb, 10000) / n) if (arity == 0).any() else 1.0
    keep = int(frac * 4294967296.0)
    thr, n_thr, _ = find_splits(x, seed, keep, arity, mpb, row_offset)
    tp, bad = bin_rows(x, thr, n_thr, arity, mpb, y)
    if bad:
        raise ValueError("categorical feature value out of range")
    feat_bins = np.where(arity > 0, arity, n_thr + 1).astype(np.int32)
    n_bins = int(feat_bins.max())
    cdf = poisson_cdf_table(subsampling_rate) if num_trees > 1 else None
    w = bag_weights(seed, num_trees, n, cdf, row_offset)
    fo = rf_train(tp, F, num_classes, w, feat_bins, kind, n_bins, m, max_depth, min_instances, min_info_gain, seed)
    return fo, dict(thresholds=thr, n_thr=n_thr, feat_bins=feat_bins, feat_kind=kind, n_bins=n_bins, m=m, tp=tp,
                    arity=arity, max_bins=mpb, w=w)
