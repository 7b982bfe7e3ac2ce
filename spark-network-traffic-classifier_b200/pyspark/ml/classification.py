"""pyspark.ml.classification shim.

RandomForestClassifier / DecisionTreeClassifier (kdd99.py:61,64; cicids17.py:65,68) are the hot path and run
entirely on the b200flow CUDA kernels (histogram build, Gini split scoring, batch predict).
LogisticRegression and NaiveBayes (kdd99.py:57,67; cicids17.py:61,71) are OUT of the kernel scope (SURVEY.md
§8f rank 4): torch fp64 implementations of MLlib's statistics / objective (b200flow/linear.py), checked against a numpy
restatement and scikit-learn in tests/test_linear_models.py.
"""
import zlib

import numpy as np
import torch

from b200flow import dist as bdist
from b200flow import forest as fr
from b200flow import linear as _linear

from . import Estimator, Model
from ..sql import ColumnData
from .feature import IllegalArgumentException, _materialize


def _default_seed(obj):
    # pyspark uses hash(type(self).__name__), which Python 3 randomises per process; a stable hash is used instead
    return zlib.crc32(type(obj).__name__.encode())


def _features_and_labels(df, est):
    fcol, lcol = est.getOrDefault("featuresCol"), est.getOrDefault("labelCol")
    for c in (fcol, lcol):
        if c not in df._cols:
            raise IllegalArgumentException("Field \"%s\" does not exist." % c)
    fc = df._cols[fcol]
    if fc.kind != "vector":
        raise IllegalArgumentException("Column %s must be of type vector" % fcol)
    x = fc.data
    y = df._column_tensor(lcol)
    lab_meta = df._cols[lcol].meta.get("ml_attr", {})
    if lab_meta.get("type") == "nominal":
        num_classes = len(lab_meta["vals"])
    else:
        if bool(((y < 0) | (y != torch.floor(y))).any().item()):
            raise IllegalArgumentException("Classifier was given dataset with invalid label. Labels must be integers in [0, numClasses).")
        num_classes = int(y.max().item()) + 1 if y.numel() else 0
    return x, y, num_classes, fc.meta.get("attrs")


def _arity_from_attrs(attrs, F):
    if not attrs or len(attrs) != F:
        return [0] * F
    return [int(a.get("arity", 0)) if a.get("type") in ("nominal", "binary") else 0 for a in attrs]


def _lazy_plan(df, fcol):
    """(encode plan, emulate-f32 flag) when the features column is still lazy — VectorAssembler over raw record fields whose
    kernel has not run — so that trees can be trained / applied straight from the records (fused encode -> bins); else None."""
    fc = df._cols.get(fcol)
    if fc is None or fc.kind != "vector" or not fc.lazy or fc.prov is None or fc.prov[0] != "plan" or df._rec is None:
        return None
    return fc.prov[1]


def _records_fit_inputs(df, est):
    """-> (records, plan with the label column set, num_classes, per-slot attrs) for the fused record path, or None."""
    import copy
    fcol, lcol = est.getOrDefault("featuresCol"), est.getOrDefault("labelCol")
    plan = _lazy_plan(df, fcol)
    lc = df._cols.get(lcol)
    if plan is None or lc is None or lc.prov is None or lc.prov[0] != "index" or not lc.lazy:
        return None
    lab_meta = lc.meta.get("ml_attr", {})
    if lab_meta.get("type") != "nominal":
        return None
    p2 = copy.copy(plan); p2.slots = list(plan.slots); p2.luts = list(plan.luts); p2._dev = None
    p2.set_label(lc.prov[1], lc.prov[2])
    return df._rec, p2, len(lab_meta["vals"]), df._cols[fcol].meta.get("attrs")


class _TreeParams:
    _defaults = {"featuresCol": "features", "labelCol": "label", "predictionCol": "prediction",
                 "probabilityCol": "probability", "rawPredictionCol": "rawPrediction", "maxDepth": 5, "maxBins": 32,
                 "minInstancesPerNode": 1, "minInfoGain": 0.0, "maxMemoryInMB": 256, "cacheNodeIds": False,
                 "checkpointInterval": 10, "impurity": "gini", "seed": None}


class _TreeClassifierBase(Estimator, _TreeParams):
    def _forest_params(self, num_trees, strategy, subsampling, bootstrap):
        imp = str(self.getOrDefault("impurity")).lower()
        if imp not in ("gini", "entropy"):
            raise IllegalArgumentException("impurity must be gini or entropy, got %r" % imp)
        if imp == "entropy":       # log is not bit-reproducible between libm and CUDA: not offered on the B200 path
            raise IllegalArgumentException("impurity='entropy' is not supported by the b200flow tree trainer; use 'gini' "
                                           "(the reference scripts take the default, kdd99.py:64 / cicids17.py:68)")
        seed = self.getOrDefault("seed")
        return fr.ForestParams(num_trees=int(num_trees), max_depth=int(self.getOrDefault("maxDepth")),
                               max_bins=int(self.getOrDefault("maxBins")),
                               min_instances_per_node=int(self.getOrDefault("minInstancesPerNode")),
                               min_info_gain=float(self.getOrDefault("minInfoGain")), feature_subset_strategy=str(strategy),
                               subsampling_rate=float(subsampling), impurity=imp,
                               seed=_default_seed(self) if seed is None else int(seed), bootstrap=bootstrap)

    def _train(self, df, params):
        from .feature import SparkException
        fused = _records_fit_inputs(df, self)
        try:
            grp = bdist.group()
            if fused is not None:                           # lazy VectorAssembler output: bin straight from the raw records
                rec, plan, C, attrs = fused
                if C > 100:
                    raise IllegalArgumentException("Classifier inferred %d classes; maximum is 100" % C)
                off, _ = bdist.global_offset(rec.shape[0], rec.device, grp)
                return fr.fit_forest_records(rec, plan, C, _arity_from_attrs(attrs, plan.n_out), params, row_offset=off, group=grp)
            x, y, C, attrs = _features_and_labels(df, self)
            if C > 100:
                raise IllegalArgumentException("Classifier inferred %d classes; maximum is 100" % C)
            off, _ = bdist.global_offset(x.shape[0], x.device, grp)
            forest = fr.fit_forest(x, y.to(torch.int32), C, _arity_from_attrs(attrs, x.shape[1]), params,
                                   row_offset=off, group=grp)
        except fr.InvalidRowsError as e:                    # NaN / null under handleInvalid="error": surfaces at the action, as in Spark
            raise SparkException("Encountered NaN/null while assembling a row with handleInvalid = \"error\" (%s)" % e)
        except ValueError as e:        # includes b200flow's UnsupportedParamError; CUDA failures propagate as they are
            raise IllegalArgumentException(str(e))
        return forest


class RandomForestClassifier(_TreeClassifierBase):
    _defaults = {"numTrees": 20, "featureSubsetStrategy": "auto", "subsamplingRate": 1.0}

    def __init__(self, featuresCol=None, labelCol=None, predictionCol=None, probabilityCol=None, rawPredictionCol=None,
                 maxDepth=None, maxBins=None, minInstancesPerNode=None, minInfoGain=None, maxMemoryInMB=None,
                 cacheNodeIds=None, checkpointInterval=None, impurity=None, numTrees=None, featureSubsetStrategy=None,
                 seed=None, subsamplingRate=None):
        kw = dict(locals()); kw.pop("self"); kw.pop("__class__", None)
        super().__init__(**kw)

    def _fit(self, df):
        p = self._forest_params(self.getOrDefault("numTrees"), self.getOrDefault("featureSubsetStrategy"),
                                self.getOrDefault("subsamplingRate"), bootstrap=True)
        m = RandomForestClassificationModel(self._train(df, p))
        m._paramMap = {k: v for k, v in self._paramMap.items() if k in m._all_defaults()}
        return m


class DecisionTreeClassifier(_TreeClassifierBase):
    """MLlib trains it as RandomForest.run(numTrees=1, featureSubsetStrategy='all'), no bagging."""

    def __init__(self, featuresCol=None, labelCol=None, predictionCol=None, probabilityCol=None, rawPredictionCol=None,
                 maxDepth=None, maxBins=None, minInstancesPerNode=None, minInfoGain=None, maxMemoryInMB=None,
                 cacheNodeIds=None, checkpointInterval=None, impurity=None, seed=None):
        kw = dict(locals()); kw.pop("self"); kw.pop("__class__", None)
        super().__init__(**kw)

    def _fit(self, df):
        p = self._forest_params(1, "all", 1.0, bootstrap=False)
        m = DecisionTreeClassificationModel(self._train(df, p))
        m._paramMap = {k: v for k, v in self._paramMap.items() if k in m._all_defaults()}
        return m


class _ForestModelBase(Model, _TreeParams):
    def __init__(self, forest):
        super().__init__()
        self._forest = forest

    @property
    def numClasses(self):
        return self._forest.C

    @property
    def numFeatures(self):
        return self._forest.F

    @property
    def totalNumNodes(self):
        return self._forest.n_nodes

    @property
    def featureImportances(self):
        from .linalg import DenseVector
        return DenseVector(self._forest.feature_importances())

    def _transform(self, df):
        fcol = self.getOrDefault("featuresCol")
        if fcol not in df._cols or df._cols[fcol].kind != "vector":
            raise IllegalArgumentException("Column %s must be of type vector" % fcol)
        plan = _lazy_plan(df, fcol)
        if plan is not None and plan.n_out == self._forest.F:   # lazy features: fused encode -> bins -> tree walk, no dense matrix
            from .feature import SparkException
            try:
                raw, prob, pred, _ = self._forest.predict_records(df._rec, plan, on_invalid="error" if plan.check_nan else "ignore")
            except fr.InvalidRowsError as e:
                raise SparkException("Encountered NaN/null while assembling a row with handleInvalid = \"error\" (%s)" % e)
        else:
            x = df._cols[fcol].data
            raw, prob, pred = self._forest.predict(x)          # R9: bin + walk all trees on the GPU
        cols = dict(df._cols)
        for name, key, kind in ((self.getOrDefault("rawPredictionCol"), raw, "vector"),
                                (self.getOrDefault("probabilityCol"), prob, "vector"),
                                (self.getOrDefault("predictionCol"), pred, "numeric")):
            if name:
                if name in cols:
                    raise IllegalArgumentException("Output column %s already exists." % name)
                cols[name] = ColumnData(kind, key, "f64")
        return df._with(cols=cols)

    def _tree_strings(self):
        ex = self._forest.export()
        thr = self._forest.thresholds.cpu().numpy()
        out = []
        for t in range(self._forest.T):
            sel = np.nonzero(ex["tree"] == t)[0]
            idx = {int(ex["nid"][i]): i for i in sel}
            lines = []

            def rec(nid, depth):
                i = idx[nid]
                pad = " " * (depth + 1)
                if ex["is_leaf"][i]:
                    lines.append("%sPredict: %.1f" % (pad, float(np.argmax(ex["counts"][i]))))
                    return
                f = int(ex["feat"][i])
                if ex["kind"][i] == 0:
                    v = thr[f, int(ex["bin_thr"][i])]
                    lines.append("%sIf (feature %d <= %s)" % (pad, f, repr(float(v)))); rec(nid * 2, depth + 1)
                    lines.append("%sElse (feature %d > %s)" % (pad, f, repr(float(v)))); rec(nid * 2 + 1, depth + 1)
                else:
                    cats = [c for c in range(256) if (int(ex["mask"][i][c >> 6]) >> (c & 63)) & 1]
                    s = "{%s}" % ",".join("%.1f" % c for c in cats)
                    lines.append("%sIf (feature %d in %s)" % (pad, f, s)); rec(nid * 2, depth + 1)
                    lines.append("%sElse (feature %d not in %s)" % (pad, f, s)); rec(nid * 2 + 1, depth + 1)
            rec(1, 0)
            out.append((len(sel), lines))
        return out


class RandomForestClassificationModel(_ForestModelBase):
    _defaults = {"numTrees": 20, "featureSubsetStrategy": "auto", "subsamplingRate": 1.0}

    @property
    def getNumTrees(self):
        return self._forest.T

    @property
    def treeWeights(self):
        return [1.0] * self._forest.T

    @property
    def toDebugString(self):
        parts = ["RandomForestClassificationModel with %d trees" % self._forest.T]
        for t, (nn, lines) in enumerate(self._tree_strings()):
            parts.append("  Tree %d (weight 1.0):" % t)
            parts += ["  " + l for l in lines]
        return "\n".join(parts) + "\n"

    def __repr__(self):
        return "RandomForestClassificationModel with %d trees" % self._forest.T


class DecisionTreeClassificationModel(_ForestModelBase):
    @property
    def numNodes(self):
        return self._forest.n_nodes

    @property
    def depth(self):
        nid = self._forest.export()["nid"].astype(np.int64)
        return int(np.floor(np.log2(nid.max()))) if len(nid) else 0

    @property
    def toDebugString(self):
        nn, lines = self._tree_strings()[0]
        return "DecisionTreeClassificationModel of depth %d with %d nodes\n%s\n" % (self.depth, nn, "\n".join(lines))

    def __repr__(self):
        return "DecisionTreeClassificationModel of depth %d with %d nodes" % (self.depth, self.numNodes)


# ------------------------------------------------------------------------------- out-of-scope models (torch)
class _ProbModel(Model):
    _defaults = {"featuresCol": "features", "labelCol": "label", "predictionCol": "prediction",
                 "probabilityCol": "probability", "rawPredictionCol": "rawPrediction"}

    def _emit(self, df, raw, prob):
        pred = torch.argmax(raw, 1).to(torch.float64)
        cols = dict(df._cols)
        cols[self.getOrDefault("rawPredictionCol")] = ColumnData("vector", raw, "f64")
        cols[self.getOrDefault("probabilityCol")] = ColumnData("vector", prob, "f64")
        cols[self.getOrDefault("predictionCol")] = ColumnData("numeric", pred, "f64")
        return df._with(cols=cols)


class LogisticRegression(Estimator):
    """Multinomial / binomial elastic-net logistic regression with MLlib's objective (standardised features, unpenalised
    intercepts), minimised by OWL-QN: b200flow/linear.py (kdd99.py:57-58, cicids17.py:61-62)."""
    _defaults = dict(_ProbModel._defaults, maxIter=100, regParam=0.0, elasticNetParam=0.0, tol=1e-6, fitIntercept=True,
                     standardization=True, family="auto", threshold=0.5)

    def __init__(self, featuresCol=None, labelCol=None, predictionCol=None, maxIter=None, regParam=None,
                 elasticNetParam=None, tol=None, fitIntercept=None, threshold=None, probabilityCol=None,
                 rawPredictionCol=None, standardization=None, family=None):
        kw = dict(locals()); kw.pop("self"); kw.pop("__class__", None)
        super().__init__(**kw)

    def _fit(self, df):
        x, y, C, _ = _features_and_labels(df, self)
        g = self.getOrDefault
        if g("family") not in ("auto", "binomial", "multinomial"):
            raise IllegalArgumentException("family must be auto, binomial or multinomial")
        try:
            fit = _linear.lr_fit(x, y, C, max_iter=int(g("maxIter")), reg_param=float(g("regParam")), elastic_net=float(g("elasticNetParam")),
                                 tol=float(g("tol")), fit_intercept=bool(g("fitIntercept")), standardization=bool(g("standardization")),
                                 family=g("family"), group=bdist.group())
        except ValueError as e:
            raise IllegalArgumentException(str(e))
        m = LogisticRegressionModel(fit, C)
        m._paramMap = {k: v for k, v in self._paramMap.items() if k in m._all_defaults()}
        return m


class _TrainingSummary:
    def __init__(self, hist, iterations):
        self.objectiveHistory, self.totalIterations = list(hist), int(iterations)


class LogisticRegressionModel(_ProbModel):
    def __init__(self, fit, C):
        super().__init__()
        self._fit_result, self.numClasses = fit, C
        self.summary = _TrainingSummary(fit.objective_history, fit.iterations)

    @property
    def coefficientMatrix(self):
        return self._fit_result.coef.cpu().numpy()

    @property
    def interceptVector(self):
        return self._fit_result.intercept.cpu().numpy()

    def _transform(self, df):
        raw = _linear.lr_raw(self._fit_result, df._cols[self.getOrDefault("featuresCol")].data)
        return self._emit(df, raw, _linear.lr_probability(self._fit_result, raw))


class NaiveBayes(Estimator):
    """Multinomial naive Bayes with MLlib's smoothed prior and likelihood (b200flow/linear.py; kdd99.py:67, cicids17.py:71);
    rejects negative features as MLlib does (the reason for the where-filters at cicids17.py:30-35)."""
    _defaults = dict(_ProbModel._defaults, smoothing=1.0, modelType="multinomial")

    def __init__(self, featuresCol=None, labelCol=None, predictionCol=None, probabilityCol=None, rawPredictionCol=None,
                 smoothing=None, modelType=None):
        kw = dict(locals()); kw.pop("self"); kw.pop("__class__", None)
        super().__init__(**kw)

    def _fit(self, df):
        if self.getOrDefault("modelType") != "multinomial":
            raise IllegalArgumentException("only modelType='multinomial' is implemented")
        x, y, C, _ = _features_and_labels(df, self)
        try:
            fit = _linear.nb_fit(x, y, C, float(self.getOrDefault("smoothing")), group=bdist.group())
        except ValueError as e:
            raise IllegalArgumentException(str(e))
        m = NaiveBayesModel(fit, C)
        m._paramMap = {k: v for k, v in self._paramMap.items() if k in m._all_defaults()}
        return m


class NaiveBayesModel(_ProbModel):
    def __init__(self, fit, C):
        super().__init__()
        self._fit_result, self.numClasses = fit, C

    @property
    def pi(self):
        return self._fit_result.pi.cpu().numpy()

    @property
    def theta(self):
        return self._fit_result.theta.cpu().numpy()

    def _transform(self, df):
        raw = _linear.nb_raw(self._fit_result, df._cols[self.getOrDefault("featuresCol")].data)
        return self._emit(df, raw, torch.softmax(raw, 1))
