"""The pyspark.ml-shaped surface the reference scripts use, end to end on the GPU:
same call sequence as code/network_traffic_classifier_{kdd99,cicids17}.py, on synthetic CSV files."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_kdd_csv(path, n, seed=1):
    from b200flow import synth
    rec, dicts = synth.make_kdd(n, 23, seed=seed, device="cuda")
    a = rec.cpu().numpy().view(synth.kdd_schema().numpy_dtype()).reshape(-1)
    with open(path, "w") as f:
        for r in a:
            cells = []
            for name in synth.KDD_COLUMNS:
                v = r[name]
                if name in dicts:
                    s = dicts[name][int(v)]
                    cells.append(s + "." if name == "label" else s)
                elif name in synth.KDD_RATE:
                    cells.append("%.2f" % v)
                else:
                    cells.append("%d" % int(v))
            f.write(",".join(cells) + "\n")
    return a, dicts


def test_kdd_script_flow(tmp_path, capsys):
    from pyspark.sql import SparkSession
    from pyspark.ml.feature import StringIndexer, VectorAssembler
    from pyspark.ml import Pipeline
    from pyspark.ml.classification import LogisticRegression, DecisionTreeClassifier, NaiveBayes, RandomForestClassifier
    from pyspark.ml.evaluation import MulticlassClassificationEvaluator
    from pyspark.sql.functions import regexp_replace
    from b200flow import synth

    csv = str(tmp_path / "kddcup.data.corrected")
    host, dicts = _write_kdd_csv(csv, 30000)
    spark = SparkSession.builder.appName("Network Attacks Classifier KDD99").master("local").getOrCreate()
    spark.sparkContext.setLogLevel("ERROR")
    dataset = spark.read.csv(csv, inferSchema=True, header=False)
    dataset = dataset.toDF(*synth.KDD_COLUMNS)
    dataset = dataset.withColumn("label", regexp_replace("label", r"\.", ""))
    assert dataset.count() == 30000 and len(dataset.columns) == 42
    cats = ["protocol_type", "service", "flag"]
    indexers = [StringIndexer(inputCol=c, outputCol=c + "_num") for c in cats]
    indexers.append(StringIndexer(inputCol="label", outputCol="label_num"))
    model = Pipeline(stages=indexers).fit(dataset)
    dataset = model.transform(dataset)
    # StringIndexer order = frequency desc
    lab_counts = {}
    for code in host["label"]:
        lab_counts[dicts["label"][code]] = lab_counts.get(dicts["label"][code], 0) + 1
    want = sorted(lab_counts, key=lambda k: (-lab_counts[k], k))
    assert model.stages[3].labels == want
    numerical = [c for c in dataset.columns if c not in cats + ["label", "label_num"]]
    assert numerical[-3:] == ["protocol_type_num", "service_num", "flag_num"] and len(numerical) == 41
    dataset = VectorAssembler(inputCols=numerical, outputCol="features").transform(dataset)
    attrs = dataset._cols["features"].meta["attrs"]
    assert [a.get("arity", 0) for a in attrs[-3:]] == [3, len(set(host["service"])), len(set(host["flag"]))]   # F7
    dataset = dataset.select(["features", "label_num"])
    train, test = dataset.randomSplit([0.75, 0.25], seed=2019)
    assert train.count() + test.count() == 30000 and abs(train.count() / 30000 - 0.75) < 0.02
    classifiers = {
        "Logistic Regression": LogisticRegression(maxIter=20, regParam=0.3, elasticNetParam=0.8, featuresCol="features",
                                                  labelCol="label_num", family="multinomial"),
        "Decision Tree": DecisionTreeClassifier(labelCol="label_num", featuresCol="features", maxBins=70),
        "Random Forest": RandomForestClassifier(labelCol="label_num", featuresCol="features", numTrees=20, maxBins=70),
        "Naive Bayes Multinomial": NaiveBayes(labelCol="label_num", featuresCol="features", smoothing=1.0, modelType="multinomial"),
    }
    scores = {}
    for name, clf in classifiers.items():
        m = clf.fit(train)
        pred = m.transform(test)
        pred.cache()
        ev = MulticlassClassificationEvaluator(labelCol="label_num", predictionCol="prediction")
        scores[name] = {}
        for metric in ["accuracy", "weightedPrecision", "weightedRecall", "f1"]:
            ev.setMetricName(metric)
            scores[name][metric] = ev.evaluate(pred)
            assert 0.0 <= scores[name][metric] <= 1.0
    assert scores["Random Forest"]["accuracy"] > 0.95 and scores["Decision Tree"]["accuracy"] > 0.95
    assert abs(scores["Random Forest"]["weightedRecall"] - scores["Random Forest"]["accuracy"]) < 1e-12
    assert spark.conf.get("spark.app.name") == "Network Attacks Classifier KDD99"
    # maxBins below the categorical arity must raise like MLlib (why the script sets maxBins=70)
    with pytest.raises(ValueError, match="maxBins"):
        DecisionTreeClassifier(labelCol="label_num", featuresCol="features").fit(train)


def test_shim_forest_equals_functional_api_and_oracle(tmp_path):
    """the shim adds nothing numerically: its RF equals b200flow.fit_forest on the same matrix, which the
    parity tests pin to the oracle."""
    import oracle
    from pyspark.sql import SparkSession
    from pyspark.ml.feature import StringIndexer, VectorAssembler
    from pyspark.ml.classification import RandomForestClassifier
    csv = str(tmp_path / "k.csv")
    _write_kdd_csv(csv, 8000, seed=4)
    from b200flow import synth
    spark = SparkSession.builder.getOrCreate()
    df = spark.read.csv(csv, inferSchema=True, header=False).toDF(*synth.KDD_COLUMNS)
    for c in ["protocol_type", "service", "flag", "label"]:
        df = StringIndexer(inputCol=c, outputCol=c + "_num").fit(df).transform(df)
    cols = [c for c in df.columns if c not in ["protocol_type", "service", "flag", "label", "label_num"]]
    df = VectorAssembler(inputCols=cols, outputCol="features").transform(df)
    rf = RandomForestClassifier(labelCol="label_num", featuresCol="features", numTrees=5, maxBins=70, maxDepth=6, seed=11)
    m = rf.fit(df)
    x = df._cols["features"].data.cpu().numpy(); y = df._cols["label_num"].data.cpu().numpy().astype(np.int32)
    arity = [int(a.get("arity", 0)) for a in df._cols["features"].meta["attrs"]]
    C = len(df._cols["label_num"].meta["ml_attr"]["vals"])
    fo, meta = oracle.fit_forest(x, y, C, arity, num_trees=5, max_bins=70, max_depth=6, seed=11)
    from util import forests_equal
    assert forests_equal(m._forest.export(), fo.export()) == []
    out = m.transform(df)
    tp, _ = oracle.bin_rows(x, meta["thresholds"], meta["n_thr"], meta["arity"], meta["max_bins"])
    raw, prob, pred = fo.predict(tp)
    assert np.array_equal(out._cols["prediction"].data.cpu().numpy(), pred)
    assert np.array_equal(out._cols["probability"].data.cpu().numpy(), prob)
    assert m.getNumTrees == 5 and m.numClasses == C and abs(m.featureImportances.toArray().sum() - 1.0) < 1e-9
    assert "Tree 0" in m.toDebugString


def test_cicids_script_flow(tmp_path):
    from pyspark.sql import SparkSession
    from pyspark.ml.feature import StringIndexer, VectorAssembler
    from pyspark.ml.classification import RandomForestClassifier, DecisionTreeClassifier
    from pyspark.ml.evaluation import MulticlassClassificationEvaluator
    from pyspark.sql.functions import regexp_replace, col
    from b200flow import synth
    rec, dicts = synth.make_cicids(12000, 6, seed=3, device="cuda", nan_fraction=0.02, n_features=20)
    a = rec.cpu().numpy().view(synth.cicids_schema(20).numpy_dtype()).reshape(-1)
    names = [" Flow Duration", " Init_Win_bytes_forward", "Fwd Header Length"] + [" Feat %d" % i for i in range(3, 19)] + \
            ["Fwd Header Length", " Label"]
    for part in range(2):                                               # two files -> glob
        with open(tmp_path / ("day%d.pcap_ISCX.csv" % part), "w", encoding="utf-8") as f:
            f.write(",".join(names) + "\n")
            for r in a[part::2]:
                lab = dicts["Label"][int(r["Label"])]
                if lab == "DoS Hulk":
                    lab = "DoS � Hulk"
                cells = ["NaN" if np.isnan(r["f%02d" % i]) else repr(float(r["f%02d" % i])) for i in range(20)]
                f.write(",".join(cells) + "," + lab + "\n")
    spark = SparkSession.builder.appName("cic").master("local").getOrCreate()
    ds = spark.read.csv(str(tmp_path / "*.pcap_ISCX.csv"), inferSchema=True, header=True, multiLine=True,
                        ignoreLeadingWhiteSpace=True, ignoreTrailingWhiteSpace=True)
    ds.printSchema()
    assert ds.count() == 12000 and ds.columns[0] == "Flow Duration" and ds.columns[2] == "Fwd Header Length2" \
        and ds.columns[19] == "Fwd Header Length19"
    ds = ds.withColumn("Label", regexp_replace("Label", u"� ", ""))
    ds.select("Label").groupBy("Label").count().orderBy("count", ascending=False).show()
    n0 = ds.count()
    ds = ds.where(col("Flow Duration") > 20).where(col("Init_Win_bytes_forward") > 1000.0)
    keep = (a["f00"] > 20) & (a["f01"] > 1000.0)
    assert ds.count() == int(keep.sum()) < n0
    feats = [f for f in ds.columns if f not in ["Label"]]
    ds = VectorAssembler(inputCols=feats, outputCol="features").setHandleInvalid("skip").transform(ds)
    nan_rows = np.isnan(np.stack([a["f%02d" % i] for i in range(20)], 1)).any(1)
    assert ds.count() == int((keep & ~nan_rows).sum())
    li = StringIndexer(inputCol="Label", outputCol="Label_Idx").setHandleInvalid("skip").fit(ds)
    ds = li.transform(ds)
    label_list = ds.select(["Label", "Label_Idx"]).distinct().orderBy("Label_Idx").select("Label").rdd.flatMap(lambda x: x).collect()
    assert label_list == li.labels and "DoS Hulk" in label_list
    ds = ds.select(["features", "Label_Idx"])
    train, test = ds.randomSplit([0.75, 0.25], seed=2019)
    for clf in (DecisionTreeClassifier(labelCol="Label_Idx", featuresCol="features", maxBins=len(feats)),
                RandomForestClassifier(labelCol="Label_Idx", featuresCol="features", numTrees=20, maxBins=len(feats))):
        pred = clf.fit(train).transform(test)
        ev = MulticlassClassificationEvaluator(labelCol="Label_Idx", predictionCol="prediction")
        assert ev.setMetricName("accuracy").evaluate(pred) > 0.8
        t = pred.select("Label_Idx").rdd.flatMap(lambda x: x).collect()
        p = pred.select("prediction").rdd.flatMap(lambda x: x).collect()
        nums = pred.select("Label_Idx").distinct().orderBy("Label_Idx").rdd.flatMap(lambda x: x).collect()
        assert len(t) == len(p) == test.count() and nums == sorted(set(t))


def test_onehot_standardscaler_pipeline_is_fused_and_matches_oracle(tmp_path):
    import oracle
    from pyspark.sql import SparkSession
    from pyspark.ml import Pipeline
    from pyspark.ml.feature import StringIndexer, OneHotEncoder, VectorAssembler, StandardScaler
    from b200flow import synth
    csv = str(tmp_path / "k.csv")
    _write_kdd_csv(csv, 6000, seed=9)
    spark = SparkSession.builder.getOrCreate()
    df = spark.read.csv(csv, inferSchema=True, header=False).toDF(*synth.KDD_COLUMNS)
    cats = ["protocol_type", "service", "flag"]
    stages = [StringIndexer(inputCol=c, outputCol=c + "_num") for c in cats]
    stages.append(OneHotEncoder(inputCols=[c + "_num" for c in cats], outputCols=[c + "_oh" for c in cats]))
    nums = [c for c in synth.KDD_COLUMNS if c not in cats + ["label"]]
    stages.append(VectorAssembler(inputCols=nums + [c + "_oh" for c in cats], outputCol="raw_features"))
    stages.append(StandardScaler(inputCol="raw_features", outputCol="features", withMean=True, withStd=True))
    model = Pipeline(stages=stages).fit(df)
    out = model.transform(df)
    assert out._cols["features"].prov is not None and out._cols["features"].prov[0] == "plan"     # fused from raw records
    raw = out._cols["raw_features"].data.cpu().numpy()
    mean, std = oracle.moments(raw)
    want = (raw - mean) * np.where(std != 0, 1.0 / np.where(std != 0, std, 1.0), 0.0)
    got = out._cols["features"].data.cpu().numpy()
    assert got.shape[1] == 38 + sum(len(s.labels) - 1 for s in model.stages[:3])
    assert np.allclose(got, want, rtol=1e-6, atol=1e-9)
    # doctest known answers through the shim
    d2 = spark.createDataFrame([(0.0,), (2.0,)], ["a"])
    d2 = VectorAssembler(inputCols=["a"], outputCol="v").transform(d2)
    sm = StandardScaler(inputCol="v", outputCol="s").fit(d2)
    assert sm.mean.tolist() == [1.0] and abs(sm.std[0] - 1.4142135623730951) < 1e-15
    assert np.allclose(sm.transform(d2)._cols["s"].data.cpu().numpy()[:, 0], [0.0, 1.4142135623730951])
    d3 = spark.createDataFrame([("a",), ("b",), ("c",), ("a",), ("a",), ("c",)], ["x"])
    d3 = StringIndexer(inputCol="x", outputCol="i").fit(d3).transform(d3)
    assert d3._cols["i"].data.cpu().numpy().tolist() == [0.0, 2.0, 1.0, 0.0, 0.0, 1.0]


def test_indexer_columns_are_lazy_and_fused_but_identical():
    """StringIndexerModel.transform defers its kernel when the record buffer's own category counts prove that no label is
    unseen; VectorAssembler fuses the lookup.  The deferred column, once read, and the fused vector hold the same values as
    the eager path; a model applied to OTHER data still checks for unseen labels."""
    from b200flow import synth
    from pyspark.ml import Pipeline
    from pyspark.ml.feature import SparkException, StringIndexer, VectorAssembler
    from pyspark.sql import DataFrame
    rec, dicts = synth.make_kdd(30011, 23, seed=8, device="cuda")
    df = DataFrame.fromRecords(rec, synth.kdd_schema(), dicts)
    cats = synth.KDD_CATEGORICAL
    stages = [StringIndexer(inputCol=c, outputCol=c + "_num") for c in cats + ["label"]]
    model = Pipeline(stages=stages).fit(df)
    out = model.transform(df)
    assert all(out._cols[c + "_num"]._data is None for c in cats + ["label"])          # nothing launched yet
    numerical = [c for c in out.columns if c not in cats + ["label", "label_num"]]
    feats = VectorAssembler(inputCols=numerical, outputCol="features").transform(out)._cols["features"].data
    raw = rec.view(torch.int32)
    schema = synth.kdd_schema()
    for j, c in enumerate(cats):
        lazy = out._cols[c + "_num"].data                                             # materialised on first read
        rank_of = {s: i for i, s in enumerate(model.stages[j].labels)}
        lut = torch.tensor([rank_of.get(s, -1) for s in dicts[c]], dtype=torch.float64, device="cuda")
        want = lut[raw[:, schema.offsets[c] // 4].long()]
        assert torch.equal(lazy, want) and torch.equal(feats[:, 38 + j].to(torch.float64), want)
    # other data with a label the model has not seen: the eager, checking path runs and raises
    few = rec[rec.view(torch.int32)[:, schema.offsets["service"] // 4] == 0]
    m_few = StringIndexer(inputCol="service", outputCol="s_num").fit(DataFrame.fromRecords(few.contiguous(), schema, dicts))
    with pytest.raises(SparkException):
        m_few.transform(df)


def test_lazy_assembled_vector_feeds_the_trees_from_raw_records():
    """VectorAssembler over raw record fields defers its kernel (like Spark's lazy transform); select / randomSplit move the
    RECORDS only, RandomForestClassifier.fit and model.transform bin straight from them (fused encode -> bins), and the result
    equals the eager path (vector materialised before the split) exactly.  A NaN under handleInvalid="error" raises at the action."""
    from b200flow import synth
    from pyspark.ml import Pipeline
    from pyspark.ml.classification import RandomForestClassifier
    from pyspark.ml.evaluation import MulticlassClassificationEvaluator
    from pyspark.ml.feature import SparkException, StringIndexer, VectorAssembler
    from pyspark.sql import DataFrame
    rec, dicts = synth.make_kdd(40000, 5, seed=21, device="cuda")
    schema = synth.kdd_schema()
    cats = synth.KDD_CATEGORICAL

    def flow(materialise, records=rec):
        df = DataFrame.fromRecords(records, schema, dicts)
        stages = [StringIndexer(inputCol=c, outputCol=c + "_num") for c in cats] + [StringIndexer(inputCol="label", outputCol="label_num")]
        df = Pipeline(stages=stages).fit(df).transform(df)
        numerical = [c for c in df.columns if c not in cats + ["label", "label_num"]]
        df = VectorAssembler(inputCols=numerical, outputCol="features").transform(df).select(["features", "label_num"])
        assert df._cols["features"].lazy
        if materialise:
            df._cols["features"].data                                               # the fused encode kernel runs now
            assert not df._cols["features"].lazy
        train, test = df.randomSplit([0.75, 0.25], seed=2019)
        assert train._cols["features"].lazy == (not materialise)
        model = RandomForestClassifier(labelCol="label_num", featuresCol="features", numTrees=6, maxBins=70, maxDepth=7, seed=5).fit(train)
        pred = model.transform(test)
        assert pred._cols["features"].lazy == (not materialise)                    # the dense matrix never existed on the lazy path
        f1 = MulticlassClassificationEvaluator(labelCol="label_num", predictionCol="prediction", metricName="f1").evaluate(pred)
        return model._forest.export(), pred._cols["prediction"].data, pred._cols["probability"].data, f1, test

    ex_l, pred_l, prob_l, f1_l, test_l = flow(False)
    ex_e, pred_e, prob_e, f1_e, _ = flow(True)
    assert all(np.array_equal(ex_l[k], ex_e[k]) for k in ex_e) and torch.equal(pred_l, pred_e) and torch.equal(prob_l, prob_e) and f1_l == f1_e
    # reading the lazy column after the fact gives the values the eager path assembled
    assert test_l._cols["features"].data.shape == (test_l.count(), 41)
    bad = rec.clone(); bad.view(torch.float32)[123, 0] = float("nan")
    with pytest.raises(SparkException):
        flow(False, bad)
