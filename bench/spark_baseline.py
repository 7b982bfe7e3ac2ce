#!/usr/bin/env python
"""BASELINE.md B3 — the reference pipeline itself on Spark `local[*]`, timed on this box's host cores.

This is the REAL reference arm: it runs code/network_traffic_classifier_kdd99.py's RandomForest flow
(StringIndexer x4 -> VectorAssembler -> randomSplit 75/25 -> RandomForestClassifier -> transform -> evaluator;
kdd99.py:34-52,64,79-91) through genuine pyspark with master("local[*]") instead of the script's master("local")
(kdd99.py:10), on the synthetic KDD99-shaped CSV written by tools/make_synthetic_csv.py.

It needs a JVM and the pyspark package.  Neither exists in this image (no java, no pyspark, no
network to install them), so here it prints one line and exits 0:
    SKIPPED: no JVM/pyspark in image
bench.py's `--impl reference` therefore times the C++/OpenMP oracle port (oracle/) — stated as such in every bench line.
On a box that has Spark, run:  python tools/make_synthetic_csv.py kdd ... ; python bench/spark_baseline.py --csv <dir>/kddcup.data.corrected [--trees 100 --depth 16]
"""
import argparse
import importlib.util
import json
import os
import shutil
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "kddcup.data.corrected"),
                    help="headerless 42-column KDD99 file (tools/make_synthetic_csv.py kdd writes one)")
    ap.add_argument("--trees", type=int, default=20)
    ap.add_argument("--depth", type=int, default=5)
    ap.add_argument("--max-bins", type=int, default=70)
    a = ap.parse_args()
    if shutil.which("java") is None or importlib.util.find_spec("pyspark") is None:
        print("SKIPPED: no JVM/pyspark in image")
        return 0
    from pyspark.ml import Pipeline
    from pyspark.ml.classification import RandomForestClassifier
    from pyspark.ml.evaluation import MulticlassClassificationEvaluator
    from pyspark.ml.feature import StringIndexer, VectorAssembler
    from pyspark.sql import SparkSession
    spark = SparkSession.builder.master("local[*]").appName("b200flow-spark-baseline").getOrCreate()
    t_read = time.perf_counter()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spark-network-traffic-classifier_b200"))
    from b200flow.synth import KDD_COLUMNS                            # the 42 names of kdd99.py:15-23
    from pyspark.sql.functions import regexp_replace
    dataset = spark.read.csv(a.csv, header=False, inferSchema=True).toDF(*KDD_COLUMNS)       # kdd99.py:25
    dataset = dataset.withColumn("label", regexp_replace("label", "\\.", "")).cache()        # kdd99.py:27
    n = dataset.count()                                             # CSV parsing excluded from the timed region
    t0 = time.perf_counter()
    cats = ["protocol_type", "service", "flag"]
    indexers = [StringIndexer(inputCol=c, outputCol=c + "_num") for c in cats] + [StringIndexer(inputCol="label", outputCol="label_num")]
    dataset = Pipeline(stages=indexers).fit(dataset).transform(dataset)
    numerical = [c for c in dataset.columns if c not in cats + ["label", "label_num"]]
    dataset = VectorAssembler(inputCols=numerical, outputCol="features").transform(dataset).select(["features", "label_num"])
    train, test = dataset.randomSplit([0.75, 0.25], seed=2019)
    rf = RandomForestClassifier(labelCol="label_num", featuresCol="features", numTrees=a.trees, maxBins=a.max_bins, maxDepth=a.depth, seed=2019)
    pred = rf.fit(train).transform(test)
    f1 = MulticlassClassificationEvaluator(labelCol="label_num", predictionCol="prediction", metricName="f1").evaluate(pred)
    dt = time.perf_counter() - t0
    print(json.dumps({"impl": "spark-local[*]", "metric": "flow-records/sec fit+transform", "value": n / dt, "unit": "records/s", "rows": n,
                      "seconds": dt, "csv_parse_seconds": t0 - t_read, "cores": os.cpu_count(), "weighted_f1": f1,
                      "num_trees": a.trees, "max_depth": a.depth, "max_bins": a.max_bins}))
    spark.stop()
    return 0


if __name__ == "__main__":
    sys.exit(main())
