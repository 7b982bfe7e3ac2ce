"""pyspark-shaped shim: the subset of pyspark.sql / pyspark.ml that
/root/reference/code/network_traffic_classifier_{kdd99,cicids17}.py import, backed by
the b200flow CUDA library.  Put `spark-network-traffic-classifier_b200/` on PYTHONPATH and the scripts
resolve `from pyspark...` to this package.  This is NOT Apache Spark: only the operators on the hot
path are implemented natively; relational plumbing is a thin columnar layer over torch CUDA tensors.
"""
__version__ = "2.4.0+b200flow"
