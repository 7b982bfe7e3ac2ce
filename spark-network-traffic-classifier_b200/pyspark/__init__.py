"""pyspark.ml-shaped shim over b200flow (filled in below)."""
