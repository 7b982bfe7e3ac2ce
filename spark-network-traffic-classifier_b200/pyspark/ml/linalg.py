"""Minimal pyspark.ml.linalg: DenseVector / Vectors for host-side inspection of vector columns."""
import numpy as np


class DenseVector:
    def __init__(self, values):
        self.values = np.asarray(values, np.float64)

    def toArray(self):
        return self.values

    def __len__(self):
        return len(self.values)

    def __getitem__(self, i):
        return float(self.values[i])

    def __eq__(self, o):
        return isinstance(o, DenseVector) and np.array_equal(self.values, o.values)

    def __repr__(self):
        return "DenseVector([%s])" % ", ".join(repr(float(v)) for v in self.values)


class Vectors:
    @staticmethod
    def dense(*v):
        return DenseVector(v[0] if len(v) == 1 and hasattr(v[0], "__len__") else v)
