"""CPU restatement (pure Python) of `spark.read.csv(path, inferSchema=..., header=...)` for unquoted CSV text — TEST
INFRASTRUCTURE ONLY (only tests/ may import it); the checker of the device reader csrc/csv.cu.  PARITY UNPINNED at the Spark
boundary (no JVM here); follows [recalled] Spark 2.4 sql/execution/datasources/csv/{CSVInferSchema,UnivocityParser}.scala as the
scripts call it (kdd99.py:25; cicids17.py:19-20): per field Integer -> Long -> Double -> String with Java's parseInt /
parseDouble grammar (Python's float() is correctly rounded like Double.parseDouble), empty field = null, blank lines skipped,
every file of a glob carries its own header line, duplicate header names get their position appended.
The record layout choices of this repo (no int64 / nullable-int fields: such columns become float64 with NaN for null; strings
become dictionary codes in order of first appearance, null = -1) are restated too, so that outputs compare byte for byte."""
import re

import numpy as np

_WS = bytes(range(0x21))
_INT = re.compile(rb"^[+-]?[0-9]+$")
_DBL = re.compile(rb"^[+-]?([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?$")
_SPECIAL = {b"NaN": float("nan"), b"Infinity": float("inf"), b"+Infinity": float("inf"), b"-Infinity": float("-inf"),
            b"Inf": float("inf"), b"+Inf": float("inf"), b"-Inf": float("-inf")}
NULL, INT, LONG, DOUBLE, STRING = range(5)


def classify(f):
    if f == b"":
        return NULL
    if _INT.match(f):
        v = int(f)
        if -2 ** 31 <= v < 2 ** 31:
            return INT
        if -2 ** 63 <= v < 2 ** 63:
            return LONG
        return DOUBLE
    t = f.strip(_WS)
    if _DBL.match(t) or t in _SPECIAL:
        return DOUBLE
    return STRING


def to_double(f):
    if f == b"":
        return float("nan")
    t = f.strip(_WS)
    return _SPECIAL[t] if t in _SPECIAL else float(t)


def _lines(data):
    for ln in data.split(b"\n"):
        if ln.endswith(b"\r"):
            ln = ln[:-1]
        if ln:
            yield ln


def read_csv(paths, header=False, infer_schema=False, strip_lead=False, strip_trail=False):
    """-> (names, types ['i32' | 'f64' | 'code'], {name: numpy column}, {name: [strings]})"""
    rows, names = [], None
    for p in paths:
        it = _lines(open(p, "rb").read())
        if header:
            h = next(it, None)
            if h is not None and names is None:
                names = [c.decode("utf-8", "replace") for c in h.split(b",")]
                names = [c.strip() if (strip_lead or strip_trail) else c for c in names]
                low = [n.lower() for n in names]
                names = [n + str(i) if low.count(n.lower()) > 1 else n for i, n in enumerate(names)]
        for ln in it:
            fs = ln.split(b",")
            if strip_lead:
                fs = [f.lstrip(_WS) for f in fs]
            if strip_trail:
                fs = [f.rstrip(_WS) for f in fs]
            rows.append(fs)
    if names is None:
        names = ["_c%d" % i for i in range(len(rows[0]) if rows else 0)]
    for r in rows:
        if len(r) != len(names):
            raise ValueError("ragged row")
    types, cols, dicts = [], {}, {}
    for c, name in enumerate(names):
        fields = [r[c] for r in rows]
        classes = [classify(f) for f in fields] if infer_schema else [STRING] * len(fields)
        k = max(classes, default=NULL)
        has_null = any(f == b"" for f in fields)
        if k == INT and not has_null:
            types.append("i32"); cols[name] = np.array([int(f) for f in fields], np.int32)
        elif k in (INT, LONG, DOUBLE) and infer_schema:
            types.append("f64"); cols[name] = np.array([to_double(f) for f in fields], np.float64)
        else:
            values, code_of, codes = [], {}, []
            for f in fields:
                if f == b"":
                    codes.append(-1); continue
                s = f.decode("utf-8", "replace")
                if s not in code_of:
                    code_of[s] = len(values); values.append(s)
                codes.append(code_of[s])
            types.append("code"); cols[name] = np.array(codes, np.int32); dicts[name] = values
    return names, types, cols, dicts
