"""the device CSV reader (csrc/csv.cu via b200flow.csvio / the shim's spark.read.csv) against the pure-Python
restatement of Spark's reader in oracle/csv_ref.py — records byte for byte, dictionaries string for string."""
import os

import numpy as np
import pytest
import torch

from oracle import csv_ref

pytestmark = pytest.mark.gpu


def _check(paths, **kw):
    from b200flow import csvio
    rec, schema, dicts = csvio.read_csv(paths, kw.get("header", False), kw.get("infer_schema", False), kw.get("strip_lead", False),
                                        kw.get("strip_trail", False))
    names, types, cols, want_dicts = csv_ref.read_csv(paths, **kw)
    assert schema.names == names and schema.types == types, (schema.names, schema.types, names, types)
    host = rec.cpu().numpy().view(schema.numpy_dtype()).reshape(-1) if rec.shape[0] else None
    n = len(next(iter(cols.values()))) if cols else 0
    assert rec.shape[0] == n
    for name, typ in zip(names, types):
        got = host[name] if host is not None else np.zeros(0)
        if typ == "f64":
            assert np.array_equal(np.asarray(got, np.float64).view(np.uint64), cols[name].view(np.uint64)), name     # bit for bit (NaN too)
        else:
            assert np.array_equal(got, cols[name]), name
    assert dicts == want_dicts
    return rec, schema, dicts


def _kdd_lines(n, seed):
    from b200flow import synth
    rec, dicts = synth.make_kdd(n, 23, seed=seed, device="cuda")
    a = rec.cpu().numpy().view(synth.kdd_schema().numpy_dtype()).reshape(-1)
    out = []
    for r in a:
        cells = []
        for name in synth.KDD_COLUMNS:
            v = r[name]
            if name in dicts:
                cells.append(dicts[name][int(v)] + ("." if name == "label" else ""))
            elif name in synth.KDD_RATE:
                cells.append("%.2f" % v)
            else:
                cells.append("%d" % int(v))
        out.append(",".join(cells))
    return out


def test_kdd_shaped_file(tmp_path):
    p = str(tmp_path / "kddcup.data.corrected")
    lines = _kdd_lines(20000, 3)
    open(p, "w").write("\n".join(lines) + "\n")
    rec, schema, dicts = _check([p], infer_schema=True)
    assert schema.types.count("code") == 4 and schema.types.count("f64") == 15 and len(schema.names) == 42
    # no trailing newline, CRLF line ends, blank lines in between: same records
    q = str(tmp_path / "crlf.csv")
    open(q, "wb").write(("\r\n".join(lines[:5000]) + "\r\n\r\n\r\n" + "\r\n".join(lines[5000:])).encode())
    rec2, schema2, dicts2 = _check([q], infer_schema=True)
    assert torch.equal(rec, rec2) and dicts == dicts2
    # inferSchema=False: every column is a string column (numeric ones have many distinct values: the hash tables grow)
    _check([p], infer_schema=False)


def test_cicids_shaped_files_glob_header_whitespace_specials(tmp_path):
    rng = np.random.default_rng(5)
    header = " Destination Port, Flow Duration,Total Fwd Packets, Fwd Header Length, Flow Bytes/s, Flow Packets/s, Fwd Header Length, Idle Min, Label"
    labels = ["BENIGN", "DoS Hulk", "PortScan", "Web Attack � Brute Force", "Bot"]
    paths = []
    for k, n in enumerate([7000, 3000, 1]):
        rows = []
        for i in range(n):
            dur = int(rng.integers(1, 120000000))
            fb = rng.random()
            flow_bytes = "NaN" if fb < 0.01 else ("Infinity" if fb < 0.02 else repr(float(rng.standard_normal() * 10 ** rng.uniform(-3, 9))))
            pk = "%.9f" % (rng.random() * 1e6) if rng.random() < 0.5 else "%.17g" % (rng.random() * 1e6)
            rows.append("%d, %d,%d,%d,%s, %s,%d,%s,%s" % (rng.integers(0, 65536), dur, rng.integers(1, 200000), rng.integers(-5, 5000) * 1000000,
                                                         flow_bytes, pk, rng.integers(0, 4000), "" if rng.random() < 0.05 else str(rng.integers(0, 10 ** 12)),
                                                         labels[int(rng.integers(0, 5))]))
        p = str(tmp_path / ("day%d.pcap_ISCX.csv" % k))
        open(p, "w", encoding="utf-8").write(header + "\n" + "\n".join(rows) + ("\n" if k != 1 else ""))
        paths.append(p)
    rec, schema, dicts = _check(paths, header=True, infer_schema=True, strip_lead=True, strip_trail=True)
    assert schema.names[3] == "Fwd Header Length3" and schema.names[6] == "Fwd Header Length6" and schema.names[0] == "Destination Port"
    assert dict(zip(schema.names, schema.types)) == {"Destination Port": "i32", "Flow Duration": "i32", "Total Fwd Packets": "i32",
                                                      "Fwd Header Length3": "f64", "Flow Bytes/s": "f64", "Flow Packets/s": "f64",
                                                      "Fwd Header Length6": "i32", "Idle Min": "f64", "Label": "code"}
    assert rec.shape[0] == 10001 and sorted(dicts["Label"]) == sorted(labels)
    # without the whitespace options " 123" is not an integer for Java's parseInt but parses as a double
    _, schema_raw, _ = _check(paths, header=True, infer_schema=True)
    assert schema_raw.type_of[" Flow Duration"] == "f64" and schema_raw.type_of["Total Fwd Packets"] == "i32"


def test_type_lattice_nulls_and_edge_literals(tmp_path):
    p = str(tmp_path / "edge.csv")
    rows = ["1,2147483647,1,,x,,1e3,-0", "-2,2147483648,2.5,,,,.5,+7", "+3,-9223372036854775808,-7,,y z,,5.,0007",
            "4,12,9007199254740993,,x,,0.30000000000000004,12", "5,13,1e-5,,tcp,,123456789012345678,-2147483648"]
    open(p, "w").write("\n".join(rows) + "\n")
    rec, schema, dicts = _check([p], infer_schema=True)
    assert schema.types == ["i32", "f64", "f64", "code", "code", "code", "f64", "i32"]      # all-null columns are (empty) string columns
    assert dicts["_c4"] == ["x", "y z", "tcp"] and dicts["_c3"] == [] and dicts["_c5"] == []


def test_inputs_the_reader_refuses(tmp_path):
    from b200flow import csvio

    def write(name, text):
        p = str(tmp_path / name); open(p, "w").write(text); return p
    with pytest.raises(csvio.CsvFormatError, match="do not have 3 fields"):
        csvio.read_csv([write("ragged.csv", "1,2,3\n4,5\n6,7,8\n")], infer_schema=True)
    with pytest.raises(csvio.CsvFormatError, match="quoted"):
        csvio.read_csv([write("quoted.csv", '1,"a,b",3\n')], infer_schema=True)
    with pytest.raises(csvio.CsvFormatError, match="longer than 4096"):
        csvio.read_csv([write("long.csv", "1,2\n" + "9" * 5000 + ",3\n")], infer_schema=True)
    with pytest.raises(csvio.CsvFormatError, match="exact"):
        csvio.read_csv([write("tiny.csv", "1.5,2\n1e-400,3\n")], infer_schema=True)
    rec, schema, dicts = csvio.read_csv([write("empty.csv", "")], infer_schema=True)
    assert rec.shape[0] == 0 and schema.names == []
    rec, schema, dicts = csvio.read_csv([write("header_only.csv", "a,b\n")], header=True, infer_schema=True)
    assert rec.shape[0] == 0 and schema.names == ["a", "b"] and schema.types == ["code", "code"]


def test_shim_reader_device_engine_equals_host_engine(tmp_path):
    from pyspark.sql import SparkSession
    p = str(tmp_path / "kdd.csv")
    open(p, "w").write("\n".join(_kdd_lines(8000, 11)) + "\n")
    spark = SparkSession.builder.appName("t").master("local").getOrCreate()
    dev = spark.read.csv(p, inferSchema=True, header=False)
    host = spark.read.option("b200flow.csvEngine", "host").csv(p, inferSchema=True, header=False)
    assert dev.columns == host.columns and dev.count() == host.count() == 8000
    assert dev._schema.types == host._schema.types and torch.equal(dev._rec, host._rec) and dev._dicts == host._dicts
