"""Host side of the fused encode path: record schema, encode plans, fit statistics.

Mirrors, for the raw-record representation, what the reference does with
StringIndexer / OneHotEncoder / StandardScaler / VectorAssembler
(kdd99.py:34-37,45-46; cicids17.py:41-46).  All arithmetic runs in libb200flow.so.
"""
import numpy as np
import torch

from . import _lib
from ._lib import SLOT_DTYPE, SRC_F32, SRC_F64, SRC_I32, SRC_INDEX, SRC_ONEHOT, call, ptr

_KIND_OF = {"f32": SRC_F32, "f64": SRC_F64, "i32": SRC_I32}
_SIZE_OF = {"f32": 4, "f64": 8, "i32": 4, "code": 4}


class RecordSchema:
    """Layout of one raw flow record (AoS): fields are f32 / f64 / i32 numbers or int32 dictionary
    codes ('code') with a host-side string dictionary.  Fields are packed in order, 4-byte aligned."""

    def __init__(self, fields):
        self.names, self.types, self.offsets = [], [], {}
        off = 0
        for name, typ in fields:
            if typ not in _SIZE_OF:
                raise ValueError("unknown field type %r" % typ)
            self.names.append(name); self.types.append(typ); self.offsets[name] = off
            off += _SIZE_OF[typ]
        self.row_bytes = off
        self.type_of = dict(zip(self.names, self.types))

    def numpy_dtype(self):
        np_of = {"f32": "<f4", "f64": "<f8", "i32": "<i4", "code": "<i4"}
        return np.dtype({"names": self.names, "formats": [np_of[t] for t in self.types],
                         "offsets": [self.offsets[n] for n in self.names], "itemsize": self.row_bytes})


def category_counts(records, schema, field, K):
    """StringIndexer.fit, counting half (R1): occurrences of each dictionary code -> int64[K] (device)."""
    counts = torch.zeros(K, dtype=torch.int64, device=records.device)
    n = records.shape[0]
    call("b200flow_category_counts", ptr(records), n, schema.row_bytes, schema.offsets[field], K, ptr(counts))
    return counts


def category_counts_multi(records, schema, fields, Ks):
    """StringIndexer.fit counting for several code fields in one pass over the records -> list of int64[K] device tensors
    (views of one concatenated buffer: read them all with a single .cpu() on `result[0]._base_all`)."""
    Ks = [max(int(k), 1) for k in Ks]
    if len(fields) > 8 or sum(Ks) > 8192:
        return [category_counts(records, schema, f, k) for f, k in zip(fields, Ks)]
    allc = torch.zeros(sum(Ks), dtype=torch.int64, device=records.device)
    offs = np.asarray([schema.offsets[f] for f in fields], np.int32)
    ks = np.asarray(Ks, np.int32)
    call("b200flow_category_counts_multi", ptr(records), records.shape[0], schema.row_bytes, len(fields), offs.ctypes.data, ks.ctypes.data,
         ptr(allc))
    out, o = [], 0
    for k in Ks:
        out.append(allc[o:o + k]); o += k
    return out


def string_index_order(counts, labels):
    """StringIndexer.fit, ordering half (A.7): frequencyDesc, ties alphabetical; never-seen codes get rank -1.
    counts: sequence of ints (host); returns (ordered labels, int32 lut code->rank)."""
    idx = [i for i in range(len(labels)) if counts[i] > 0]
    idx.sort(key=lambda i: (-int(counts[i]), labels[i]))
    lut = np.full(len(labels), -1, np.int32)
    for rank, i in enumerate(idx):
        lut[i] = rank
    return [labels[i] for i in idx], lut


class EncodePlan:
    """One descriptor per output slot of the assembled feature vector + the LUT pool."""

    def __init__(self, schema):
        self.schema = schema
        self.slots = []          # tuples (kind, src_off, lut_off, lut_len, hot, mean, scale)
        self.luts = []           # list of int32 arrays, concatenated into the pool
        self._lut_total = 0
        self.label = None        # (field offset, lut_off, lut_len)
        self.check_nan = 0
        self._dev = None

    # ---- building -------------------------------------------------------------------
    def _add_lut(self, lut):
        lut = np.ascontiguousarray(lut, np.int32)
        off = self._lut_total
        self.luts.append(lut); self._lut_total += len(lut)
        return off, len(lut)

    def add_numeric(self, field, mean=0.0, scale=1.0):
        typ = self.schema.type_of[field]
        if typ == "code":
            raise ValueError("field %s is a dictionary code; use add_index/add_onehot" % field)
        self.slots.append((_KIND_OF[typ], self.schema.offsets[field], 0, 0, 0, mean, scale)); self._dev = None
        return self

    def add_index(self, field, lut, mean=0.0, scale=1.0):
        off, ln = self._add_lut(lut)
        self.slots.append((SRC_INDEX, self.schema.offsets[field], off, ln, 0, mean, scale)); self._dev = None
        return self

    def add_onehot(self, field, lut, n_categories, drop_last=True, means=None, scales=None):
        off, ln = self._add_lut(lut)
        width = n_categories - 1 if drop_last else n_categories
        for k in range(width):
            self.slots.append((SRC_ONEHOT, self.schema.offsets[field], off, ln, k,
                               0.0 if means is None else float(means[k]), 1.0 if scales is None else float(scales[k])))
        self._dev = None
        return self

    def set_label(self, field, lut):
        off, ln = self._add_lut(lut)
        self.label = (self.schema.offsets[field], off, ln); self._dev = None
        return self

    def set_scaling(self, mean, scale):
        """StandardScaler: per-slot (mean, scale) applied on top of the assembled vector."""
        assert len(mean) == len(self.slots) and len(scale) == len(self.slots)
        self.slots = [s[:5] + (float(mu), float(sc)) for s, mu, sc in zip(self.slots, mean, scale)]
        self._dev = None
        return self

    @property
    def n_out(self):
        return len(self.slots)

    def slot_array(self):
        a = np.zeros(len(self.slots), SLOT_DTYPE)
        for i, (kind, so, lo, ll, hot, mean, scale) in enumerate(self.slots):
            a[i] = (kind, so, lo, ll, hot, 0, mean, scale)
        return a

    def lut_array(self):
        return np.concatenate(self.luts).astype(np.int32) if self.luts else np.zeros(0, np.int32)

    def algorithmic_bytes_per_row(self, out_dtype=torch.float32):
        """record in + dense vector out + 4-byte label."""
        return self.schema.row_bytes + self.n_out * (4 if out_dtype == torch.float32 else 8) + (4 if self.label else 0)

    # ---- running --------------------------------------------------------------------
    def _device_tables(self, device):
        if self._dev is None or self._dev[0] != device:
            slots = _lib.h2d(self.slot_array().view(np.uint8), device)
            lut = self.lut_array()
            lut_t = _lib.h2d(lut, device) if len(lut) else None
            self._dev = (device, slots, lut_t, len(lut))
        return self._dev

    def run(self, records, out_dtype=torch.float32, out=None, want_valid=True, want_label=True):
        """records: uint8 [n, row_bytes] CUDA tensor -> (features [n, n_out], label int32 [n] | None, valid uint8 [n] | None)."""
        if records.dtype != torch.uint8 or records.dim() != 2 or records.shape[1] != self.schema.row_bytes:
            raise ValueError("records must be uint8 [n, %d]" % self.schema.row_bytes)
        n = records.shape[0]
        _, slots, lut_t, lut_total = self._device_tables(records.device)
        if out is None:
            out = torch.empty((n, self.n_out), dtype=out_dtype, device=records.device)
        label_out = torch.empty(n, dtype=torch.int32, device=records.device) if (self.label and want_label) else None
        valid = torch.empty(n, dtype=torch.uint8, device=records.device) if want_valid else None
        loff, llo, lln = self.label if self.label else (-1, 0, 0)
        from .forest import _timed
        _timed("encode", "b200flow_encode", ptr(records), n, self.schema.row_bytes, ptr(slots), self.n_out, ptr(lut_t), lut_total,
             loff, llo, lln, int(self.check_nan), ptr(out), _lib.dtype_code(out), ptr(label_out), ptr(valid))
        return out, label_out, valid


def column_moments(x, group=None):
    """StandardScaler.fit (R3c): per-column mean and unbiased std of a dense [n, D] CUDA matrix, corrected
    two-pass in fp64; with a process group the partial sums are all-reduced between the passes."""
    import torch.distributed as dist
    n, D = x.shape
    dev = x.device
    buf = torch.zeros(2 * D + 1, dtype=torch.float64, device=dev)
    call("b200flow_column_moments", ptr(x), _lib.dtype_code(x), n, D, x.stride(0), None, ptr(buf[:D]), ptr(buf[D:2 * D]))
    buf[2 * D] = float(n)
    if group is not None:
        from .dist import all_reduce_
        all_reduce_(buf, group)
    n_tot = buf[2 * D]
    mean = (buf[:D] / torch.clamp(n_tot, min=1.0)).contiguous()
    buf2 = torch.zeros(2 * D, dtype=torch.float64, device=dev)
    call("b200flow_column_moments", ptr(x), _lib.dtype_code(x), n, D, x.stride(0), ptr(mean), ptr(buf2[:D]), ptr(buf2[D:]))
    if group is not None:
        all_reduce_(buf2, group)
    m2 = buf2[D:] - buf2[:D] * buf2[:D] / torch.clamp(n_tot, min=1.0)
    var = torch.where(n_tot > 1, m2 / torch.clamp(n_tot - 1.0, min=1.0), torch.zeros_like(m2))
    return mean, torch.sqrt(torch.clamp(var, min=0.0))
