"""ctypes binding of libb200flow.so (include/b200flow.h).

The CUDA library is the product path: there is no CPU fallback.  Importing this module
without a built libb200flow.so, or calling into it without a CUDA device, raises.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200flow.so")

F32, F64 = 0, 1
SRC_F32, SRC_F64, SRC_I32, SRC_INDEX, SRC_ONEHOT = 0, 1, 2, 3, 4

SLOT_DTYPE = np.dtype([("kind", "<i4"), ("src_off", "<i4"), ("lut_off", "<i4"), ("lut_len", "<i4"),
                       ("hot", "<i4"), ("reserved", "<i4"), ("mean", "<f8"), ("scale", "<f8")])
SPLIT_DTYPE = np.dtype([("feat", "<i4"), ("kind", "<i4"), ("bin_thr", "<i4"), ("flags", "<i4"),
                        ("gain", "<f8"), ("impurity", "<f8"), ("mask", "<u8", (4,))])
NODE_DTYPE = np.dtype([("feat", "<i4"), ("kind_bin", "<i4"), ("left", "<i4"), ("nid", "<u4")])
assert SLOT_DTYPE.itemsize == 40 and SPLIT_DTYPE.itemsize == 64 and NODE_DTYPE.itemsize == 16


class B200FlowError(RuntimeError):
    pass


class UnsupportedParamError(B200FlowError, ValueError):
    """a parameter value MLlib accepts but the B200 path does not implement (entropy impurity, maxBins > 256, ...): a
    ValueError, so the pyspark shim reports it as IllegalArgumentException; CUDA/runtime failures stay B200FlowError."""


_P, _I32, _I64, _U64, _F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double

# name -> argtypes, exactly the prototypes of include/b200flow.h
_SIGNATURES = {
    "b200flow_category_counts": [_P, _I64, _I32, _I32, _I32, _P, _P],
    "b200flow_category_counts_multi": [_P, _I64, _I32, _I32, _P, _P, _P, _P],
    "b200flow_encode": [_P, _I64, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _I32, _P, _I32, _P, _P, _P],
    "b200flow_sample_records": [_P, _I64, _I32, _P, _I32, _P, _I32, _U64, _U64, _I64, _P, _I64, _P, _P],
    "b200flow_encode_bins": [_P, _I64, _I32, _P, _I32, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _I32, _P, _I32, _P, _P, _P],
    "b200flow_column_moments": [_P, _I32, _I64, _I32, _I64, _P, _P, _P, _P],
    "b200flow_sample_rows": [_P, _I32, _I64, _I32, _I64, _U64, _U64, _I64, _P, _I64, _P, _P],
    "b200flow_find_splits": [_P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P],
    "b200flow_bin_rows": [_P, _I32, _I64, _I32, _I64, _P, _P, _P, _I32, _P, _P, _I32, _P, _P],
    "b200flow_dedup_rows": [_P, _I64, _I32, _I32, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P],
    "b200flow_bag_weights": [_U64, _I32, _I64, _I64, _P, _P, _P, _P, _I64, _P, _P],
    "b200flow_group_rows": [_P, _I64, _I64, _P, _P, _P, _P, _P, _P],
    "b200flow_bag_count": [_P, _I32, _I64, _P, _P],
    "b200flow_bag_fill": [_P, _I32, _I64, _P, _P, _P],
    "b200flow_exclusive_scan_i32_to_i64": [_P, _I64, _P, _P, _P],
    "b200flow_feature_subsets": [_U64, _I32, _P, _P, _I32, _I32, _P, _P],
    "b200flow_hist_level": [_P, _I32, _I32, _P, _I32, _P, _P, _P, _I64, _I32, _P, _I32, _I32, _I32, _P, _P],
    "b200flow_score_level": [_P, _I32, _P, _I32, _I32, _I32, _P, _P, _I32, _I32, _I32, _F64, _P, _P, _P, _P, _P],
    "b200flow_grow_level": [_I32, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P],
    "b200flow_route_hist_level": [_P, _I32, _I32, _P, _P, _I32, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _I32, _I32,
                                  _I32, _P, _I32, _P],
    "b200flow_partition_level": [_P, _I32, _P, _P, _I32, _P, _P, _P, _I64, _I32, _P, _P, _P],
    "b200flow_plan_route": [_I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P],
    "b200flow_next_segments": [_I32, _P, _P, _P, _P, _P, _P, _P, _P],
    "b200flow_finalize_forest": [_I64, _P, _I32, _P, _P],
    "b200flow_predict": [_P, _I32, _I64, _P, _P, _P, _P, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P],
    "b200flow_build_top_nodes": [_P, _P, _I64, _I32, _I32, _P, _P],
    "b200flow_gather_rows": [_P, _I32, _P, _I64, _P, _P],
    "b200flow_confusion": [_P, _P, _I64, _I32, _P, _P],
    "b200flow_random_split": [_U64, _I64, _I64, _P, _I32, _P, _P],
    "b200flow_compact_rows": [_P, _I64, _I32, _P, _I32, _P, _P, _P, _P],
    "b200flow_csv_count_lines": [_P, _I64, _P, _P, _P],
    "b200flow_csv_line_starts": [_P, _I64, _P, _P, _P],
    "b200flow_csv_infer": [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _P, _P],
    "b200flow_csv_dictionary": [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _P, _I32, _P, _P],
    "b200flow_csv_parse": [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _P, _P, _I32, _P, _I32, _P, _P],
}
EXPORTS = sorted(list(_SIGNATURES) + ["b200flow_last_error", "b200flow_version", "b200flow_route_hist_config"])

_lib = None
launches = 0   # kernels of OURS launched so far (counted per C-ABI call); bench.py reads the delta over the timed region
_KERNELS_PER_CALL = {"b200flow_grow_level": 3, "b200flow_compact_rows": 3, "b200flow_route_hist_level": 2, "b200flow_dedup_rows": 5, "b200flow_group_rows": 3}


def load():
    """dlopen libb200flow.so and declare every prototype; raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200FlowError("libb200flow.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; "
                                "g.build()'` or `make -C spark-network-traffic-classifier_b200/csrc`. There is no CPU "
                                "fallback for the product path." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        lib.b200flow_last_error.restype = C.c_char_p
        lib.b200flow_version.restype = C.c_int
        lib.b200flow_route_hist_config.argtypes = [_I32] * 4 + [C.POINTER(_I32), C.POINTER(_I32)]
        lib.b200flow_route_hist_config.restype = C.c_int
        _lib = lib
    return _lib


def route_hist_config(F, m, n_bins, n_classes):
    """launch shape of the fused route + histogram kernel: (chunk_rows, m_pass) or None when it cannot run (host-only call)."""
    ch, mp = _I32(0), _I32(0)
    ok = load().b200flow_route_hist_config(int(F), int(m), int(n_bins), int(n_classes), C.byref(ch), C.byref(mp))
    return (int(ch.value), int(mp.value)) if ok else None


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise B200FlowError("b200flow kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)
    if not t.is_contiguous():
        raise B200FlowError("b200flow kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """invoke one entry point on torch's current stream and raise on a non-zero return code."""
    global launches
    lib = load()
    rc = getattr(lib, name)(*args, stream())
    launches += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise B200FlowError("%s failed (%d): %s" % (name, rc, lib.b200flow_last_error().decode()))


def h2d(a, device):
    """small host array -> device tensor through a pinned staging block, asynchronously.  A pageable cudaMemcpy blocks the
    host until the stream reaches it, which stops the level loop from running ahead of the GPU; pinned + non_blocking does
    not (torch's caching host allocator keeps the block alive until the copy has run)."""
    t = torch.from_numpy(np.ascontiguousarray(a))
    if torch.device(device).type != "cuda":
        return t.clone()
    return t.pin_memory().to(device, non_blocking=True)


def require_cuda():
    if not torch.cuda.is_available():
        raise B200FlowError("no CUDA device: the b200flow product path has no CPU fallback")
    load()


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise B200FlowError("dense matrices must be float32 or float64, got %s" % t.dtype)
