"""CPU-side checks of the drop-in boundary: libb200flow.so loads and exports every symbol that
include/b200flow.h declares; the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from b200flow import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200flow.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200flow_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libb200flow.so does not export %s" % n
    assert sorted(names) == _lib.EXPORTS, "ctypes binding and header disagree"
    assert _lib.load().b200flow_version() >= 100


def test_struct_layouts_match_header():
    assert _lib.SLOT_DTYPE.itemsize == 40 and _lib.SLOT_DTYPE.fields["mean"][1] == 24
    assert _lib.SPLIT_DTYPE.itemsize == 64 and _lib.SPLIT_DTYPE.fields["mask"][1] == 32
    assert _lib.NODE_DTYPE.itemsize == 16


def test_no_cpu_fallback():
    with pytest.raises(_lib.B200FlowError):
        _lib.ptr(torch.zeros(4))                     # CPU tensor is refused
    if not torch.cuda.is_available():
        with pytest.raises(_lib.B200FlowError):
            _lib.require_cuda()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spark-network-traffic-classifier_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, os.path.join(d, f)
