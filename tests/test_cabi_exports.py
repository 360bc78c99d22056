"""The C-ABI library loads without a GPU and exports every symbol include/fmgpu.h declares; without a
device the product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "fmgpu.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from index4j_b200.fm_index import native
    lib = native()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.fmgpu_version().startswith(b"fmgpu")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from index4j_b200 import FmIndex, build_index
    from index4j_b200.fm_index import FmIndexError
    with pytest.raises(FmIndexError, match="no CUDA device"):
        FmIndex.read(build_index("some text\nto index\n"))


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "index4j_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc")):
                with open(os.path.join(base, f), errors="ignore") as fh:
                    src = fh.read()
                assert "pyoracle" not in src and "liboracle" not in src.replace("oracle/liboracle.so", "") or f == "_build.py", f


def test_null_arguments_are_rejected():
    from index4j_b200.fm_index import native
    lib = native()
    assert lib.fmgpu_index_load_serialized(None, 0, None, None) == -1
    assert b"null" in lib.fmgpu_last_error()
    assert lib.fmgpu_count_batch(None, None, None, 0, None, None) == -1
    assert lib.fmgpu_input_length(None) == -1
