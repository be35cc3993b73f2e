"""The C-ABI library loads without a GPU and exports every symbol include/yaha_b200.h declares."""
import ctypes
import os
import re

import pytest

import yaha_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "yaha_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ya_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(yaha_b200.EXPORTS)


def test_library_exports_all_symbols():
    if not os.path.exists(yaha_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(yaha_b200.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    # Fragment_t compatibility (Math.h:448-455) and the POD layouts the ABI promises
    assert yaha_b200.FRAG_DT.itemsize == 12
    assert yaha_b200.JOB_DT.itemsize == 16
    assert yaha_b200.RES_DT.itemsize == 16
    assert yaha_b200.OP_DT.itemsize == 4
    assert yaha_b200.STRAND_DT.itemsize == 16
    assert ctypes.sizeof(yaha_b200.Params) == 48


def test_no_cpu_fallback(small):
    """Without a GPU the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(yaha_b200.YahaError):
        yaha_b200.Aligner(small.nib, small.idx)
