"""CPU-side checks of the C ABI: the in-tree library loads without a GPU, exports every function include/ucdir_b200.h
declares, agrees with the header on the ABI version and record size, and validates op records without a device."""
import ctypes
import os
import re

import pytest

from ucdir_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "ucdir_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ucdir_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load(require_device=False)
    names = declared_functions()
    assert len(names) >= 10 and "ucdir_run_ops" in names and "ucdir_graph_capture" in names
    for n in names:
        assert hasattr(lib, n), "header declares %s but libucdir_b200.so does not export it" % n
    assert set(_lib.EXPORTS) == set(names), "ucdir_b200/_lib.py EXPORTS and the header disagree"


def test_abi_version_and_record_size():
    lib = _lib.load(require_device=False)
    assert lib.ucdir_abi_version() == _lib.C["UCDIR_ABI_VERSION"]
    assert lib.ucdir_op_sizeof() == ctypes.sizeof(_lib.Op)


def test_check_ops_rejects_bad_records_without_a_device():
    bad = (_lib.Op * 1)(_lib.make_op("UCDIR_OP_CONV_F32"))          # all pointers null
    with pytest.raises(_lib.UcdirLibraryError):
        _lib.check_ops(bad, 1)
    unknown = (_lib.Op * 1)(_lib.make_op(9999))
    with pytest.raises(_lib.UcdirLibraryError):
        _lib.check_ops(unknown, 1)
