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


def test_tc_schedule_routing():
    """Which kernel a TC_CONV record is routed to is decided on the host (ucdir_tc_schedule): the halo kernels take the
    integration-module convs with C <= 256 and the gn-folded 3x3 convs with 64 / 128 output channels, only when asked to."""
    import torch
    from ucdir_b200 import engine as E

    def act(C, H, W):
        return E.Act(torch.zeros(1), C, H, W, 1, True)

    def mix(C, halo):
        ol = E.OpList()
        kc, kb, nt, nsplit = E.tc_mix_tiling(C)
        E._tc_op(ol, src0=act(C, 32, 32), w=1, tb=1, tg=1, gn=1, ncls=9, groups=8, kc=kc, kb=kb, nsplit=nsplit, nt=nt, mode=1, att=1,
                 attw=1, attw_stride=8, res=act(C, 32, 32), dst=act(C, 32, 32), ntot=8 * C, B=2, halo=halo)
        return _lib.tc_schedule(ol.array()[0])

    def dense(C0, Cout, halo, ks=3, gn=1, stride=1):
        ol = E.OpList()
        E._tc_op(ol, src0=act(C0, 32 * stride, 32 * stride), w=1, tb=1, tg=1 if gn else 0, gn=gn, ncls=9 if (gn and ks == 3) else 1, nty=ks,
                 ntx=ks, oy0=-(ks // 2), ox0=-(ks // 2), stride=stride, act=1, dst=act(Cout, 32, 32), ntot=Cout, B=2, nt=E._tc_nt(Cout), halo=halo)
        return _lib.tc_schedule(ol.array()[0])

    assert [mix(C, 1) for C in (64, 128, 256, 512)] == [1, 1, 1, 0]
    assert [mix(C, 0) for C in (64, 128, 256, 512)] == [0, 0, 0, 0]
    assert dense(64, 64, 1) == 2 and dense(192, 64, 1) == 2 and dense(384, 128, 1) == 2
    assert dense(64, 64, 0) == 0 and dense(256, 256, 1) == 0 and dense(64, 64, 1, ks=1) == 0
    assert dense(64, 64, 1, gn=0) == 2 and dense(64, 64, 1, stride=2) == 0
