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

    # fp32_tc (SPLIT) records: the halo mix kernel has a split form for C = 64 / 128 / 256, the halo dense kernel for the 16-channel
    # in-conv and for conv1 with the fused res_conv; every other split record runs the streamed kernel
    def sact(C, H, W):
        return E.Act(torch.zeros(1), C, H, W, 1, True, True)

    def mix_split(C):
        ol = E.OpList()
        kc, kb, nt, nsplit = E.tc_mix_tiling(C)
        E._tc_op(ol, split=1, src0=sact(C, 32, 32), w=1, tb=1, tg=1, gn=1, ncls=9, groups=8, kc=kc, kb=kb, nsplit=nsplit, nt=nt, mode=1, att=1,
                 attw=1, attw_stride=8, res=sact(C, 32, 32), dst=sact(C, 32, 32), ntot=8 * C, B=2, halo=1)
        return _lib.tc_schedule(ol.array()[0])

    def dense_split(C0, Cout, kc=64, fused=False, gn=1):
        ol = E.OpList()
        extra = dict(w2=1, tb2=1, dst_res=sact(Cout, 32, 32)) if fused else {}
        E._tc_op(ol, split=1, **extra, src0=sact(C0, 32, 32), w=1, tb=1, tg=1 if gn else 0, gn=gn, ncls=9 if gn else 1, act=1, dst=sact(Cout, 32, 32),
                 ntot=Cout, B=2, nt=E._tc_nt(Cout), halo=1, kc=kc)
        _lib.check_ops(ol.array(), 1)
        return _lib.tc_schedule(ol.array()[0])

    assert [mix_split(C) for C in (64, 128, 256, 512)] == [1, 1, 1, 0]
    assert dense_split(16, 64, kc=16, gn=0) == 2 and dense_split(128, 64, fused=True) == 2 and dense_split(64, 128, fused=True) == 2
    assert dense_split(128, 64) == 0 and dense_split(256, 256) == 0


def test_op_flags_are_part_of_the_record():
    """UCDIR_OP_FLAG_BRANCH / JOIN (ABI 14): carried in ucdir_op_t.flags, accepted by the argument checker, named in the header."""
    import torch
    from ucdir_b200 import engine as E
    assert _lib.C["UCDIR_OP_FLAG_BRANCH"] == 1 and _lib.C["UCDIR_OP_FLAG_JOIN"] == 2 and _lib.C["UCDIR_ABI_VERSION"] >= 14
    a = lambda C: E.Act(torch.zeros(1), C, 32, 32, 1, True)
    ol = E.OpList()
    E._tc_op(ol, src0=a(256), w=1, tb=1, nty=1, ntx=1, oy0=0, ox0=0, dst=a(256), ntot=256, B=2, nt=256, flags=_lib.C["UCDIR_OP_FLAG_BRANCH"])
    E._tc_op(ol, src0=a(256), w=1, tb=1, tg=1, gn=1, ncls=9, act=1, dst=a(256), ntot=256, B=2, nt=256)
    arr = ol.array()
    assert arr[0].flags == 1 and arr[1].flags == 0
    _lib.check_ops(arr, 2)


def test_c_abi_weight_packers_match_the_engine(tmp_path):
    """ucdir_pack_* (host-side C++, csrc/ucdir_pack.cu) against the torch packers the engine uses (engine.pack_tc_* / pack_conv_f32):
    bf16 operands bit for bit, fp32 tables up to summation order -- so a non-Python host gets the same kernel inputs."""
    import numpy as np
    import torch
    from ucdir_b200 import engine as E
    lib = _lib.load(require_device=False)
    F, U16, I = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int)
    fp = lambda a: a.ctypes.data_as(F) if a is not None else None
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    bits = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16)

    def np32(t):
        return None if t is None else np.ascontiguousarray(t.numpy(), dtype=np.float32)

    # dense 3x3 / 1x1, with and without folded GroupNorm, plain and split (concatenated input: c0 = 64 of 96 channels)
    for (co, ci, ks, nt, gn, split, c0) in [(64, 96, 3, 64, True, False, 0), (64, 96, 3, 64, True, True, 64), (48, 32, 1, 64, False, False, 0),
                                            (3, 64, 3, 16, False, True, 0), (128, 64, 1, 128, True, True, 0)]:
        w, b = r(co, ci, ks, ks) * 0.2, r(co)
        gam, bet = (r(ci) * 0.5 + 1, r(ci) * 0.1) if gn else (None, None)
        pw, ptb, ptg = E.pack_tc_dense(w, b, nt, gam, bet, split=split, c0=(c0 or None))
        we, ncls, ntot = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.ucdir_pack_tc_dense_sizes(co, ci, ks, nt, int(split), int(gn), ctypes.byref(we), ctypes.byref(ncls), ctypes.byref(ntot)) == 0
        assert we.value == pw.numel() and ntot.value == pw.shape[0] and ncls.value == ptb.shape[0]
        ow = np.zeros(we.value, np.uint16); otb = np.zeros((ncls.value, ntot.value), np.float32); otg = np.zeros_like(otb)
        wn, bn, gn_, ben = np32(w), np32(b), np32(gam), np32(bet)
        rc = lib.ucdir_pack_tc_dense(fp(wn), fp(bn), fp(gn_), fp(ben), co, ci, ks, nt, int(split), c0, ow.ctypes.data_as(U16), fp(otb), fp(otg) if gn else None)
        assert rc == ncls.value, _lib.last_error()
        assert np.array_equal(ow, bits(pw).reshape(-1)), (co, ci, ks, split)
        np.testing.assert_allclose(otb, ptb.numpy(), rtol=1e-5, atol=1e-5)
        if gn:
            np.testing.assert_allclose(otg, ptg.numpy(), rtol=1e-5, atol=1e-5)
    # grouped spdyconv for C = 64 (Cg = 8 < KC: zero-filled chunks) and C = 512, plain and split
    for (C, split) in [(64, False), (64, True), (512, False)]:
        cg, co = C // 8, 8 * C
        w, b, gam, bet = r(co, cg, 3, 3) * 0.2, r(co), r(C) * 0.5 + 1, r(C) * 0.1
        kc = E.tc_mix_tiling(C)[1]
        pw, ptb, ptg = E.pack_tc_grouped(w, b, 8, kc, gam, bet, split=split)
        we = ctypes.c_int()
        assert lib.ucdir_pack_tc_grouped_sizes(co, cg, 8, kc, int(split), ctypes.byref(we)) == 0 and we.value == pw.numel()
        ow = np.zeros(we.value, np.uint16); otb = np.zeros((9, co), np.float32); otg = np.zeros_like(otb)
        wn, bn, gn_, ben = np32(w), np32(b), np32(gam), np32(bet)
        assert lib.ucdir_pack_tc_grouped(fp(wn), fp(bn), fp(gn_), fp(ben), co, cg, 8, kc, int(split), ow.ctypes.data_as(U16), fp(otb), fp(otg)) == 9
        assert np.array_equal(ow, bits(pw).reshape(-1)), (C, split)
        np.testing.assert_allclose(otb, ptb.numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(otg, ptg.numpy(), rtol=1e-5, atol=1e-5)
    # upsample phases
    w, b = r(64, 64, 3, 3) * 0.2, r(64)
    for split in (False, True):
        for py in range(2):
            for px in range(2):
                pw, ptb = E.pack_tc_up_phase(w, b, py, px, 64, split=split)
                we, ntot = ctypes.c_int(), ctypes.c_int()
                assert lib.ucdir_pack_tc_up_phase_sizes(64, 64, 64, int(split), ctypes.byref(we), ctypes.byref(ntot)) == 0 and we.value == pw.numel()
                ow = np.zeros(we.value, np.uint16); otb = np.zeros(ntot.value, np.float32)
                wn, bn = np32(w), np32(b)
                assert lib.ucdir_pack_tc_up_phase(fp(wn), fp(bn), 64, 64, py, px, 64, int(split), ow.ctypes.data_as(U16), fp(otb)) == 0
                # the pre-summed weights are sums of up to four fp32 values: order-dependent in the last bit before the bf16 rounding
                diff = np.abs(ow.view(np.int16).astype(np.int32) - bits(pw).reshape(-1).view(np.int16).astype(np.int32))
                assert (diff <= (1 if not split else 2 ** 15)).all() and (diff != 0).mean() < (0.01 if not split else 0.2), (split, py, px)
                np.testing.assert_allclose(otb, ptb.numpy().reshape(-1), rtol=0, atol=0)
    # SIMT layout
    w = r(24, 6, 3, 3)
    want = E.pack_conv_f32(w, pad_cin_to=8)
    n = lib.ucdir_pack_conv_f32_size(24, 6, 3, 1, 8)
    assert n == want.numel()
    out = np.zeros(n, np.float32)
    wn = np32(w)
    assert lib.ucdir_pack_conv_f32(fp(wn), 24, 6, 3, 1, 8, fp(out)) == 0
    assert np.array_equal(out, want.numpy().reshape(-1))
