"""ctypes binding of libucdir_b200.so (C ABI in include/ucdir_b200.h).

The library is the product: if it is missing or the device is not sm_100 this module raises --
there is no PyTorch / CPU fallback path anywhere in ucdir_b200.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "ucdir_b200.h")
LIB_PATH = os.path.join(HERE, "libucdir_b200.so")


def _parse_header(path):
    """Pull every `UCDIR_X = n` enumerator and `#define UCDIR_X n` out of the header so the Python side
    can never drift from the C side."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    consts = {}
    for m in re.finditer(r"#define\s+(UCDIR_\w+)\s+(\d+)", txt):
        consts[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"\b(UCDIR_\w+)\s*=\s*(\d+)", txt):
        consts[m.group(1)] = int(m.group(2))
    return consts


C = _parse_header(HEADER)
NPTR, NINT, NFLT = C["UCDIR_OP_NPTR"], C["UCDIR_OP_NINT"], C["UCDIR_OP_NFLT"]


class Op(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("flags", ctypes.c_int32), ("p", ctypes.c_void_p * NPTR),
                ("i", ctypes.c_int32 * NINT), ("f", ctypes.c_float * NFLT)]


EXPORTS = ["ucdir_run_ops", "ucdir_check_ops", "ucdir_abi_version", "ucdir_op_sizeof", "ucdir_last_error",
           "ucdir_launch_count", "ucdir_device_ok", "ucdir_profile_begin", "ucdir_profile_end", "ucdir_graph_capture",
           "ucdir_graph_launch", "ucdir_graph_destroy", "ucdir_tc_schedule",
           "ucdir_pack_tc_dense_sizes", "ucdir_pack_tc_dense", "ucdir_pack_tc_grouped_sizes", "ucdir_pack_tc_grouped",
           "ucdir_pack_tc_up_phase_sizes", "ucdir_pack_tc_up_phase", "ucdir_pack_conv_f32_size", "ucdir_pack_conv_f32"]

_lib = None
_device_ok = False


class UcdirLibraryError(RuntimeError):
    pass


def load(require_device=True):
    """Load the shared library (building is __graft_entry__.build()'s / ucdir_b200.build's job)."""
    global _lib, _device_ok
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UcdirLibraryError(
                "libucdir_b200.so is missing (%s). Build it with `python -m ucdir_b200.build`; "
                "ucdir_b200 has no CPU / PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.ucdir_run_ops.argtypes = [ctypes.POINTER(Op), ctypes.c_int, ctypes.c_void_p]
        lib.ucdir_run_ops.restype = ctypes.c_int
        lib.ucdir_check_ops.argtypes = [ctypes.POINTER(Op), ctypes.c_int]
        lib.ucdir_check_ops.restype = ctypes.c_int
        lib.ucdir_abi_version.restype = ctypes.c_int
        lib.ucdir_op_sizeof.restype = ctypes.c_int
        lib.ucdir_last_error.restype = ctypes.c_char_p
        lib.ucdir_launch_count.restype = ctypes.c_longlong
        lib.ucdir_device_ok.restype = ctypes.c_int
        lib.ucdir_tc_schedule.argtypes = [ctypes.POINTER(Op)]
        lib.ucdir_tc_schedule.restype = ctypes.c_int
        lib.ucdir_graph_capture.argtypes = [ctypes.POINTER(Op), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        lib.ucdir_graph_capture.restype = ctypes.c_int
        lib.ucdir_graph_launch.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.ucdir_graph_launch.restype = ctypes.c_int
        lib.ucdir_graph_destroy.argtypes = [ctypes.c_void_p]
        lib.ucdir_graph_destroy.restype = ctypes.c_int
        lib.ucdir_profile_begin.restype = ctypes.c_int
        lib.ucdir_profile_end.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        lib.ucdir_profile_end.restype = ctypes.c_int
        if lib.ucdir_abi_version() != C["UCDIR_ABI_VERSION"]:
            raise UcdirLibraryError("libucdir_b200.so ABI %d != header ABI %d: rebuild" % (
                lib.ucdir_abi_version(), C["UCDIR_ABI_VERSION"]))
        if lib.ucdir_op_sizeof() != ctypes.sizeof(Op):
            raise UcdirLibraryError("ucdir_op_t size mismatch: C %d vs ctypes %d" % (lib.ucdir_op_sizeof(), ctypes.sizeof(Op)))
        _lib = lib
    if require_device and not _device_ok:
        # once per process: ucdir_device_ok() is a cudaGetDeviceProperties call (milliseconds), and this function sits on the
        # path of every stateless p_sample() / forward() call
        rc = _lib.ucdir_device_ok()
        if rc != 0:
            raise UcdirLibraryError("ucdir_b200 needs a B200 (sm_100) CUDA device: %s" % last_error())
        _device_ok = True
    return _lib


def last_error():
    return (_lib.ucdir_last_error() or b"").decode() if _lib is not None else ""


def make_op(kind, p=None, i=None, f=None):
    op = Op()
    op.kind = C[kind] if isinstance(kind, str) else kind
    for k, v in (p or {}).items():
        op.p[C[k] if isinstance(k, str) else k] = v
    for k, v in (i or {}).items():
        op.i[C[k] if isinstance(k, str) else k] = int(v)
    for k, v in (f or {}).items():
        op.f[C[k] if isinstance(k, str) else k] = float(v)
    return op


def run_ops(ops, n, stream):
    """ops: ctypes array of Op.  stream: integer cudaStream_t handle."""
    rc = load().ucdir_run_ops(ops, n, ctypes.c_void_p(stream))
    if rc != 0:
        raise UcdirLibraryError("ucdir_run_ops failed (%d): %s" % (rc, last_error()))


def check_ops(ops, n):
    rc = load(require_device=False).ucdir_check_ops(ops, n)
    if rc != 0:
        raise UcdirLibraryError("ucdir_check_ops failed (%d): %s" % (rc, last_error()))


def tc_schedule(op) -> int:
    """0 = streamed tc_conv_kernel, 1 = halo mix kernel, 2 = halo dense kernel (see ucdir_tc_schedule in the header)."""
    return int(load(require_device=False).ucdir_tc_schedule(ctypes.byref(op)))


def launch_count():
    return int(load(require_device=False).ucdir_launch_count())


def profile_begin():
    load().ucdir_profile_begin()


def profile_end(cap=1 << 20):
    """-> list of (ms, op_index, kind) for every op run since profile_begin()."""
    ms = (ctypes.c_float * cap)()
    idx = (ctypes.c_int * cap)()
    kind = (ctypes.c_int * cap)()
    n = load().ucdir_profile_end(ms, idx, kind, cap)
    if n < 0:
        raise UcdirLibraryError("ucdir_profile_end failed: %s" % last_error())
    return [(ms[k], idx[k], kind[k]) for k in range(n)]


class Graph:
    """A captured op list (CUDA graph).  Replay = one launch; see ucdir_graph_capture in the header."""

    def __init__(self, ops, n):
        h = ctypes.c_void_p()
        rc = load().ucdir_graph_capture(ops, n, ctypes.byref(h))
        if rc != 0:
            raise UcdirLibraryError("ucdir_graph_capture failed (%d): %s" % (rc, last_error()))
        self.h = h

    def launch(self, stream):
        rc = _lib.ucdir_graph_launch(self.h, ctypes.c_void_p(stream))
        if rc != 0:
            raise UcdirLibraryError("ucdir_graph_launch failed (%d): %s" % (rc, last_error()))

    def __del__(self):
        try:
            if _lib is not None and self.h:
                _lib.ucdir_graph_destroy(self.h)
        except Exception:
            pass
