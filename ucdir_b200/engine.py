"""Host-side engine: turns the reference's module graph into arrays of `ucdir_op_t` records and hands
them to libucdir_b200.so (C ABI, include/ucdir_b200.h).  One `ucdir_run_ops` call per denoising step.

What is here is plumbing only: weight re-packing (OIHW fp32 -> kernel layouts, once per load), tile
geometry (utils/util.py:108-146 and model/ucdir.py:295-307 of the reference restated as index tables),
buffer planning, and op-list construction.  Every arithmetic step of the hot path is a CUDA kernel in
the library; torch is used for device memory and streams.  There is no eager / CPU compute path.

Data layout in HBM: activations are NHWC ("pixel-major, channel-innermost") over a *tile batch*
[BT, TH, TW, C]; images at the boundary stay in the reference's NCHW fp32.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import C as K_

F32 = torch.float32

# Test seam: tests/ may swap the op runner for the CPU interpreter in tests/op_emulator.py to check graph
# construction without a GPU.  The product never sets these; on a non-CUDA device every entry point raises.
_RUNNER = _lib.run_ops
_TEST_CPU_PLAN = False


def _run_ops(arr, n, stream):
    return _RUNNER(arr, n, stream)


def _require_cuda(dev):
    if _TEST_CPU_PLAN:
        return
    if dev.type != "cuda":
        raise _lib.UcdirLibraryError("ucdir_b200 runs on a CUDA device only (module is on %s); there is no CPU "
                                     "fallback" % dev)
    _lib.load()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0


class _NullGuard:
    def __enter__(self): return self
    def __exit__(self, *a): return False


_NULL_GUARD = _NullGuard()


_NVTX = os.environ.get("UCDIR_NVTX", "0") == "1"
"""UCDIR_NVTX=1: NVTX ranges around the host-visible phases (bind / guidance maps, a denoising step, the tile all-gather) here and
around every op inside libucdir_b200.so (c_abi.cu) -- for nsys / `ncu --nvtx` captures.  Off by default (push/pop cost)."""


class _Range:
    def __init__(self, name): self.name = name
    def __enter__(self):
        if _NVTX: torch.cuda.nvtx.range_push(self.name)
    def __exit__(self, *a):
        if _NVTX: torch.cuda.nvtx.range_pop()
        return False


def _on_device(fn):
    """Method decorator: run with self.dev as the current CUDA device (see _device_guard)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        g = _device_guard(self.dev)
        if g is _NULL_GUARD:
            return fn(self, *a, **k)
        with g:
            return fn(self, *a, **k)
    return wrapper


def _device_guard(dev):
    """Make `dev` the current CUDA device around library calls: ucdir_run_ops / graph capture launch on the CURRENT device
    with a stream handle of `dev`, so a process that touches a second GPU must switch first.  A no-op when it already is
    current (the one-process-per-GPU launch) and on the CPU test seam."""
    if dev.type != "cuda" or _TEST_CPU_PLAN or torch.cuda.current_device() == (dev.index or 0):
        return _NULL_GUARD
    return torch.cuda.device(dev)


# ======================================================================================
# geometry: which windows of which (reflect padded) image form the tile batch
# ======================================================================================
def tile_windows(length: int, skip: int, padding: int) -> List[int]:
    """Window origins along one padded axis, reference order (utils/util.py:122-134): i in
    arange(0, L, skip - 2*padding), a window that would overrun is shifted back to end at L."""
    shift = skip - 2 * padding
    if shift <= 0:
        raise ValueError("tiler stride skip-2*padding must be positive (skip=%d padding=%d)" % (skip, padding))
    return [min(i, length - skip) for i in range(0, length, shift)]


@dataclass
class Geometry:
    """Tile batch description.  Tile index = (img * nty + ty) * ntx + tx."""
    B: int
    IH: int
    IW: int
    TH: int
    TW: int
    PD: int                      # reflect padding applied on every side before windowing (0: bottom/right overhang)
    ys: List[int]                # unique window origins (padded coordinates)
    xs: List[int]
    owner_y: np.ndarray          # int32[IH]: window row that owns image row y (last writer wins), -1 none
    owner_x: np.ndarray
    kind: str = "direct"

    @property
    def nty(self): return len(self.ys)
    @property
    def ntx(self): return len(self.xs)
    @property
    def tiles_per_image(self): return self.nty * self.ntx
    @property
    def n_tiles(self): return self.B * self.tiles_per_image

    def table(self) -> np.ndarray:
        t = np.zeros((self.n_tiles, 3), dtype=np.int32)
        k = 0
        for b in range(self.B):
            for y0 in self.ys:
                for x0 in self.xs:
                    t[k] = (b, y0, x0)
                    k += 1
        return t


def _owners(n: int, pd: int, starts: Sequence[int], skip: int, padding: int) -> Tuple[List[int], np.ndarray]:
    """Dedupe window origins (a shifted-back last window can coincide with its predecessor; both produce the
    same values) and compute, for each image row, the last window in reference order whose interior
    [s+padding, s+skip-padding) covers it (utils/util.py:144-145: later writes overwrite earlier ones)."""
    uniq: List[int] = []
    for s in starts:
        if s not in uniq:
            uniq.append(s)
    own = np.full(n, -1, dtype=np.int32)
    for s in starts:                       # reference order, later wins
        lo, hi = s + padding - pd, s + skip - padding - pd
        lo, hi = max(lo, 0), min(hi, n)
        if hi > lo:
            own[lo:hi] = uniq.index(s)
    return uniq, own


def geometry_direct(B: int, h: int, w: int, fac: int = 32) -> Geometry:
    """model/ucdir.py:302-307: reflect-pad bottom/right to (h//32+1)*32 (always adds 1..32), run, crop."""
    TH, TW = (h // fac + 1) * fac, (w // fac + 1) * fac
    if TH - h >= h or TW - w >= w:
        raise ValueError("reflect padding %dx%d -> %dx%d needs pad < dim (F.pad raises in the reference too)" % (h, w, TH, TW))
    return Geometry(B, h, w, TH, TW, 0, [0], [0], np.zeros(h, np.int32), np.zeros(w, np.int32), "direct")


def geometry_naive(B: int, h: int, w: int) -> Geometry:
    """model/ucdir.py:270-293 called directly: no padding; four stride-2 stages need h, w % 16 == 0."""
    if h % 16 or w % 16:
        raise ValueError("naiveforward needs H, W multiples of 16 (got %dx%d): skip concat shapes would differ" % (h, w))
    return Geometry(B, h, w, h, w, 0, [0], [0], np.zeros(h, np.int32), np.zeros(w, np.int32), "naive")


def geometry_tiled(B: int, h: int, w: int, skip: int, padding: int) -> Geometry:
    """utils/util.py:108-146."""
    if skip % 16:
        raise ValueError("tile size must be a multiple of 16")
    m = min(h, w)
    pd = skip - m + padding if m < skip else padding
    if pd >= h or pd >= w:
        raise ValueError("tiler reflect pad %d >= image dim %dx%d (F.pad raises in the reference too)" % (pd, h, w))
    ys, oy = _owners(h, pd, tile_windows(h + 2 * pd, skip, padding), skip, padding)
    xs, ox = _owners(w, pd, tile_windows(w + 2 * pd, skip, padding), skip, padding)
    return Geometry(B, h, w, skip, skip, pd, ys, xs, oy, ox, "tiled")


# ======================================================================================
# device buffers
# ======================================================================================
class Pool:
    """Size-bucketed free list over torch allocations.  Ops execute in stream order, so a buffer can be
    handed to a later op as soon as its last reader has been *emitted*."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[int, List[torch.Tensor]] = {}
        self.all: List[torch.Tensor] = []

    def get(self, nbytes: int) -> torch.Tensor:
        nbytes = (int(nbytes) + 255) & ~255
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.all.append(t)
        return t

    def put(self, t: torch.Tensor):
        self.free.setdefault(t.numel(), []).append(t)

    def total_bytes(self):
        return sum(t.numel() for t in self.all)


@dataclass
class Act:
    """An NHWC activation of the tile batch plus its GroupNorm(1,C) statistics slot."""
    buf: torch.Tensor
    C: int
    H: int
    W: int
    stats: int                   # device pointer to double[BT][2] or 0
    keep: bool = False
    split: bool = False          # fp32_tc: (hi, lo) bf16 plane pairs, pixel row = [hi: C | lo: C]

    @property
    def ptr(self): return self.buf.data_ptr()


# ======================================================================================
# weight packing (runs once per load; layouts documented in include/ucdir_b200.h)
# ======================================================================================
def pack_conv_f32(w: torch.Tensor, groups: int = 1, pad_cin_to: Optional[int] = None) -> torch.Tensor:
    """OIHW fp32 -> [groups][K = ks*ks*Cg (tap-major, channel-minor)][ldw = roundup4(Cout/groups)]."""
    co, cg, kh, kw = w.shape
    if pad_cin_to is not None and cg < pad_cin_to:
        w = torch.cat([w, w.new_zeros(co, pad_cin_to - cg, kh, kw)], dim=1)
        cg = pad_cin_to
    ng = co // groups
    ldw = (ng + 3) & ~3
    p = w.view(groups, ng, cg, kh, kw).permute(0, 3, 4, 2, 1).reshape(groups, kh * kw * cg, ng)
    if ldw != ng:
        p = torch.cat([p, p.new_zeros(groups, kh * kw * cg, ldw - ng)], dim=2)
    return p.contiguous().float()


def pack_convT_phase_f32(w: torch.Tensor, py: int, px: int) -> torch.Tensor:
    """ConvTranspose2d(2, stride 2) weight [Cin, Cout, 2, 2] -> the 1x1 GEMM [K=Cin][Cout] of phase (py,px)."""
    return w[:, :, py, px].contiguous().float().unsqueeze(0)


class WeightStore:
    """Flat device copies of every tensor the kernels read, keyed by name."""

    def __init__(self, device):
        self.device = device
        self.t: Dict[str, torch.Tensor] = {}

    def put(self, name: str, t: torch.Tensor) -> int:
        t = t.detach().to(device=self.device).contiguous()
        self.t[name] = t
        return t.data_ptr()

    def ptr(self, name: str) -> int:
        return self.t[name].data_ptr()

    def has(self, name): return name in self.t


# ======================================================================================
# op-list builder
# ======================================================================================
class OpList:
    def __init__(self):
        self.ops: List[_lib.Op] = []
        self._arr = None

    def add(self, kind, p=None, i=None, f=None, flags: int = 0) -> int:
        op = _lib.make_op(kind, p, i, f)
        op.flags = flags
        self.ops.append(op)
        self._arr = None
        return len(self.ops) - 1

    def extend(self, other: "OpList"):
        self.ops.extend(other.ops)
        self._arr = None

    def array(self):
        if self._arr is None:
            arr = (_lib.Op * len(self.ops))()
            for k, o in enumerate(self.ops):
                ctypes.memmove(ctypes.addressof(arr[k]), ctypes.addressof(o), ctypes.sizeof(_lib.Op))
            self._arr = arr
        return self._arr

    def __len__(self): return len(self.ops)


def _conv_op(ol: OpList, *, src0: Act, w: int, dst: Act, cout: int, B: int, src1: Optional[Act] = None,
             bias: int = 0, gamma: int = 0, beta: int = 0, ks: int = 3, stride: int = 1, up: int = 0,
             groups: int = 1, pre: int = 0, act: int = 0, mode: int = 0, res: Optional[Act] = None,
             att: int = 0, attw: int = 0, attw_stride: int = 0, dst_c: Optional[int] = None, dst_coff: int = 0,
             dst_up: int = 0, dst_py: int = 0, dst_px: int = 0, c0: Optional[int] = None, eps: float = 1e-5):
    """Append one UCDIR_OP_CONV_F32.  Output spatial size: dst.H x dst.W (or half of it when dst_up)."""
    H, W = (dst.H // 2, dst.W // 2) if dst_up else (dst.H, dst.W)
    p = {"UCDIR_CONV_P_SRC0": src0.ptr, "UCDIR_CONV_P_W": w, "UCDIR_CONV_P_DST": dst.ptr}
    if src1 is not None:
        p["UCDIR_CONV_P_SRC1"] = src1.ptr
    if bias: p["UCDIR_CONV_P_BIAS"] = bias
    if pre:
        p["UCDIR_CONV_P_GAMMA"], p["UCDIR_CONV_P_BETA"] = gamma, beta
        p["UCDIR_CONV_P_STATS0"] = src0.stats
        if src1 is not None:
            p["UCDIR_CONV_P_STATS1"] = src1.stats
    if res is not None: p["UCDIR_CONV_P_RES"] = res.ptr
    if att: p["UCDIR_CONV_P_ATT"] = att
    if attw: p["UCDIR_CONV_P_ATTW"] = attw
    if dst.stats: p["UCDIR_CONV_P_DST_STATS"] = dst.stats
    i = {"UCDIR_CONV_I_B": B, "UCDIR_CONV_I_H": H, "UCDIR_CONV_I_W": W,
         "UCDIR_CONV_I_C0": src0.C if c0 is None else c0, "UCDIR_CONV_I_C1": src1.C if src1 is not None else 0,
         "UCDIR_CONV_I_COUT": cout, "UCDIR_CONV_I_KSIZE": ks, "UCDIR_CONV_I_STRIDE": stride, "UCDIR_CONV_I_UP": up,
         "UCDIR_CONV_I_GROUPS": groups, "UCDIR_CONV_I_PRE": pre, "UCDIR_CONV_I_ACT": act, "UCDIR_CONV_I_MODE": mode,
         "UCDIR_CONV_I_SRC_H": src0.H, "UCDIR_CONV_I_SRC_W": src0.W,
         "UCDIR_CONV_I_DST_C": dst.C if dst_c is None else dst_c, "UCDIR_CONV_I_DST_COFF": dst_coff,
         "UCDIR_CONV_I_DST_UP": dst_up, "UCDIR_CONV_I_DST_PY": dst_py, "UCDIR_CONV_I_DST_PX": dst_px,
         "UCDIR_CONV_I_RES_C": res.C if res is not None else 0, "UCDIR_CONV_I_ATTW_STRIDE": attw_stride,
         "UCDIR_CONV_I_GN_GROUPS": 1}
    ol.add("UCDIR_OP_CONV_F32", p, i, {"UCDIR_CONV_F_EPS": eps})


# ======================================================================================
# bf16 / tcgen05 path: weight packing (layouts documented in include/ucdir_b200.h, UCDIR_OP_TC_CONV)
# ======================================================================================
BF16 = torch.bfloat16
_CLS_TAPS = [[(ty, tx) for ty in range(3) for tx in range(3)
              if not (cy == 0 and ty == 0) and not (cy == 2 and ty == 2) and not (cx == 0 and tx == 0) and not (cx == 2 and tx == 2)]
             for cy in range(3) for cx in range(3)]
"""taps of a 3x3 pad-1 convolution that fall inside the image, per border class cls = cy*3 + cx
(cy: 0 top row, 1 interior, 2 bottom row; same for cx)."""


def _cls_mask(device) -> torch.Tensor:
    m = torch.zeros(9, 3, 3)
    for cls, taps in enumerate(_CLS_TAPS):
        for (ty, tx) in taps:
            m[cls, ty, tx] = 1.0
    return m.to(device)


def _round_up(v, m):
    return (v + m - 1) // m * m


def split_hi_lo(t: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi + lo == t to 16 mantissa bits (UCDIR_TC_I_SPLIT)."""
    t = t.float()
    hi = t.to(BF16)
    return hi, (t - hi.float()).to(BF16)


def _split_k(wt: torch.Tensor, c0: Optional[int] = None) -> torch.Tensor:
    """[..., taps, ci] fp32 -> [..., taps, 3*ci] bf16 in the K order of a SPLIT op: per tap
    [W_hi(src0) | W_hi(src0) | W_hi(src1) | W_hi(src1) | W_lo(src0) | W_lo(src1)] (src0 = channels [0, c0))."""
    ci = wt.shape[-1]
    c0 = ci if c0 is None else c0
    hi, lo = split_hi_lo(wt)
    return torch.cat([hi[..., :c0], hi[..., :c0], hi[..., c0:], hi[..., c0:], lo[..., :c0], lo[..., c0:]], dim=-1)


def pack_tc_dense(w: torch.Tensor, bias, nt: int, gamma=None, beta=None, split: bool = False, c0: Optional[int] = None):
    """OIHW fp32 conv weight -> (W bf16 [Ntot][K = (tap, c)], TB fp32 [ncls][Ntot], TG or None).
    With gamma/beta the GroupNorm(1,C) in front of the conv is folded (see ucdir_tc.cu header).
    split: fp32_tc operand pairs (three K passes per tap, see _split_k); tables from the fp32 weights."""
    co, ci, kh, kw = w.shape
    w = w.float()
    wg = w * gamma.float().view(1, ci, 1, 1) if gamma is not None else w
    ntot = _round_up(co, nt)
    wq = wg.to(BF16)
    kmul = 3 if split else 1
    packed = torch.zeros(ntot, kh * kw * ci * kmul, dtype=BF16, device=w.device)
    if split:
        packed[:co] = _split_k(wg.permute(0, 2, 3, 1).reshape(co, kh * kw, ci), c0).reshape(co, kh * kw * ci * 3)
    else:
        packed[:co] = wq.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
    b = bias.float() if bias is not None else torch.zeros(co, device=w.device)
    if gamma is None:
        tb = torch.zeros(1, ntot, device=w.device); tb[0, :co] = b
        return packed.contiguous(), tb.contiguous(), None
    wq32 = wg if split else wq.float()
    if kh == 3:
        ncls = 9
        tg = torch.zeros(ncls, ntot, device=w.device); tb = torch.zeros(ncls, ntot, device=w.device)
        wb = (w * beta.float().view(1, ci, 1, 1)).sum(1)           # [co, 3, 3]
        wgs = wq32.sum(1)
        mask = _cls_mask(w.device)
        tg[:, :co] = torch.einsum("cyx,nyx->cn", mask, wgs)
        tb[:, :co] = torch.einsum("cyx,nyx->cn", mask, wb) + b.view(1, co)
    else:
        tg = torch.zeros(1, ntot, device=w.device); tb = torch.zeros(1, ntot, device=w.device)
        tg[0, :co] = wq32.sum((1, 2, 3))
        tb[0, :co] = (w * beta.float().view(1, ci, 1, 1)).sum((1, 2, 3)) + b
    return packed.contiguous(), tb.contiguous(), tg.contiguous()


def pack_tc_grouped(w: torch.Tensor, bias, groups: int, kc: int, gamma, beta, split: bool = False):
    """Grouped 3x3 conv [G*Ng, Cg, 3, 3] (spdyconv, model/ucdir.py:116) with folded GroupNorm.  When Cg < KC
    the K chunk spans KC/Cg neighbouring groups and the rows carry zeros for the foreign channels.
    split: per tap [W_hi | W_hi | W_lo] over the group's chunk (UCDIR_TC_I_SPLIT)."""
    co, cg, kh, kw = w.shape
    cin = cg * groups
    ng = co // groups
    cg_eff = max(cg, kc)
    w = w.float()
    dev = w.device
    packed32 = torch.zeros(co, kh * kw, cg_eff, device=dev)
    tg = torch.zeros(9, co, device=dev); tb = torch.zeros(9, co, device=dev)
    mask = _cls_mask(dev)
    for g in range(groups):
        cbase = (g * cg) // cg_eff * cg_eff
        off = g * cg - cbase
        rows = slice(g * ng, (g + 1) * ng)
        wg = w[rows] * gamma.float()[g * cg:(g + 1) * cg].view(1, cg, 1, 1)
        packed32[rows, :, off:off + cg] = wg.permute(0, 2, 3, 1).reshape(ng, kh * kw, cg)
        wgs = (wg if split else wg.to(BF16).float()).sum(1)
        wb = (w[rows] * beta.float()[g * cg:(g + 1) * cg].view(1, cg, 1, 1)).sum(1)
        tg[:, rows] = torch.einsum("cyx,nyx->cn", mask, wgs)
        tb[:, rows] = torch.einsum("cyx,nyx->cn", mask, wb)
    tb += bias.float().view(1, co)
    if split:
        packed = _split_k(packed32).reshape(co, kh * kw * cg_eff * 3)
    else:
        packed = packed32.to(BF16).reshape(co, kh * kw * cg_eff)
    return packed.contiguous(), tb.contiguous(), tg.contiguous()


def pack_tc_up_phase(w: torch.Tensor, bias, py: int, px: int, nt: int, split: bool = False):
    """Nearest-2x upsample followed by a 3x3 conv (model/ucdir.py:53-60) == four 2x2-tap convolutions on the
    source grid, one per output parity (py, px); row set of tap t: py=0 -> {0}, {1,2}; py=1 -> {0,1}, {2}."""
    co, ci, _, _ = w.shape
    w = w.float()
    rs = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    wp = torch.zeros(co, 2, 2, ci, device=w.device)
    for ty in range(2):
        for tx in range(2):
            acc = 0
            for dy in rs[py][ty]:
                for dx in rs[px][tx]:
                    acc = acc + w[:, :, dy, dx]
            wp[:, ty, tx, :] = acc
    ntot = _round_up(co, nt)
    packed = torch.zeros(ntot, 4 * ci * (3 if split else 1), dtype=BF16, device=w.device)
    packed[:co] = _split_k(wp.reshape(co, 4, ci)).reshape(co, 12 * ci) if split else wp.reshape(co, 4 * ci).to(BF16)
    tb = torch.zeros(1, ntot, device=w.device); tb[0, :co] = bias.float()
    return packed.contiguous(), tb.contiguous()


def _tc_nt(cout: int) -> int:
    for nt in (256, 128, 64):
        if cout % nt == 0:
            return nt
    raise ValueError("no NT tile for Cout=%d" % cout)


def _tc_op(ol: OpList, *, src0: Act, w: int, tb: int, dst: Act, ntot: int, B: int, nt: int, kc: int = 64,
           src1: Optional[Act] = None, tg: int = 0, gn: int = 0, ncls: int = 1, nty: int = 3, ntx: int = 3, oy0: int = -1,
           ox0: int = -1, stride: int = 1, groups: int = 1, act: int = 0, mode: int = 0, res: Optional[Act] = None,
           att: int = 0, attw: int = 0, attw_stride: int = 0, dst_f32: int = 0, ncol_valid: int = 0, dst_up: int = 0,
           dst_py: int = 0, dst_px: int = 0, eps: float = 1e-5, kb: int = 0, nsplit: int = 1, src_cstride: int = 0,
           w_batched: int = 0, w_rowstride: int = 0, w_batchstride: int = 0, alpha: float = 0.0, dst2: int = 0, t_col0: int = 0,
           t_ld: int = 0, w_rows: int = 0, row3: Optional[int] = None, halo: Optional[int] = None, src_gn_swish: int = 0,
           src_gamma: int = 0, src_beta: int = 0, w2: int = 0, tb2: int = 0, dst_res: Optional[Act] = None,
           split: int = 0, src_lo_off: int = 0, w_lo_off: int = 0, dst_crop: int = 0, phases: int = 0, flags: int = 0):
    H, W = (dst.H // 2, dst.W // 2) if dst_up else (dst.H, dst.W)
    p = {"UCDIR_TC_P_SRC0": src0.ptr, "UCDIR_TC_P_W": w, "UCDIR_TC_P_TB": tb, "UCDIR_TC_P_DST": dst.ptr}
    if src1 is not None: p["UCDIR_TC_P_SRC1"] = src1.ptr
    if tg: p["UCDIR_TC_P_TG"] = tg
    if gn:
        p["UCDIR_TC_P_STATS0"] = src0.stats
        if src1 is not None: p["UCDIR_TC_P_STATS1"] = src1.stats
    if res is not None: p["UCDIR_TC_P_RES"] = res.ptr
    if att: p["UCDIR_TC_P_ATT"] = att
    if attw: p["UCDIR_TC_P_ATTW"] = attw
    if dst.stats: p["UCDIR_TC_P_DST_STATS"] = dst.stats
    i = {"UCDIR_TC_I_B": B, "UCDIR_TC_I_H": H, "UCDIR_TC_I_W": W, "UCDIR_TC_I_SRC_H": src0.H, "UCDIR_TC_I_SRC_W": src0.W,
         "UCDIR_TC_I_C0": src0.C, "UCDIR_TC_I_C1": src1.C if src1 is not None else 0, "UCDIR_TC_I_NTOT": ntot,
         "UCDIR_TC_I_NCOL_VALID": ncol_valid, "UCDIR_TC_I_NTY": nty, "UCDIR_TC_I_NTX": ntx, "UCDIR_TC_I_OY0": oy0,
         "UCDIR_TC_I_OX0": ox0, "UCDIR_TC_I_STRIDE": stride, "UCDIR_TC_I_GROUPS": groups, "UCDIR_TC_I_KC": kc,
         "UCDIR_TC_I_NT": nt, "UCDIR_TC_I_GN": gn, "UCDIR_TC_I_NCLS": ncls, "UCDIR_TC_I_ACT": act, "UCDIR_TC_I_MODE": mode,
         "UCDIR_TC_I_DST_F32": dst_f32, "UCDIR_TC_I_DST_C": dst.C, "UCDIR_TC_I_DST_COFF": 0, "UCDIR_TC_I_DST_UP": dst_up,
         "UCDIR_TC_I_DST_PY": dst_py, "UCDIR_TC_I_DST_PX": dst_px, "UCDIR_TC_I_RES_C": res.C if res is not None else 0,
         "UCDIR_TC_I_ATTW_STRIDE": attw_stride, "UCDIR_TC_I_KB": kb or kc, "UCDIR_TC_I_NSPLIT": nsplit,
         "UCDIR_TC_I_SRC_CSTRIDE": src_cstride, "UCDIR_TC_I_W_BATCHED": w_batched, "UCDIR_TC_I_W_ROWSTRIDE": w_rowstride,
         "UCDIR_TC_I_W_BATCHSTRIDE_LO": w_batchstride & 0x7FFFFFFF, "UCDIR_TC_I_W_BATCHSTRIDE_HI": w_batchstride >> 31,
         "UCDIR_TC_I_T_COL0": t_col0, "UCDIR_TC_I_T_LD": t_ld, "UCDIR_TC_I_W_ROWS": w_rows,
         "UCDIR_TC_I_ROW3": _TC_ROW3 if row3 is None else row3, "UCDIR_TC_I_HALO": _TC_HALO if halo is None else halo}
    if split:                                # the halo request stays: the library applies it where a split schedule exists
        i["UCDIR_TC_I_SPLIT"] = 1
        i["UCDIR_TC_I_SRC_LO_OFF"], i["UCDIR_TC_I_W_LO_OFF"] = src_lo_off, w_lo_off
    if dst2: p["UCDIR_TC_P_DST2"] = dst2
    if dst_crop: i["UCDIR_TC_I_DST_CROP"] = dst_crop
    if phases: i["UCDIR_TC_I_PHASES"] = phases
    if src_gn_swish:
        p["UCDIR_TC_P_SRC_GAMMA"], p["UCDIR_TC_P_SRC_BETA"], p["UCDIR_TC_P_STATS0"] = src_gamma, src_beta, src0.stats
        i["UCDIR_TC_I_SRC_GN_SWISH"] = 1
    if dst_res is not None:                  # fused 1x1 res_conv of the same input (halo schedule only)
        p["UCDIR_TC_P_W2"], p["UCDIR_TC_P_TB2"], p["UCDIR_TC_P_DST_RES"] = w2, tb2, dst_res.ptr
        i["UCDIR_TC_I_RES_FUSED"], i["UCDIR_TC_I_DST_RES_C"] = 1, dst_res.C
    ol.add("UCDIR_OP_TC_CONV", p, i, {"UCDIR_TC_F_EPS": eps, "UCDIR_TC_F_ALPHA": alpha}, flags=flags)


_TC_ROW3 = 0 if os.environ.get("UCDIR_TC_ROW3", "1") == "0" else 1
"""Request the shared-activation-row schedule for dense 3x3 convs on 128-pixel-wide rows (the library applies it only
where its preconditions hold).  On unless UCDIR_TC_ROW3=0."""


_TC_FUSE_RES = 0 if os.environ.get("UCDIR_TC_FUSE_RES", "1") == "0" else 1
"""Fuse a block's 1x1 res_conv into its conv1 launch where the halo schedule applies (csrc/ucdir_dhalo.cu, RES).  On unless
UCDIR_TC_FUSE_RES=0."""
_TC_FUSE_PHASES = 0 if os.environ.get("UCDIR_TC_FUSE_PHASES", "1") == "0" else 1
"""One launch for the four phase convolutions of an Upsample layer (UCDIR_TC_I_PHASES).  UCDIR_TC_FUSE_PHASES=0: four launches."""
_TC_FLASH = 0 if os.environ.get("UCDIR_TC_FLASH", "1") == "0" else 1
"""Fused attention core (csrc/ucdir_attn.cu: QK^T -> online softmax -> PV in one kernel, scores / probabilities never in HBM) for
the 512-channel attention layers of the bf16 path.  UCDIR_TC_FLASH=0 keeps the three-launch form (score GEMM, softmax, PV GEMM)."""
_TC_HALO = 0 if os.environ.get("UCDIR_TC_HALO", "1") == "0" else 1
"""Request the halo / weight-stationary schedule (csrc/ucdir_mix.cu) for the integration-module convs with C = 64 / 128 /
256 (the library applies it only where its preconditions hold).  On unless UCDIR_TC_HALO=0."""


def tc_mix_tiling(cout: int):
    """(KC, KB, NT, NSPLIT) of the grouped spdyconv + mix op for C = cout (8 groups of C/8 channels -> 8 x C
    columns): a 256-column work item covers 256/C groups which share one 32-channel activation slab."""
    cg = cout // 8
    if cout >= 512:
        return 64, 64, 256, 1
    return 32, max(cg, 16), 256, max(256 // cout, 1)


class _Builder:
    """Shared helpers of the UNet / predictor graph builders."""

    def __init__(self, pool: Pool, BT: int, stats: torch.Tensor, elem: int = 4, split: bool = False):
        self.pool, self.BT, self.elem = pool, BT, elem * (2 if split else 1)
        self.split = split
        self.stats = stats               # double[n_slots][BT][2]
        self.next_slot = 0
        self.ops = OpList()

    def new(self, C, H, W, with_stats=True, keep=False, cpad: Optional[int] = None) -> Act:
        buf = self.pool.get(self.BT * H * W * (cpad or C) * self.elem)
        st = 0
        if with_stats:
            if self.next_slot >= self.stats.shape[0]:
                raise RuntimeError("statistics arena exhausted")
            st = self.stats.data_ptr() + self.next_slot * self.BT * 16
            self.next_slot += 1
        return Act(buf, cpad or C, H, W, st, keep, self.split)

    def release(self, a: Optional[Act]):
        if a is not None and not a.keep:
            self.pool.put(a.buf)


# ======================================================================================
# UNet engine
# ======================================================================================
MAX_STAT_SLOTS = 192
PRECISIONS = ("fp32", "bf16", "fp32_tc")
"""fp32: SIMT FFMA kernels (debug / bit-level reference of the kernel semantics).  bf16: tcgen05, bf16 operands, stated bf16
tolerance.  fp32_tc: tcgen05 with split operands (x = hi + lo, both bf16; hi*hi + hi*lo + lo*hi accumulated in fp32 TMEM),
meets the reference's fp32 tolerance (rtol 1e-3 / atol 1e-4) at one third of the bf16 tensor rate (SURVEY 8d "Tolerances")."""


def _max_chunk_pixels():
    return int(os.environ.get("UCDIR_CHUNK_PIXELS", 4 * 1024 * 1024 + 512 * 1024))


class UNetEngine:
    """Owns packed weights of one DY3h module and builds/runs forward plans."""

    def __init__(self, module):
        self.m = module
        self.ws: Optional[WeightStore] = None
        self.blocks: List[Tuple[str, object]] = []
        self._sessions: Dict[tuple, "Session"] = {}
        self._geo_cache: Dict[tuple, Geometry] = {}
        self._params: Optional[list] = None
        self._packed_version = -1
        self.precision = os.environ.get("UCDIR_PRECISION", "fp32_tc")    # see PRECISIONS; default = the tensor-core mode that meets the reference's fp32 tolerance
        if self.precision not in PRECISIONS:
            raise ValueError("UCDIR_PRECISION must be one of %s" % (PRECISIONS,))
        self.shard_mode = os.environ.get("UCDIR_SHARD", "none")       # "none" | "tiles" | "batch" 
        if self.shard_mode not in ("none", "tiles", "batch"):
            raise ValueError("UCDIR_SHARD must be none, tiles or batch")

    # ---- weights ------------------------------------------------------------------------
    def invalidate_weights(self):
        self.ws = None
        self._params = None
        self._sessions.clear()

    def invalidate_schedule(self):
        pass                            # levels are per-step kernel arguments; nothing cached per schedule

    def set_shard_mode(self, mode: str):
        if mode not in ("none", "tiles", "batch"):
            raise ValueError("shard mode must be none, tiles or batch")
        if mode != self.shard_mode:
            self.shard_mode = mode
            self._sessions.clear()

    def set_precision(self, precision: str):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (PRECISIONS,))
        if precision != self.precision:
            self.precision = precision
            self.invalidate_weights()

    def device(self):
        return next(self.m.parameters()).device

    def _iter_blocks(self):
        from .model import ucdir as U
        for grp in ("downs", "mid", "ups"):
            for k, layer in enumerate(getattr(self.m, grp)):
                if isinstance(layer, U.ResnetBlocWithAttn):
                    yield "%s.%d" % (grp, k), layer

    def _param_version(self) -> int:
        """Sum of the parameters' in-place version counters: changes when an optimizer step, the reference's EMA update
        (model/model.py:81-91, `p.data.mul_().add_()`) or any other in-place write touches a weight after it was packed."""
        if self._params is None:
            self._params = list(self.m.parameters())
        return sum(p._version for p in self._params)

    def ensure_weights(self):
        dev = self.device()
        _require_cuda(dev)
        ver = self._param_version()
        if self.ws is not None and self.ws.device == dev and ver == self._packed_version:
            return
        if self.ws is not None:
            self._sessions.clear()               # plans and captured graphs point at the old packed weights
        self._packed_version = ver
        ws = WeightStore(dev)
        m = self.m
        inner = m.cfg["inner_channel"]
        ws.put("temb.w1", m.noise_level_mlp[1].weight.float()); ws.put("temb.b1", m.noise_level_mlp[1].bias.float())
        ws.put("temb.w2", m.noise_level_mlp[3].weight.float()); ws.put("temb.b2", m.noise_level_mlp[3].bias.float())
        recs = []
        self.blocks = list(self._iter_blocks())
        for name, layer in self.blocks:
            rb = layer.res_block
            recs.append(torch.cat([rb.noise_func[0].weight.float().reshape(-1), rb.noise_func[0].bias.float(),
                                   rb.noise_func[2].weight.float().reshape(-1), rb.noise_func[2].bias.float()]))
            ws.put(name + ".norm1.w", rb.norm1.weight.float()); ws.put(name + ".norm1.b", rb.norm1.bias.float())
            ws.put(name + ".norm2.w", rb.norm2.weight.float()); ws.put(name + ".norm2.b", rb.norm2.bias.float())
            ws.put(name + ".conv1.w", pack_conv_f32(rb.conv1.weight)); ws.put(name + ".conv1.b", rb.conv1.bias.float())
            ws.put(name + ".spdy.w", pack_conv_f32(rb.spdyconv.weight, groups=rb.nset))
            ws.put(name + ".spdy.b", rb.spdyconv.bias.float())
            ws.put(name + ".g.w0", rb.conv2[0].weight.float().reshape(16, 3)); ws.put(name + ".g.b0", rb.conv2[0].bias.float())
            ws.put(name + ".g.w2", rb.conv2[2].weight.float()); ws.put(name + ".g.b2", rb.conv2[2].bias.float())
            if isinstance(rb.res_conv, torch.nn.Conv2d):
                ws.put(name + ".res.w", pack_conv_f32(rb.res_conv.weight)); ws.put(name + ".res.b", rb.res_conv.bias.float())
            if layer.with_attn:
                at = layer.attn
                ws.put(name + ".attn.norm.w", at.norm.weight.float()); ws.put(name + ".attn.norm.b", at.norm.bias.float())
                ws.put(name + ".attn.qkv.w", pack_conv_f32(at.qkv.weight))
                ws.put(name + ".attn.out.w", pack_conv_f32(at.out.weight)); ws.put(name + ".attn.out.b", at.out.bias.float())
        ws.put("temb.blk", torch.stack(recs))
        from .model import ucdir as U
        for grp in ("downs", "ups"):
            for k, layer in enumerate(getattr(m, grp)):
                name = "%s.%d" % (grp, k)
                if isinstance(layer, torch.nn.Conv2d):          # in-conv: 6 -> 8 zero-padded input channels
                    ws.put(name + ".w", pack_conv_f32(layer.weight, pad_cin_to=8)); ws.put(name + ".b", layer.bias.float())
                elif isinstance(layer, (U.Downsample, U.Upsample)):
                    ws.put(name + ".w", pack_conv_f32(layer.conv.weight)); ws.put(name + ".b", layer.conv.bias.float())
        fc = m.final_conv
        ws.put("final.norm.w", fc[0].weight.float()); ws.put("final.norm.b", fc[0].bias.float())
        ws.put("final.w", pack_conv_f32(fc[3].weight)); ws.put("final.b", fc[3].bias.float())
        self.inner = inner
        self.ws = ws
        # channels of the running activation in front of every block (first source of the up blocks' concatenated input)
        from .model import ucdir as U2
        self._block_c0 = {}
        c = None
        for grp in ("downs", "mid", "ups"):
            for k, layer in enumerate(getattr(m, grp)):
                if isinstance(layer, torch.nn.Conv2d):
                    c = layer.out_channels
                elif isinstance(layer, U2.ResnetBlocWithAttn):
                    self._block_c0["%s.%d" % (grp, k)] = c
                    c = layer.res_block.dim_out
        if self.precision in ("bf16", "fp32_tc"):
            self._ensure_weights_bf16(split=self.precision == "fp32_tc")

    # ---- graph --------------------------------------------------------------------------
    def n_blocks(self):
        return len(self.blocks)

    def build_guidance_ops(self, ol: OpList, pool: Pool, BT: int, TH: int, TW: int, guide_tiles: torch.Tensor
                           ) -> List[torch.Tensor]:
        """27 step-invariant guidance maps [BT, H, W, 8] (model/ucdir.py:133-135 without attw), one per block."""
        from .model import ucdir as U
        ws = self.ws
        maps = []
        res = {}                                                 # block name -> (H, W)
        H, W = TH, TW
        for k, layer in enumerate(self.m.downs):
            if isinstance(layer, U.Downsample):
                H, W = H // 2, W // 2
            elif isinstance(layer, U.ResnetBlocWithAttn):
                res["downs.%d" % k] = (H, W)
        for k, _ in enumerate(self.m.mid):
            res["mid.%d" % k] = (H, W)
        for k, layer in enumerate(self.m.ups):
            if isinstance(layer, U.Upsample):
                H, W = H * 2, W * 2
            else:
                res["ups.%d" % k] = (H, W)
        for name, _ in self.blocks:
            h, w = res[name]
            t = torch.empty(BT * h * w * 8, dtype=F32, device=pool.device)
            maps.append(t)
            ol.add("UCDIR_OP_GUIDANCE",
                   {"UCDIR_GUID_P_GUIDE": guide_tiles.data_ptr(), "UCDIR_GUID_P_W0": ws.ptr(name + ".g.w0"),
                    "UCDIR_GUID_P_B0": ws.ptr(name + ".g.b0"), "UCDIR_GUID_P_W2": ws.ptr(name + ".g.w2"),
                    "UCDIR_GUID_P_B2": ws.ptr(name + ".g.b2"), "UCDIR_GUID_P_DST": t.data_ptr()},
                   {"UCDIR_GUID_I_B": BT, "UCDIR_GUID_I_GH": TH, "UCDIR_GUID_I_GW": TW, "UCDIR_GUID_I_H": h,
                    "UCDIR_GUID_I_W": w})
        return maps

    def time_embed_op(self, ol: OpList, dst: torch.Tensor, L: int, levels_ptr: int = 0, level: float = 0.0) -> int:
        ws = self.ws
        return ol.add("UCDIR_OP_TIME_EMBED",
                      {"UCDIR_TEMB_P_LEVELS": levels_ptr, "UCDIR_TEMB_P_W1": ws.ptr("temb.w1"),
                       "UCDIR_TEMB_P_B1": ws.ptr("temb.b1"), "UCDIR_TEMB_P_W2": ws.ptr("temb.w2"),
                       "UCDIR_TEMB_P_B2": ws.ptr("temb.b2"), "UCDIR_TEMB_P_BLK": ws.ptr("temb.blk"),
                       "UCDIR_TEMB_P_DST": dst.data_ptr()},
                      {"UCDIR_TEMB_I_L": L, "UCDIR_TEMB_I_NBLK": len(self.blocks), "UCDIR_TEMB_I_INNER": self.inner},
                      {"UCDIR_TEMB_F_LEVEL": level})

    def build_forward_ops(self, pool: Pool, BT: int, TH: int, TW: int, x_in: torch.Tensor, gmaps: List[torch.Tensor],
                          attw: torch.Tensor, attw_stride: int, eps_ptr: int, stats: torch.Tensor) -> OpList:
        """DY3h.naiveforward (model/ucdir.py:270-293) over the tile batch x_in[BT,TH,TW,8] -> eps[BT,TH,TW,4]."""
        from .model import ucdir as U
        ws, m = self.ws, self.m
        bld = _Builder(pool, BT, stats)
        ol = bld.ops
        nbytes = stats.numel() * stats.element_size()
        ol.add("UCDIR_OP_MEMSET", {0: stats.data_ptr()}, {0: nbytes & 0x7FFFFFFF, 1: nbytes >> 31})
        blk_index = {name: k for k, (name, _) in enumerate(self.blocks)}
        nblk = len(self.blocks)

        def block(name, layer, x: Act, skip: Optional[Act]) -> Act:
            rb = layer.res_block
            cout = rb.dim_out
            k = blk_index[name]
            h1 = bld.new(cout, x.H, x.W)
            _conv_op(ol, src0=x, src1=skip, w=ws.ptr(name + ".conv1.w"), bias=ws.ptr(name + ".conv1.b"),
                     gamma=ws.ptr(name + ".norm1.w"), beta=ws.ptr(name + ".norm1.b"), pre=1, act=1, dst=h1, cout=cout, B=BT)
            if ws.has(name + ".res.w"):
                res = bld.new(cout, x.H, x.W, with_stats=False)
                _conv_op(ol, src0=x, src1=skip, w=ws.ptr(name + ".res.w"), bias=ws.ptr(name + ".res.b"), ks=1, dst=res,
                         cout=cout, B=BT)
                own_res = True
            else:
                if skip is not None:
                    raise RuntimeError("identity residual with a concatenated input")
                res, own_res = x, False
            out = bld.new(cout, x.H, x.W)
            _conv_op(ol, src0=h1, w=ws.ptr(name + ".spdy.w"), bias=ws.ptr(name + ".spdy.b"),
                     gamma=ws.ptr(name + ".norm2.w"), beta=ws.ptr(name + ".norm2.b"), pre=1, groups=rb.nset, mode=1,
                     att=gmaps[k].data_ptr(), attw=attw.data_ptr() + k * 8 * 4, attw_stride=attw_stride, res=res,
                     dst=out, cout=cout * rb.nset, B=BT, dst_c=cout)
            bld.release(h1)
            if own_res:
                bld.release(res)
            bld.release(x)
            bld.release(skip)
            if layer.with_attn:
                out = attention(name, out)
            return out

        def attention(name, x: Act) -> Act:
            """SelfAttention.forward, model/ucdir.py:165-182 (n_head = 1, d = C)."""
            C, N = x.C, x.H * x.W
            qkv = bld.new(3 * C, x.H, x.W, with_stats=False)
            _conv_op(ol, src0=x, w=ws.ptr(name + ".attn.qkv.w"), gamma=ws.ptr(name + ".attn.norm.w"),
                     beta=ws.ptr(name + ".attn.norm.b"), pre=1, ks=1, dst=qkv, cout=3 * C, B=BT)
            if N * N >= 2 ** 31 or N * 3 * C >= 2 ** 31:
                raise RuntimeError("attention over %d tokens exceeds the fp32 path's 32-bit strides" % N)
            S = pool.get(BT * N * N * 4)
            ol.add("UCDIR_OP_SGEMM_F32",
                   {"UCDIR_SGEMM_P_A": qkv.ptr, "UCDIR_SGEMM_P_B": qkv.ptr + C * 4, "UCDIR_SGEMM_P_C": S.data_ptr()},
                   {"UCDIR_SGEMM_I_BATCH": BT, "UCDIR_SGEMM_I_M": N, "UCDIR_SGEMM_I_N": N, "UCDIR_SGEMM_I_K": C,
                    "UCDIR_SGEMM_I_LDA": 3 * C, "UCDIR_SGEMM_I_LDB": 3 * C, "UCDIR_SGEMM_I_LDC": N,
                    "UCDIR_SGEMM_I_SA": N * 3 * C, "UCDIR_SGEMM_I_SB": N * 3 * C, "UCDIR_SGEMM_I_SC": N * N,
                    "UCDIR_SGEMM_I_TRANSB": 1},
                   {"UCDIR_SGEMM_F_ALPHA": 1.0 / math.sqrt(C)})
            ol.add("UCDIR_OP_SOFTMAX_F32", {"UCDIR_SOFTMAX_P_X": S.data_ptr()},
                   {"UCDIR_SOFTMAX_I_ROWS": BT * N, "UCDIR_SOFTMAX_I_COLS": N})
            o = bld.new(C, x.H, x.W, with_stats=False)
            ol.add("UCDIR_OP_SGEMM_F32",
                   {"UCDIR_SGEMM_P_A": S.data_ptr(), "UCDIR_SGEMM_P_B": qkv.ptr + 2 * C * 4, "UCDIR_SGEMM_P_C": o.ptr},
                   {"UCDIR_SGEMM_I_BATCH": BT, "UCDIR_SGEMM_I_M": N, "UCDIR_SGEMM_I_N": C, "UCDIR_SGEMM_I_K": N,
                    "UCDIR_SGEMM_I_LDA": N, "UCDIR_SGEMM_I_LDB": 3 * C, "UCDIR_SGEMM_I_LDC": C,
                    "UCDIR_SGEMM_I_SA": N * N, "UCDIR_SGEMM_I_SB": N * 3 * C, "UCDIR_SGEMM_I_SC": N * C,
                    "UCDIR_SGEMM_I_TRANSB": 0},
                   {"UCDIR_SGEMM_F_ALPHA": 1.0})
            pool.put(S)
            bld.release(qkv)
            y = bld.new(C, x.H, x.W)
            _conv_op(ol, src0=o, w=ws.ptr(name + ".attn.out.w"), bias=ws.ptr(name + ".attn.out.b"), ks=1, res=x, dst=y,
                     cout=C, B=BT)
            bld.release(o)
            bld.release(x)
            return y

        feats: List[Act] = []
        x = Act(x_in, 8, TH, TW, 0, keep=True)
        for k, layer in enumerate(m.downs):
            name = "downs.%d" % k
            if isinstance(layer, torch.nn.Conv2d):
                y = bld.new(layer.out_channels, x.H, x.W)
                _conv_op(ol, src0=x, w=ws.ptr(name + ".w"), bias=ws.ptr(name + ".b"), dst=y, cout=layer.out_channels, B=BT)
                x = y
            elif isinstance(layer, U.Downsample):
                y = bld.new(x.C, x.H // 2, x.W // 2)
                _conv_op(ol, src0=x, w=ws.ptr(name + ".w"), bias=ws.ptr(name + ".b"), stride=2, dst=y, cout=x.C, B=BT)
                x = y
            else:
                x.keep = True                                    # the input of a down block is always a stored feature
                x = block(name, layer, x, None)
            x.keep = True
            feats.append(x)
        x = feats[-1]
        for k, layer in enumerate(m.mid):
            x = block("mid.%d" % k, layer, x, None)              # first input is feats[-1] (kept); later ones are released
        for k, layer in enumerate(m.ups):
            name = "ups.%d" % k
            if isinstance(layer, U.Upsample):
                y = bld.new(x.C, x.H * 2, x.W * 2)
                _conv_op(ol, src0=x, w=ws.ptr(name + ".w"), bias=ws.ptr(name + ".b"), up=1, dst=y, cout=x.C, B=BT)
                bld.release(x)
                x = y
            else:
                skip = feats.pop()
                skip.keep = False
                x = block(name, layer, x, skip)                  # torch.cat((x, feats.pop())) is never materialised
        eps_dst = Act(_PtrBuf(eps_ptr), 4, TH, TW, 0, keep=True)      # type: ignore[arg-type]
        _conv_op(ol, src0=x, w=ws.ptr("final.w"), bias=ws.ptr("final.b"), gamma=ws.ptr("final.norm.w"),
                 beta=ws.ptr("final.norm.b"), pre=2, dst=eps_dst, cout=m.cfg["out_channel"], B=BT)
        bld.release(x)
        self.last_stat_slots = bld.next_slot
        return ol

    def attw_rows(self, levels: torch.Tensor) -> torch.Tensor:
        """The per-block timestep weights attw[L][n_blocks][8] (PositionalEncoding + noise_level_mlp + every block's
        noise_func, model/ucdir.py:24-29,212-214,106,125) for L noise levels in ONE launch.  The level of step t is a
        function of the schedule only (model/diffusion.py:162-163), so the samplers build this table once per schedule and a
        step copies its row next to the other per-step scalars: the denoising step itself has no timestep-embedding launch."""
        self.ensure_weights()
        dev = self.device()
        levels = levels.detach().to(device=dev, dtype=F32).contiguous()
        L = int(levels.numel())
        out = torch.empty((L, len(self.blocks), 8), dtype=F32, device=dev)
        ol = OpList()
        self.time_embed_op(ol, out, L, levels.data_ptr(), 0.0)
        with _device_guard(dev):
            _run_ops(ol.array(), len(ol), _stream(dev))
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()        # `levels` may be a temporary
        return out

    # ---- bf16 / tcgen05 graph --------------------------------------------------------------
    def _ensure_weights_bf16(self, split: bool = False):
        """Packed bf16 operands + fp32 epilogue tables for every TC op (once per load).  split: fp32_tc operand pairs."""
        from .model import ucdir as U
        ws, m = self.ws, self.m
        sp = dict(split=True) if split else {}

        def c0_of(name, layer):
            """channels of the first source of a block's conv1 / res_conv (the up blocks read torch.cat((x, skip)))."""
            return self._block_c0.get(name)

        def put3(name, triple):
            w, tb, tg = triple
            ws.put(name + ".tcw", w); ws.put(name + ".tb", tb)
            if tg is not None:
                ws.put(name + ".tg", tg)

        for name, layer in self.blocks:
            rb = layer.res_block
            cout = rb.dim_out
            c0 = dict(c0=c0_of(name, layer)) if split else {}
            put3(name + ".conv1", pack_tc_dense(rb.conv1.weight, rb.conv1.bias, _tc_nt(cout), rb.norm1.weight, rb.norm1.bias, **sp, **c0))
            put3(name + ".spdy", pack_tc_grouped(rb.spdyconv.weight, rb.spdyconv.bias, rb.nset, tc_mix_tiling(cout)[1],
                                                 rb.norm2.weight, rb.norm2.bias, **sp))
            if isinstance(rb.res_conv, torch.nn.Conv2d):
                put3(name + ".res", pack_tc_dense(rb.res_conv.weight, rb.res_conv.bias, _tc_nt(cout), **sp, **c0))
            if layer.with_attn:
                at = layer.attn
                put3(name + ".attn.qkv", pack_tc_dense(at.qkv.weight, None, 256, at.norm.weight, at.norm.bias, **sp))
                put3(name + ".attn.out", pack_tc_dense(at.out.weight, at.out.bias, _tc_nt(at.out.out_channels), **sp))
        for grp in ("downs", "ups"):
            for k, layer in enumerate(getattr(m, grp)):
                name = "%s.%d" % (grp, k)
                if isinstance(layer, torch.nn.Conv2d):          # in-conv: 6 -> 16 zero-padded input channels (KC = 16)
                    w = layer.weight
                    w = torch.cat([w, w.new_zeros(w.shape[0], 16 - w.shape[1], 3, 3)], dim=1)
                    put3(name, pack_tc_dense(w, layer.bias, _tc_nt(layer.out_channels), **sp))
                elif isinstance(layer, U.Downsample):
                    put3(name, pack_tc_dense(layer.conv.weight, layer.conv.bias, _tc_nt(layer.conv.out_channels), **sp))
                elif isinstance(layer, U.Upsample):
                    blocks = []
                    for py in range(2):
                        for px in range(2):
                            w, tb = pack_tc_up_phase(layer.conv.weight, layer.conv.bias, py, px, _tc_nt(layer.conv.out_channels), **sp)
                            ws.put("%s.p%d%d.tcw" % (name, py, px), w); ws.put("%s.p%d%d.tb" % (name, py, px), tb)
                            blocks.append(w)
                    ws.put(name + ".ph.tcw", torch.cat(blocks, dim=0))       # the four phases stacked along N: one launch (PHASES = 4)
        fc = m.final_conv
        put3("final", pack_tc_dense(fc[3].weight, fc[3].bias, 16, **sp))
        ws.put("zeros", torch.zeros(65536, dtype=F32, device=ws.device))    # bias table of the attention GEMMs

    def build_forward_ops_bf16(self, pool: Pool, BT: int, TH: int, TW: int, x_in: torch.Tensor, gmaps: List[torch.Tensor],
                               attw: torch.Tensor, attw_stride: int, eps_ptr: int, stats: torch.Tensor, split: bool = False,
                               eps_crop: int = 0) -> OpList:
        """Same graph as build_forward_ops on the tcgen05 path: bf16 NHWC activations, x_in[BT,TH,TW,16] bf16.
        split (precision fp32_tc): every activation is a (hi, lo) bf16 plane pair and every op a three-pass split-operand
        MMA on the streamed kernel (include/ucdir_b200.h, UCDIR_TC_I_SPLIT); x_in[BT,TH,TW,2*16]."""
        from .model import ucdir as U
        ws, m = self.ws, self.m
        bld = _Builder(pool, BT, stats, elem=2, split=split)
        sp = 1 if split else 0
        ol = bld.ops
        nbytes = stats.numel() * stats.element_size()
        ol.add("UCDIR_OP_MEMSET", {0: stats.data_ptr()}, {0: nbytes & 0x7FFFFFFF, 1: nbytes >> 31})
        blk_index = {name: k for k, (name, _) in enumerate(self.blocks)}

        def block(name, layer, x: Act, skip: Optional[Act]) -> Act:
            rb = layer.res_block
            cout = rb.dim_out
            k = blk_index[name]
            nt = _tc_nt(cout)
            h1 = bld.new(cout, x.H, x.W)
            has_res = ws.has(name + ".res.tcw")
            # the halo schedule computes the 1x1 res_conv from the activation box conv1 already holds in shared memory
            # (fp32_tc too: the split form of the halo kernel visits the hi and lo boxes of every chunk)
            fuse_res = (has_res and _TC_HALO and _TC_FUSE_RES and cout in (64, 128) and x.C % 64 == 0 and
                        (skip is None or skip.C % 64 == 0))
            res = bld.new(cout, x.H, x.W, with_stats=False) if has_res else None
            # a separate 1x1 res_conv reads the same input as conv1 and is first read by the integration conv: it is emitted as a
            # BRANCH of the step graph (captured on a forked stream, joined at the integration conv) -- at the coarse levels of a small
            # tile batch neither kernel fills the GPU
            branch = has_res and not fuse_res
            if branch:
                _tc_op(ol, split=sp, src0=x, src1=skip, w=ws.ptr(name + ".res.tcw"), tb=ws.ptr(name + ".res.tb"), nty=1, ntx=1, oy0=0,
                       ox0=0, dst=res, ntot=cout, B=BT, nt=nt, flags=_lib.C["UCDIR_OP_FLAG_BRANCH"])
            _tc_op(ol, split=sp, src0=x, src1=skip, w=ws.ptr(name + ".conv1.tcw"), tb=ws.ptr(name + ".conv1.tb"),
                   tg=ws.ptr(name + ".conv1.tg"), gn=1, ncls=9, act=1, dst=h1, ntot=cout, B=BT, nt=nt,
                   **(dict(w2=ws.ptr(name + ".res.tcw"), tb2=ws.ptr(name + ".res.tb"), dst_res=res) if fuse_res else {}))
            if has_res:
                own_res = True
            else:
                if skip is not None:
                    raise RuntimeError("identity residual with a concatenated input")
                res, own_res = x, False
            out = bld.new(cout, x.H, x.W)
            kc, kb, mnt, nsplit = tc_mix_tiling(cout)
            _tc_op(ol, split=sp, src0=h1, w=ws.ptr(name + ".spdy.tcw"), tb=ws.ptr(name + ".spdy.tb"), tg=ws.ptr(name + ".spdy.tg"),
                   gn=1, ncls=9, groups=rb.nset, kc=kc, kb=kb, nsplit=nsplit, nt=mnt, mode=1, att=gmaps[k].data_ptr(),
                   attw=attw.data_ptr() + k * 8 * 4, attw_stride=attw_stride, res=res, dst=out, ntot=cout * rb.nset, B=BT,
                   flags=_lib.C["UCDIR_OP_FLAG_JOIN"] if branch else 0)
            bld.release(h1)
            if own_res:
                bld.release(res)
            bld.release(x)
            bld.release(skip)
            if layer.with_attn:
                out = attention(name, out)
            return out

        def attention(name, x: Act) -> Act:
            """SelfAttention.forward (model/ucdir.py:165-182) on the tensor cores: the two einsums are batched GEMMs
            of the same tcgen05 kernel -- Q (a channel slice of the qkv conv's output) is the activation operand,
            K resp. V^T are per-image "weights" (3-D tensor map); V^T is written by the qkv conv's epilogue."""
            C, N = x.C, x.H * x.W
            if split:
                return attention_split(name, x)
            NP = (N + 7) & ~7                                   # 16-byte aligned row pitch of P and V^T
            nt_s = 256 if N % 256 == 0 else (128 if N % 128 == 0 else 64)
            NS = (N + nt_s - 1) // nt_s * nt_s                  # score columns incl. zero padding
            if N * 2 * C >= 2 ** 31 or C * NP >= 2 ** 62:
                raise RuntimeError("attention over %d tokens exceeds the tensor-map strides" % N)
            qk = bld.new(2 * C, x.H, x.W, with_stats=False)
            vt = pool.get(BT * C * NP * 2)
            _tc_op(ol, split=sp, src0=x, w=ws.ptr(name + ".attn.qkv.tcw"), tb=ws.ptr(name + ".attn.qkv.tb"),
                   tg=ws.ptr(name + ".attn.qkv.tg"), gn=1, ncls=1, nty=1, ntx=1, oy0=0, ox0=0, dst=qk, ntot=3 * C, B=BT, nt=256,
                   dst2=vt.data_ptr(), t_col0=2 * C, t_ld=NP)
            if _TC_FLASH and C == 512:
                o = bld.new(C, x.H, x.W, with_stats=False)
                ol.add("UCDIR_OP_TC_ATTN", {"UCDIR_ATTN_P_QK": qk.ptr, "UCDIR_ATTN_P_VT": vt.data_ptr(), "UCDIR_ATTN_P_O": o.ptr},
                       {"UCDIR_ATTN_I_B": BT, "UCDIR_ATTN_I_N": N, "UCDIR_ATTN_I_C": C, "UCDIR_ATTN_I_QK_LD": 2 * C,
                        "UCDIR_ATTN_I_VT_LD": NP, "UCDIR_ATTN_I_O_LD": C}, {"UCDIR_ATTN_F_SCALE": 1.0 / math.sqrt(C)})
                pool.put(vt)
                bld.release(qk)
                y = bld.new(C, x.H, x.W)
                _tc_op(ol, split=sp, src0=o, w=ws.ptr(name + ".attn.out.tcw"), tb=ws.ptr(name + ".attn.out.tb"), nty=1, ntx=1, oy0=0,
                       ox0=0, res=x, dst=y, ntot=C, B=BT, nt=_tc_nt(C))
                bld.release(o)
                bld.release(x)
                return y
            S = pool.get(BT * N * NS * 4)
            zeros = ws.ptr("zeros")
            q_act = Act(qk.buf, C, x.H, x.W, 0, keep=True)
            s_dst = Act(S, NS, x.H, x.W, 0, keep=True)
            _tc_op(ol, split=sp, src0=q_act, src_cstride=2 * C, w=qk.ptr + C * 2, w_batched=1, w_rowstride=2 * C, w_batchstride=N * 2 * C,
                   w_rows=N, tb=zeros, nty=1, ntx=1, oy0=0, ox0=0, dst=s_dst, ntot=NS, B=BT, nt=nt_s, dst_f32=1, ncol_valid=NS,
                   alpha=1.0 / math.sqrt(C))
            P = pool.get(BT * N * NP * 2)
            ol.add("UCDIR_OP_SOFTMAX_F32", {"UCDIR_SOFTMAX_P_X": S.data_ptr(), "UCDIR_SOFTMAX_P_OUT_BF16": P.data_ptr()},
                   {"UCDIR_SOFTMAX_I_ROWS": BT * N, "UCDIR_SOFTMAX_I_COLS": N, "UCDIR_SOFTMAX_I_IN_LD": NS, "UCDIR_SOFTMAX_I_OUT_LD": NP})
            o = bld.new(C, x.H, x.W, with_stats=False)
            p_act = Act(P, N, x.H, x.W, 0, keep=True)
            _tc_op(ol, split=sp, src0=p_act, src_cstride=NP, w=vt.data_ptr(), w_batched=1, w_rowstride=NP, w_batchstride=C * NP, tb=zeros,
                   nty=1, ntx=1, oy0=0, ox0=0, dst=o, ntot=C, B=BT, nt=_tc_nt(C))
            pool.put(S); pool.put(P); pool.put(vt)
            bld.release(qk)
            y = bld.new(C, x.H, x.W)
            _tc_op(ol, split=sp, src0=o, w=ws.ptr(name + ".attn.out.tcw"), tb=ws.ptr(name + ".attn.out.tb"), nty=1, ntx=1, oy0=0, ox0=0,
                   res=x, dst=y, ntot=C, B=BT, nt=_tc_nt(C))
            bld.release(o)
            bld.release(x)
            return y

        def attention_split(name, x: Act) -> Act:
            """The same chain with (hi, lo) operand pairs: qk rows [q_hi k_hi | q_lo k_lo], V^T rows [hi: NP | lo: NP]
            (zeroed first: the K tail of P x V^T must not meet stale bytes), fp32 scores, fp32 softmax in place, probabilities
            split into plane pairs by UCDIR_OP_CAST mode 2."""
            C, N = x.C, x.H * x.W
            NP = _round_up(N, 64)                               # plane pitch of P and V^T: whole K chunks, zero padded
            nt_s = 256 if N % 256 == 0 else (128 if N % 128 == 0 else 64)
            NS = _round_up(N, nt_s)
            if N * 4 * C >= 2 ** 31 or N * NS >= 2 ** 62:
                raise RuntimeError("attention over %d tokens exceeds the tensor-map strides" % N)
            qk = bld.new(2 * C, x.H, x.W, with_stats=False)      # physical rows of 4C
            vt = pool.get(BT * C * 2 * NP * 2)
            nb = BT * C * 2 * NP * 2
            ol.add("UCDIR_OP_MEMSET", {0: vt.data_ptr()}, {0: nb & 0x7FFFFFFF, 1: nb >> 31})
            _tc_op(ol, split=1, src0=x, w=ws.ptr(name + ".attn.qkv.tcw"), tb=ws.ptr(name + ".attn.qkv.tb"),
                   tg=ws.ptr(name + ".attn.qkv.tg"), gn=1, ncls=1, nty=1, ntx=1, oy0=0, ox0=0, dst=qk, ntot=3 * C, B=BT, nt=256,
                   dst2=vt.data_ptr(), t_col0=2 * C, t_ld=NP)
            S = pool.get(BT * N * NS * 4)
            zeros = ws.ptr("zeros")
            q_act = Act(qk.buf, C, x.H, x.W, 0, keep=True, split=True)
            s_dst = Act(S, NS, x.H, x.W, 0, keep=True)
            _tc_op(ol, split=1, src0=q_act, src_cstride=4 * C, src_lo_off=2 * C, w=qk.ptr + C * 2, w_batched=1, w_rowstride=4 * C,
                   w_lo_off=2 * C, w_batchstride=N * 4 * C, w_rows=N, tb=zeros, nty=1, ntx=1, oy0=0, ox0=0, dst=s_dst, ntot=NS, B=BT,
                   nt=nt_s, dst_f32=1, ncol_valid=NS, alpha=1.0 / math.sqrt(C))
            ol.add("UCDIR_OP_SOFTMAX_F32", {"UCDIR_SOFTMAX_P_X": S.data_ptr()},
                   {"UCDIR_SOFTMAX_I_ROWS": BT * N, "UCDIR_SOFTMAX_I_COLS": N, "UCDIR_SOFTMAX_I_IN_LD": NS})
            P = pool.get(BT * N * 2 * NP * 2)
            rows = BT * N
            ol.add("UCDIR_OP_CAST", {0: S.data_ptr(), 1: P.data_ptr()}, {0: rows & 0x7FFFFFFF, 1: rows >> 31, 2: 2, 3: N, 4: NS, 5: NP})
            o = bld.new(C, x.H, x.W, with_stats=False)
            p_act = Act(P, N, x.H, x.W, 0, keep=True, split=True)
            _tc_op(ol, split=1, src0=p_act, src_cstride=2 * NP, src_lo_off=NP, w=vt.data_ptr(), w_batched=1, w_rowstride=2 * NP,
                   w_lo_off=NP, w_batchstride=C * 2 * NP, tb=zeros, nty=1, ntx=1, oy0=0, ox0=0, dst=o, ntot=C, B=BT, nt=_tc_nt(C))
            pool.put(S); pool.put(P); pool.put(vt)
            bld.release(qk)
            y = bld.new(C, x.H, x.W)
            _tc_op(ol, split=1, src0=o, w=ws.ptr(name + ".attn.out.tcw"), tb=ws.ptr(name + ".attn.out.tb"), nty=1, ntx=1, oy0=0, ox0=0,
                   res=x, dst=y, ntot=C, B=BT, nt=_tc_nt(C))
            bld.release(o)
            bld.release(x)
            return y

        feats: List[Act] = []
        x = Act(x_in, 16, TH, TW, 0, keep=True, split=split)
        for k, layer in enumerate(m.downs):
            name = "downs.%d" % k
            if isinstance(layer, torch.nn.Conv2d):
                y = bld.new(layer.out_channels, x.H, x.W)
                _tc_op(ol, split=sp, src0=x, w=ws.ptr(name + ".tcw"), tb=ws.ptr(name + ".tb"), kc=16, dst=y, ntot=layer.out_channels,
                       B=BT, nt=_tc_nt(layer.out_channels))
                x = y
            elif isinstance(layer, U.Downsample):
                y = bld.new(x.C, x.H // 2, x.W // 2)
                _tc_op(ol, split=sp, src0=x, w=ws.ptr(name + ".tcw"), tb=ws.ptr(name + ".tb"), stride=2, dst=y, ntot=x.C, B=BT,
                       nt=_tc_nt(x.C))
                x = y
            else:
                x.keep = True
                x = block(name, layer, x, None)
            x.keep = True
            feats.append(x)
        x = feats[-1]
        for k, layer in enumerate(m.mid):
            x = block("mid.%d" % k, layer, x, None)
        for k, layer in enumerate(m.ups):
            name = "ups.%d" % k
            if isinstance(layer, U.Upsample):
                y = bld.new(x.C, x.H * 2, x.W * 2)
                if _TC_FUSE_PHASES:
                    _tc_op(ol, split=sp, src0=x, w=ws.ptr(name + ".ph.tcw"), tb=ws.ptr(name + ".p00.tb"), nty=2, ntx=2, oy0=-1, ox0=-1,
                           dst=y, ntot=x.C, B=BT, nt=_tc_nt(x.C), dst_up=1, phases=4)
                else:
                    for py in range(2):
                        for px in range(2):
                            _tc_op(ol, split=sp, src0=x, w=ws.ptr("%s.p%d%d.tcw" % (name, py, px)), tb=ws.ptr("%s.p%d%d.tb" % (name, py, px)),
                                   nty=2, ntx=2, oy0=py - 1, ox0=px - 1, dst=y, ntot=x.C, B=BT, nt=_tc_nt(x.C), dst_up=1,
                                   dst_py=py, dst_px=px)
                bld.release(x)
                x = y
            else:
                skip = feats.pop()
                skip.keep = False
                x = block(name, layer, x, skip)
        # final_conv: GN -> Swish -> conv (model/ucdir.py:266-268); Swish sits between the norm and the conv, so the norm
        # cannot be folded into the weights.  Halo schedule: ONE kernel applies GN + Swish to the landed activation box in
        # shared memory and runs the 16-column (3 valid) conv to fp32 eps (csrc/ucdir_fhalo.cu).  Otherwise: one elementwise
        # pass, then the streamed conv.
        eps_dst = Act(_PtrBuf(eps_ptr), 4, TH, TW, 0, keep=True)      # type: ignore[arg-type]
        if eps_crop and not self.fused_final_crop(split, x.C):
            raise RuntimeError("eps_crop needs the fused final conv")
        if _TC_HALO and not split and x.C % 64 == 0 and x.C <= 128:
            _tc_op(ol, split=sp, src0=x, w=ws.ptr("final.tcw"), tb=ws.ptr("final.tb"), dst=eps_dst, ntot=16, B=BT, nt=16, dst_f32=1,
                   ncol_valid=m.cfg["out_channel"], src_gn_swish=1, src_gamma=ws.ptr("final.norm.w"), src_beta=ws.ptr("final.norm.b"),
                   dst_crop=eps_crop)
            bld.release(x)
        else:
            xn = bld.new(x.C, x.H, x.W, with_stats=False)
            ol.add("UCDIR_OP_GN_APPLY_BF16",
                   {"UCDIR_GNA_P_SRC": x.ptr, "UCDIR_GNA_P_DST": xn.ptr, "UCDIR_GNA_P_GAMMA": ws.ptr("final.norm.w"),
                    "UCDIR_GNA_P_BETA": ws.ptr("final.norm.b"), "UCDIR_GNA_P_STATS": x.stats},
                   {"UCDIR_GNA_I_B": BT, "UCDIR_GNA_I_HW": x.H * x.W, "UCDIR_GNA_I_C": x.C, "UCDIR_GNA_I_SWISH": 1,
                    "UCDIR_GNA_I_SPLIT": sp}, {0: 1e-5})
            bld.release(x)
            _tc_op(ol, split=sp, src0=xn, w=ws.ptr("final.tcw"), tb=ws.ptr("final.tb"), dst=eps_dst, ntot=16, B=BT, nt=16, dst_f32=1,
                   ncol_valid=m.cfg["out_channel"])
            bld.release(xn)
        self.last_stat_slots = bld.next_slot
        return ol

    def fused_final_crop(self, split: bool, c_final: Optional[int] = None) -> bool:
        """True when final_conv runs as the fused halo kernel, which can write the tile interiors straight into the compact eps
        buffer (UCDIR_TC_I_DST_CROP) -- no separate crop pass."""
        c = self.m.prec if c_final is None else c_final
        return bool(_TC_HALO and not split and c % 64 == 0 and c <= 128)

    # ---- sessions -----------------------------------------------------------------------
    def session(self, cond: torch.Tensor, guide: torch.Tensor, geometry: Optional[Geometry] = None) -> "Session":
        """A session binds one conditioning image batch (+ guidance) to a resident plan."""
        self.ensure_weights()
        B, _, h, w = cond.shape
        if geometry is None:
            geometry = self.default_geometry(B, h, w)
        key = (B, h, w, geometry.kind, geometry.TH, geometry.TW, geometry.PD, cond.shape[1], self.precision, self.shard_mode)
        s = self._sessions.get(key)
        if s is None:
            self._sessions.clear()                               # one resident plan at a time: plans hold GBs
            s = Session(self, geometry, cond.shape[1])
            self._sessions[key] = s
        s.bind(cond, guide)
        return s

    def default_geometry(self, B, h, w) -> Geometry:
        """model/ucdir.py:295-307 dispatch.  Memoised: building the tile tables costs ~2.5 ms of numpy, which a stateless
        p_sample() per step paid on every call (a fifth of a bf16 step; far more on a loaded host)."""
        m = self.m
        key = (B, h, w, m.tile_trigger, m.tile_skip, m.tile_padding)
        hit = self._geo_cache.get(key)
        if hit is None:
            if len(self._geo_cache) > 16:
                self._geo_cache.clear()
            hit = geometry_tiled(B, h, w, m.tile_skip, m.tile_padding) if h * w > m.tile_trigger else geometry_direct(B, h, w)
            self._geo_cache[key] = hit
        return hit

    # ---- reference-signature entry points ------------------------------------------------
    def forward(self, x, time, guide):
        """DY3h.forward(x[B,6,h,w], time[B,1], guide[B,3,h,w]) -> eps[B,3,h,w]  (model/ucdir.py:295-307)."""
        return self._forward_generic(x, time, guide, None)

    def forward_plain(self, x, time, guide):
        """DY3h.naiveforward: no padding (model/ucdir.py:270-293)."""
        self.ensure_weights()
        B, _, h, w = x.shape
        return self._forward_generic(x, time, guide, geometry_naive(B, h, w))

    def _forward_generic(self, x, time, guide, geometry):
        self.ensure_weights()
        x = x.contiguous().float()
        sess = self.session(x, guide.contiguous().float(), geometry=geometry)
        out = torch.empty((x.shape[0], self.m.cfg["out_channel"], x.shape[2], x.shape[3]), device=x.device, dtype=F32)
        sess.eps_only(time.reshape(-1).float().contiguous(), out)
        return out


class _PtrBuf:
    """Adapter: a raw device pointer where an Act expects a tensor."""

    def __init__(self, ptr): self._p = ptr
    def data_ptr(self): return self._p


# ======================================================================================
# session: resident per-image-batch state + the per-step op array
# ======================================================================================
class Session:
    """Everything that is invariant across the T denoising steps of one image batch: tile table, padded
    guidance tiles and the 27 guidance maps per chunk, workspace, and the op array of one step:

        TIME_EMBED -> per chunk [MEMSET stats, GATHER_TILES, UNet ops] -> (all-gather) -> SCATTER(+posterior)

    Tile sharding: with a process group given, rank r computes tiles [r*per, (r+1)*per) and one
    all_gather_into_tensor on the eps tile buffer reassembles the step; the posterior update runs
    redundantly on every rank with identical noise.
    """

    def __init__(self, eng: UNetEngine, geo: Geometry, in_channels: int):
        self.eng, self.geo = eng, geo
        dev = eng.device()
        self.dev = dev
        self.in_channels = in_channels          # 3: cond only (x_t supplied per step); 6: already concatenated
        self.group = None
        self.rank, self.world = 0, 1
        # Tile sharding is opt-in (UCDIR_SHARD=tiles or engine.shard_mode = "tiles"): the reference's own launcher line runs
        # a DIFFERENT image on every rank (data/__init__.py:30), which must keep working untouched under torchrun.
        if eng.shard_mode == "tiles" and torch.distributed.is_available() \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            self.group = torch.distributed.group.WORLD
            self.rank, self.world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        g = geo
        NT = g.n_tiles
        self.per_rank = (NT + self.world - 1) // self.world
        self.NT_pad = self.per_rank * self.world
        lo = min(self.rank * self.per_rank, NT)
        hi = min(lo + self.per_rank, NT)
        self.my_tiles = (lo, hi)
        tab = g.table()
        self.tab = torch.from_numpy(tab).to(dev)
        self.owner_y = torch.from_numpy(g.owner_y).to(dev)
        self.owner_x = torch.from_numpy(g.owner_x).to(dev)
        self.y0 = torch.tensor(g.ys, dtype=torch.int32, device=dev)
        self.x0 = torch.tensor(g.xs, dtype=torch.int32, device=dev)
        self.cond = torch.empty((g.B, in_channels, g.IH, g.IW), dtype=F32, device=dev)
        self.guide = torch.empty((g.B, 3, g.IH, g.IW), dtype=F32, device=dev)
        # only tile interiors are stitched (utils/util.py:144-145): the step's eps lives in a compact
        # [tile, TH-2c, TW-2c, 4] buffer, which is also what the sharded mode all-gathers
        self.crop = eng.m.tile_padding if g.kind == "tiled" else 0
        self.eps = torch.empty((self.NT_pad, g.TH - 2 * self.crop, g.TW - 2 * self.crop, 4), dtype=F32, device=dev)
        self.time_collective = False
        self.collective_events: List[tuple] = []
        per_chunk = max(1, _max_chunk_pixels() // (g.TH * g.TW))
        self.chunks = [(a, min(a + per_chunk, hi)) for a in range(lo, hi, per_chunk)]
        self.pool = Pool(dev)
        nblk = eng.n_blocks()
        self.levels = torch.zeros(max(NT, 1), dtype=F32, device=dev)      # per-tile noise levels (generic forward)
        self.attw = torch.empty((max(NT, 1), nblk, 8), dtype=F32, device=dev)
        self._mix_ops: List[tuple] = []          # (op, first tile of its chunk, attw base pointer, pointer key, stride key)
        self.static_ops = OpList()
        self.step_ops = OpList()
        self.idx_gather: List[int] = []
        self.gmaps_all = []
        self.guide_tiles_all = []
        maxbt = max((b - a) for a, b in self.chunks) if self.chunks else 1
        self.stats = torch.empty((MAX_STAT_SLOTS, maxbt, 2), dtype=torch.float64, device=dev)
        self.bf16 = eng.precision in ("bf16", "fp32_tc")          # tensor-core path (bf16 storage; fp32_tc: hi/lo plane pairs)
        self.split = eng.precision == "fp32_tc"
        self.x_tiles = torch.empty((maxbt, g.TH, g.TW, 32 if self.split else 16), dtype=BF16, device=dev) if self.bf16 else \
            torch.empty((maxbt, g.TH, g.TW, 8), dtype=F32, device=dev)
        # step op 0: timestep embedding -> attw table
        self.idx_temb = eng.time_embed_op(self.step_ops, self.attw, 1, 0, 0.0)
        for (a, b) in self.chunks:
            BT = b - a
            tab_ptr = self.tab.data_ptr() + a * 3 * 4
            gt = torch.empty((BT, g.TH, g.TW, 4), dtype=F32, device=dev)
            self.guide_tiles_all.append(gt)
            self.static_ops.add("UCDIR_OP_GATHER_TILES",
                                {"UCDIR_GATHER_P_SRC_A": self.guide.data_ptr(), "UCDIR_GATHER_P_TAB": tab_ptr,
                                 "UCDIR_GATHER_P_DST": gt.data_ptr()},
                                {"UCDIR_GATHER_I_BT": BT, "UCDIR_GATHER_I_TH": g.TH, "UCDIR_GATHER_I_TW": g.TW,
                                 "UCDIR_GATHER_I_IMG_H": g.IH, "UCDIR_GATHER_I_IMG_W": g.IW, "UCDIR_GATHER_I_PD": g.PD,
                                 "UCDIR_GATHER_I_CA": 3, "UCDIR_GATHER_I_CB": 0, "UCDIR_GATHER_I_CD": 4})
            gmaps = eng.build_guidance_ops(self.static_ops, self.pool, BT, g.TH, g.TW, gt)
            self.gmaps_all.append(gmaps)
            ca = in_channels
            self.idx_gather.append(self.step_ops.add(
                "UCDIR_OP_GATHER_TILES",
                {"UCDIR_GATHER_P_SRC_A": self.cond.data_ptr(), "UCDIR_GATHER_P_TAB": tab_ptr,
                 "UCDIR_GATHER_P_DST": self.x_tiles.data_ptr()},
                {"UCDIR_GATHER_I_BT": BT, "UCDIR_GATHER_I_TH": g.TH, "UCDIR_GATHER_I_TW": g.TW,
                 "UCDIR_GATHER_I_IMG_H": g.IH, "UCDIR_GATHER_I_IMG_W": g.IW, "UCDIR_GATHER_I_PD": g.PD,
                 "UCDIR_GATHER_I_CA": ca, "UCDIR_GATHER_I_CB": 6 - ca, "UCDIR_GATHER_I_CD": 16 if self.bf16 else 8,
                 "UCDIR_GATHER_I_OUT_BF16": (2 if self.split else 1) if self.bf16 else 0}))
            ih, iw = g.TH - 2 * self.crop, g.TW - 2 * self.crop
            fuse_crop = bool(self.crop) and self.bf16 and eng.fused_final_crop(self.split)
            if self.crop and not fuse_crop:
                if not hasattr(self, "eps_full"):
                    self.eps_full = torch.empty((maxbt, g.TH, g.TW, 4), dtype=F32, device=dev)
                eps_ptr = self.eps_full.data_ptr()
            else:
                eps_ptr = self.eps.data_ptr() + a * ih * iw * 4 * 4
            stats_view = self.stats.view(-1)[:MAX_STAT_SLOTS * BT * 2].view(MAX_STAT_SLOTS, BT, 2)   # same storage
            if self.bf16:
                sub = eng.build_forward_ops_bf16(self.pool, BT, g.TH, g.TW, self.x_tiles, gmaps, self.attw, 0, eps_ptr, stats_view,
                                                 split=self.split, eps_crop=self.crop if fuse_crop else 0)
            else:
                sub = eng.build_forward_ops(self.pool, BT, g.TH, g.TW, self.x_tiles, gmaps, self.attw, 0, eps_ptr, stats_view)
            self._chunk_attw_fix(sub, a)
            self.step_ops.extend(sub)
            if self.crop and not fuse_crop:
                self.step_ops.add("UCDIR_OP_CROP_TILES",
                                  {"UCDIR_CROP_P_SRC": self.eps_full.data_ptr(), "UCDIR_CROP_P_DST": self.eps.data_ptr() + a * ih * iw * 16},
                                  {"UCDIR_CROP_I_BT": BT, "UCDIR_CROP_I_TH": g.TH, "UCDIR_CROP_I_TW": g.TW, "UCDIR_CROP_I_IH": ih,
                                   "UCDIR_CROP_I_IW": iw, "UCDIR_CROP_I_OY": self.crop, "UCDIR_CROP_I_OX": self.crop})
        self.n_unet_ops = len(self.step_ops)
        # the scatter (+ posterior) closes the step; kept as a separate one-op list so that a collective can
        # be issued between the UNet ops and it
        self.tail_ops = OpList()
        self.tail_ops.add("UCDIR_OP_SCATTER",
                          {"UCDIR_SCATTER_P_EPS": self.eps.data_ptr(), "UCDIR_SCATTER_P_OWNER_Y": self.owner_y.data_ptr(),
                           "UCDIR_SCATTER_P_OWNER_X": self.owner_x.data_ptr(), "UCDIR_SCATTER_P_Y0": self.y0.data_ptr(),
                           "UCDIR_SCATTER_P_X0": self.x0.data_ptr()},
                          {"UCDIR_SCATTER_I_BIMG": g.B, "UCDIR_SCATTER_I_IMG_H": g.IH, "UCDIR_SCATTER_I_IMG_W": g.IW,
                           "UCDIR_SCATTER_I_NTY": g.nty, "UCDIR_SCATTER_I_NTX": g.ntx, "UCDIR_SCATTER_I_TH": g.TH - 2 * self.crop,
                           "UCDIR_SCATTER_I_TW": g.TW - 2 * self.crop, "UCDIR_SCATTER_I_PD": g.PD - self.crop, "UCDIR_SCATTER_I_CE": 4,
                           "UCDIR_SCATTER_I_MODE": 0, "UCDIR_SCATTER_I_CLIP": 1, "UCDIR_SCATTER_I_C": 3})
        self._attw_stride = 0
        self._bound = False
        self._guide_ref, self._guide_ver = None, -1
        self._cond_ref, self._cond_ver = None, -1
        self.n_binds = 0                         # guidance recomputations (tests assert on it)

    # per-sample attw addressing: sample b of chunk starting at tile a reads attw[(a + b) * stride]
    def _chunk_attw_fix(self, sub: OpList, a: int):
        for o in sub.ops:
            if o.kind == K_["UCDIR_OP_CONV_F32"] and o.i[K_["UCDIR_CONV_I_MODE"]] == 1:
                self._mix_ops.append((o, a, int(o.p[K_["UCDIR_CONV_P_ATTW"]]), "UCDIR_CONV_P_ATTW", "UCDIR_CONV_I_ATTW_STRIDE"))
            elif o.kind == K_["UCDIR_OP_TC_CONV"] and o.i[K_["UCDIR_TC_I_MODE"]] == 1:
                self._mix_ops.append((o, a, int(o.p[K_["UCDIR_TC_P_ATTW"]]), "UCDIR_TC_P_ATTW", "UCDIR_TC_I_ATTW_STRIDE"))

    def _set_attw_mode(self, per_tile: bool):
        """per_tile=False: one level for the whole batch (sampler).  True: levels[tile] (generic forward)."""
        nblk = self.eng.n_blocks()
        stride = nblk * 8 if per_tile else 0
        if stride == self._attw_stride and self._bound:
            return
        for o, a, base, pk, ik in self._mix_ops:
            o.i[K_[ik]] = stride
            o.p[K_[pk]] = base + a * stride * 4
        self._attw_stride = stride
        self.step_ops._arr = None

    # ---- binding -------------------------------------------------------------------------
    def stream(self) -> int:
        return _stream(self.dev)

    def bind(self, cond: torch.Tensor, guide: torch.Tensor):
        """Copy the conditioning batch in and (re)compute the step-invariant guidance maps.

        `cond` is copied on every call (12.6 MB D2D at 1024^2: microseconds next to a UNet step).  The guidance maps depend
        on `guide` only; they are recomputed unless the caller passes the very same tensor OBJECT again, unmodified.  The
        session keeps a strong reference to the bound guide, so "same object" cannot be a recycled allocation (round 1
        compared data_ptr / _version / shape, which a fresh tensor landing on a freed block of the caching allocator also
        satisfies -> the previous image's guidance was silently reused)."""
        self.cond.copy_(cond)
        same = self._bound and (guide is self._guide_ref) and guide._version == self._guide_ver
        if self.group is not None and not (same and cond is self._cond_ref and cond._version == self._cond_ver):
            self._check_ranks_agree(cond, guide)
            self._cond_ref, self._cond_ver = cond, cond._version
        if same:
            return
        self.guide.copy_(guide)
        self._guide_ref, self._guide_ver = guide, guide._version
        with _device_guard(self.dev), _Range("ucdir bind: guidance maps"):
            if len(self.static_ops):
                _run_ops(self.static_ops.array(), len(self.static_ops), self.stream())
        self._bound = True
        self.n_binds += 1

    def _check_ranks_agree(self, cond, guide):
        """Tile sharding computes ONE image with all ranks: every rank must hold the same conditioning batch.  (The
        reference's own multi-GPU mode gives every rank a different image, data/__init__.py:30; mixing the two would
        stitch tiles of different images.)  One small all-gather of two checksums per image."""
        mine = torch.stack([cond.double().sum(), guide.double().sum()]).to(self.dev)
        allv = torch.empty(self.world * 2, dtype=torch.float64, device=self.dev)
        torch.distributed.all_gather_into_tensor(allv, mine, group=self.group)
        allv = allv.view(self.world, 2)
        if not bool((allv == allv[0:1]).all().item()):
            raise RuntimeError("UCDIR_SHARD=tiles: ranks hold different conditioning images (checksums %s); tile sharding "
                               "needs the same image on every rank -- use the default (one image per rank) otherwise"
                               % allv.tolist())

    # ---- one UNet evaluation over all tiles, result stitched to NCHW ---------------------------
    def _run_unet(self, x_t: Optional[torch.Tensor]):
        ops = self.step_ops
        if x_t is not None:
            for k in self.idx_gather:
                ops.ops[k].p[K_["UCDIR_GATHER_P_SRC_B"]] = x_t.data_ptr()
            ops._arr = None
        _run_ops(ops.array(), len(ops), self.stream())
        if self.group is not None:
            self._all_gather()

    @_on_device
    def eps_only(self, levels: torch.Tensor, out: torch.Tensor):
        """Generic DY3h.forward: per-image noise levels from a device tensor, eps stitched into `out`."""
        g = self.geo
        self._set_attw_mode(True)
        self.levels[:g.n_tiles].copy_(levels.repeat_interleave(g.tiles_per_image))
        o = self.step_ops.ops[self.idx_temb]
        o.p[K_["UCDIR_TEMB_P_LEVELS"]] = self.levels.data_ptr()
        o.i[K_["UCDIR_TEMB_I_L"]] = g.n_tiles
        self.step_ops._arr = None
        self._run_unet(None)
        t = self.tail_ops.ops[0]
        t.i[K_["UCDIR_SCATTER_I_MODE"]] = 0
        t.p[K_["UCDIR_SCATTER_P_OUT"]] = out.data_ptr()
        self.tail_ops._arr = None
        _run_ops(self.tail_ops.array(), 1, self.stream())

    @_on_device
    def step(self, x_t: torch.Tensor, out: torch.Tensor, level: float, scalars, noise: Optional[torch.Tensor],
             clip: bool = True):
        """One p_sample (model/diffusion.py:160-183): eps = UNet(cat[cond, x_t], level); posterior update."""
        if self.in_channels != 3:
            raise RuntimeError("session was bound with a pre-concatenated input; step() needs cond only")
        self._set_attw_mode(False)
        o = self.step_ops.ops[self.idx_temb]
        o.p[K_["UCDIR_TEMB_P_LEVELS"]] = None
        o.i[K_["UCDIR_TEMB_I_L"]] = 1
        o.f[K_["UCDIR_TEMB_F_LEVEL"]] = level
        self._run_unet(x_t)
        a, b, c1, c2, sigma = scalars
        t = self.tail_ops.ops[0]
        t.i[K_["UCDIR_SCATTER_I_MODE"]] = 1
        t.i[K_["UCDIR_SCATTER_I_CLIP"]] = 1 if clip else 0
        t.p[K_["UCDIR_SCATTER_P_XT"]] = x_t.data_ptr()
        t.p[K_["UCDIR_SCATTER_P_NOISE"]] = noise.data_ptr() if noise is not None else None
        t.p[K_["UCDIR_SCATTER_P_OUT"]] = out.data_ptr()
        for k, v in zip(("A", "B", "C1", "C2", "SIGMA"), (a, b, c1, c2, sigma)):
            t.f[K_["UCDIR_SCATTER_F_" + k]] = v
        self.tail_ops._arr = None
        _run_ops(self.tail_ops.array(), 1, self.stream())

    # ---- resident stepping: state, noise and per-step scalars live on the device; one graph launch per step ------
    PARAM_FLOATS = 9          # {level, A, B, C1, C2, SIGMA, clip, use_noise, C3} followed by attw[n_blocks][8] of that level

    def param_floats(self) -> int:
        return self.PARAM_FLOATS + self.eng.n_blocks() * 8

    def ensure_resident(self):
        if getattr(self, "_resident", False):
            return
        if self.in_channels != 3:
            raise RuntimeError("resident stepping needs a session bound with cond only")
        g, dev = self.geo, self.dev
        self._set_attw_mode(False)
        self.xbuf = [torch.zeros((g.B, 3, g.IH, g.IW), dtype=F32, device=dev) for _ in range(2)]
        self.noise = torch.zeros((g.B, 3, g.IH, g.IW), dtype=F32, device=dev)
        self.params = torch.zeros(self.param_floats(), dtype=F32, device=dev)
        attw_base = self.params.data_ptr() + self.PARAM_FLOATS * 4
        mix_ids = {id(o): (pk, ik, int(o.p[K_[pk]]) - (self.attw.data_ptr() + a * self._attw_stride * 4))
                   for o, a, base, pk, ik in self._mix_ops}       # id(op) -> (pointer key, stride key, byte offset of its block row)
        self.cur = 0
        self.res_ops: List[Tuple[OpList, OpList]] = []
        clone = lambda o: _lib.Op.from_buffer_copy(o)
        for par in (0, 1):
            body, tail = OpList(), OpList()
            for idx, o in enumerate(self.step_ops.ops):
                if idx == self.idx_temb:
                    continue                     # attw of the step's level arrives with the parameter row (engine.attw_rows)
                c = clone(o)
                if id(o) in mix_ids:
                    pk, ik, off = mix_ids[id(o)]
                    c.p[K_[pk]] = attw_base + off
                    c.i[K_[ik]] = 0
                if idx in self.idx_gather:
                    c.p[K_["UCDIR_GATHER_P_SRC_B"]] = self.xbuf[par].data_ptr()
                body.ops.append(c)
            t = clone(self.tail_ops.ops[0])
            t.i[K_["UCDIR_SCATTER_I_MODE"]] = 1
            t.p[K_["UCDIR_SCATTER_P_XT"]] = self.xbuf[par].data_ptr()
            t.p[K_["UCDIR_SCATTER_P_NOISE"]] = self.noise.data_ptr()
            t.p[K_["UCDIR_SCATTER_P_OUT"]] = self.xbuf[1 - par].data_ptr()
            t.p[K_["UCDIR_SCATTER_P_PARAMS"]] = self.params.data_ptr() + 4
            tail.ops.append(t)
            self.res_ops.append((body, tail))
        self.graphs = None
        self.use_graphs = (dev.type == "cuda") and not _TEST_CPU_PLAN and os.environ.get("UCDIR_CUDA_GRAPH", "1") != "0"
        self._resident = True

    def _capture_graphs(self):
        """Run each parity once eagerly (one-time function attributes, lazy module loading), then capture."""
        graphs = []
        for body, tail in self.res_ops:
            if len(body):
                _run_ops(body.array(), len(body), self.stream())
            _run_ops(tail.array(), 1, self.stream())
        torch.cuda.synchronize(self.dev)
        for body, tail in self.res_ops:
            if self.group is None:
                both = OpList(); both.ops = body.ops + tail.ops
                graphs.append((_lib.Graph(both.array(), len(both)), None))
            else:                                # a rank can own no tiles at all (more ranks than tiles): empty body
                graphs.append((_lib.Graph(body.array(), len(body)) if len(body) else None, _lib.Graph(tail.array(), 1)))
        self.graphs = graphs

    @_on_device
    def load_state(self, x: torch.Tensor):
        self.ensure_resident()
        if self.use_graphs and self.graphs is None:
            self._capture_graphs()
        self.cur = 0
        self.xbuf[0].copy_(x)

    def state(self) -> torch.Tensor:
        return self.xbuf[self.cur]

    @_on_device
    def step_resident(self, params_row: torch.Tensor):
        """One p_sample on the resident state.  params_row: device float[param_floats()] for this step -- the nine scalars and
        the attw rows of the step's level -- D2D copied into the slot the ops read; the caller has already filled self.noise
        when the step uses noise."""
        self.params.copy_(params_row)
        body, tail = self.res_ops[self.cur]
        st = self.stream()
        if _NVTX:
            torch.cuda.nvtx.range_push("ucdir p_sample step")
        if self.graphs is not None:
            gb, gt = self.graphs[self.cur]
            if gb is not None:
                gb.launch(st)
            if self.group is not None:
                self._all_gather()
                gt.launch(st)
        else:
            if len(body):
                _run_ops(body.array(), len(body), st)
            if self.group is not None:
                self._all_gather()
            _run_ops(tail.array(), 1, st)
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        self.cur ^= 1

    def _all_gather(self):
        if self.time_collective:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        with _Range("ucdir tile all-gather"):
            torch.distributed.all_gather_into_tensor(self.eps, self.eps[self.rank * self.per_rank:(self.rank + 1) * self.per_rank],
                                                     group=self.group)
        if self.time_collective:
            e1.record()
            self.collective_events.append((e0, e1))

    def launches_per_step(self) -> int:
        n = 0
        for o in self.step_ops.ops + self.tail_ops.ops:
            n += 0 if o.kind == K_["UCDIR_OP_MEMSET"] else 1
        return n


# ======================================================================================
# predictor engine (UNetSeeInDark, model/ucdir.py:310-416)
# ======================================================================================
def _pad_c(c: int) -> int:
    """channel count of a predictor tensor on the tensor-core path: multiples of 64 (32-channel levels carry 32 zero channels)."""
    return _round_up(max(c, 64), 64)


class PredictorEngine:
    """UNetSeeInDark (model/ucdir.py:310-416), once per image.  Two plans: "fp32" = SIMT FFMA kernels; "tc" = the tensor-core path
    with split (hi + lo bf16) operands -- fp32-class accuracy (the predictor's output is ADDED to the sampler's result,
    model/diffusion.py:478, so it is never run with plain bf16 operands), LeakyReLU in the conv epilogue, ConvTranspose2d(2, 2) as
    four 1x1 phase GEMMs, max-pool on plane pairs.  The sampler picks "tc" whenever the denoiser runs on the tensor cores."""

    def __init__(self, module):
        self.m = module
        self.ws: Optional[WeightStore] = None
        self._plans: Dict[tuple, tuple] = {}
        self.mode = "fp32" if os.environ.get("UCDIR_PRECISION", "fp32_tc") == "fp32" else "tc"

    def set_mode(self, mode: str):
        if mode not in ("fp32", "tc"):
            raise ValueError("predictor mode must be fp32 or tc")
        if mode != self.mode:
            self.mode = mode
            self.invalidate_weights()

    def invalidate_weights(self):
        self.ws = None
        self._params = None
        self._plans.clear()

    def device(self):
        return next(self.m.parameters()).device

    def ensure_weights(self):
        dev = self.device()
        _require_cuda(dev)
        if getattr(self, "_params", None) is None:
            self._params = list(self.m.parameters())
        ver = sum(p._version for p in self._params)
        if self.ws is not None and self.ws.device == dev and ver == getattr(self, "_packed_version", -1):
            return
        self._packed_version = ver
        self._plans.clear()
        ws = WeightStore(dev)
        if self.mode == "tc":
            self._pack_tc(ws)
            self.ws = ws
            return
        for name, layer in self.m.named_children():
            if isinstance(layer, torch.nn.ConvTranspose2d):
                for py in range(2):
                    for px in range(2):
                        ws.put("%s.w%d%d" % (name, py, px), pack_convT_phase_f32(layer.weight, py, px))
                ws.put(name + ".b", layer.bias.float())
            elif isinstance(layer, torch.nn.Conv2d):
                ws.put(name + ".w", pack_conv_f32(layer.weight, pad_cin_to=8 if layer.in_channels < 8 else None))
                ws.put(name + ".b", layer.bias.float())
        self.ws = ws

    # ---- tensor-core plan (split operands) ------------------------------------------------------------------------------
    def _pack_tc(self, ws: WeightStore):
        from .model.ucdir import predictor_plan
        cin_layout = {}                               # conv name -> list of (real channels, padded channels) per source
        c_prev, enc = (3, 16), []
        for name, kind, cin, cout in predictor_plan():
            if kind == "conv3" or kind == "conv1":
                if name.endswith("_1") and int(name[4:].split("_")[0]) in (6, 7, 8, 9):      # torch.cat([up, skip], 1)
                    srcs = [(cout, _pad_c(cout)), (cout, _pad_c(cout))]
                else:
                    srcs = [c_prev]
                cin_layout[name] = srcs
                c_prev = (cout, _pad_c(cout))
            elif kind == "convT":
                cin_layout[name] = [c_prev]
                c_prev = (cout, _pad_c(cout))
        for name, kind, cin, cout in predictor_plan():
            layer = getattr(self.m, name)
            if kind == "pool":
                continue
            srcs = cin_layout[name]
            cip = sum(pp for _, pp in srcs)
            if kind == "convT":                       # [Cin][Cout][2][2] -> four 1x1 GEMMs [Cout][Cin]
                cop = _pad_c(cout)
                for py in range(2):
                    for px in range(2):
                        wp = torch.zeros(cop, cip, 1, 1, device=layer.weight.device)
                        wp[:cout, :cin, 0, 0] = layer.weight[:, :, py, px].t().float()
                        bp = torch.zeros(cop, device=layer.weight.device); bp[:cout] = layer.bias.float()
                        w_, tb_, _ = pack_tc_dense(wp, bp, _tc_nt(cop), split=True)
                        ws.put("%s.p%d%d.tcw" % (name, py, px), w_); ws.put("%s.p%d%d.tb" % (name, py, px), tb_)
                continue
            last = kind == "conv1"
            cop = 16 if last else _pad_c(cout)
            ks = layer.kernel_size[0]
            wp = torch.zeros(cop, cip, ks, ks, device=layer.weight.device)
            off_real = off_pad = 0
            for real, padded in srcs:
                wp[:cout, off_pad:off_pad + real] = layer.weight[:, off_real:off_real + real].float()
                off_real += real; off_pad += padded
            bp = torch.zeros(cop, device=layer.weight.device); bp[:cout] = layer.bias.float()
            w_, tb_, _ = pack_tc_dense(wp, bp, 16 if last else _tc_nt(cop), split=True, c0=srcs[0][1])
            ws.put(name + ".tcw", w_); ws.put(name + ".tb", tb_)

    def _plan_tc(self, B, h, w):
        from .model.ucdir import PREDICTOR_WIDTHS
        dev = self.device()
        ws = self.ws
        geo = geometry_direct(B, h, w)                            # model/ucdir.py:352-358: same pad rule
        TH, TW = geo.TH, geo.TW
        pool = Pool(dev)
        x_img = torch.empty((B, 3, h, w), dtype=F32, device=dev)
        x_tiles = torch.empty((B, TH, TW, 32), dtype=BF16, device=dev)         # (hi, lo) planes of 16 channels (3 valid)
        tab = torch.from_numpy(geo.table()).to(dev)
        out_tiles = torch.empty((B, TH, TW, 4), dtype=F32, device=dev)
        dummy_stats = torch.zeros((1, B, 2), dtype=torch.float64, device=dev)
        bld = _Builder(pool, B, dummy_stats, elem=2, split=True)
        ol = bld.ops
        ol.add("UCDIR_OP_GATHER_TILES",
               {"UCDIR_GATHER_P_SRC_A": x_img.data_ptr(), "UCDIR_GATHER_P_TAB": tab.data_ptr(), "UCDIR_GATHER_P_DST": x_tiles.data_ptr()},
               {"UCDIR_GATHER_I_BT": B, "UCDIR_GATHER_I_TH": TH, "UCDIR_GATHER_I_TW": TW, "UCDIR_GATHER_I_IMG_H": h,
                "UCDIR_GATHER_I_IMG_W": w, "UCDIR_GATHER_I_PD": 0, "UCDIR_GATHER_I_CA": 3, "UCDIR_GATHER_I_CB": 0,
                "UCDIR_GATHER_I_CD": 16, "UCDIR_GATHER_I_OUT_BF16": 2})

        def conv(name, x: Act, cout, skip: Optional[Act] = None, kc=64) -> Act:
            cop = _pad_c(cout)
            y = bld.new(cop, x.H, x.W, with_stats=False)
            _tc_op(ol, split=1, src0=x, src1=skip, w=ws.ptr(name + ".tcw"), tb=ws.ptr(name + ".tb"), kc=kc, act=2, dst=y, ntot=cop, B=B,
                   nt=_tc_nt(cop))
            return y

        def pool2(x: Act) -> Act:
            y = bld.new(x.C, x.H // 2, x.W // 2, with_stats=False)
            ol.add("UCDIR_OP_MAXPOOL2", {"UCDIR_POOL_P_SRC": x.ptr, "UCDIR_POOL_P_DST": y.ptr},
                   {"UCDIR_POOL_I_B": B, "UCDIR_POOL_I_H": y.H, "UCDIR_POOL_I_W": y.W, "UCDIR_POOL_I_C": x.C, "UCDIR_POOL_I_SPLIT": 1})
            return y

        def upconv(name, x: Act, cout) -> Act:
            cop = _pad_c(cout)
            y = bld.new(cop, x.H * 2, x.W * 2, with_stats=False)
            for py in range(2):
                for px in range(2):
                    _tc_op(ol, split=1, src0=x, w=ws.ptr("%s.p%d%d.tcw" % (name, py, px)), tb=ws.ptr("%s.p%d%d.tb" % (name, py, px)), nty=1,
                           ntx=1, oy0=0, ox0=0, dst=y, ntot=cop, B=B, nt=_tc_nt(cop), dst_up=1, dst_py=py, dst_px=px)
            return y

        x = Act(x_tiles, 16, TH, TW, 0, keep=True, split=True)
        enc = []
        for lvl, c in enumerate(PREDICTOR_WIDTHS, start=1):
            a = conv("conv%d_1" % lvl, x, c, kc=16 if lvl == 1 else 64)
            if lvl > 1:
                bld.release(x)
            x = conv("conv%d_2" % lvl, a, c)
            bld.release(a)
            if lvl < len(PREDICTOR_WIDTHS):
                x.keep = True
                enc.append(x)
                x = pool2(x)
        for lvl, c in zip(range(6, 10), reversed(PREDICTOR_WIDTHS[:-1])):
            u = upconv("upv%d" % lvl, x, c)
            bld.release(x)
            skip = enc.pop()
            skip.keep = False
            a = conv("conv%d_1" % lvl, u, c, skip=skip)           # torch.cat([up, conv_k], 1) as a dual-source K loop
            bld.release(u); bld.release(skip)
            x = conv("conv%d_2" % lvl, a, c)
            bld.release(a)
        dst = Act(out_tiles, 4, TH, TW, 0, keep=True)
        _tc_op(ol, split=1, src0=x, w=ws.ptr("conv10_1.tcw"), tb=ws.ptr("conv10_1.tb"), nty=1, ntx=1, oy0=0, ox0=0, dst=dst, ntot=16, B=B,
               nt=16, dst_f32=1, ncol_valid=3)
        bld.release(x)
        owner_y = torch.from_numpy(geo.owner_y).to(dev); owner_x = torch.from_numpy(geo.owner_x).to(dev)
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        idx_scatter = ol.add("UCDIR_OP_SCATTER",
                             {"UCDIR_SCATTER_P_EPS": out_tiles.data_ptr(), "UCDIR_SCATTER_P_OWNER_Y": owner_y.data_ptr(),
                              "UCDIR_SCATTER_P_OWNER_X": owner_x.data_ptr(), "UCDIR_SCATTER_P_Y0": z.data_ptr(),
                              "UCDIR_SCATTER_P_X0": z.data_ptr()},
                             {"UCDIR_SCATTER_I_BIMG": B, "UCDIR_SCATTER_I_IMG_H": h, "UCDIR_SCATTER_I_IMG_W": w,
                              "UCDIR_SCATTER_I_NTY": 1, "UCDIR_SCATTER_I_NTX": 1, "UCDIR_SCATTER_I_TH": TH,
                              "UCDIR_SCATTER_I_TW": TW, "UCDIR_SCATTER_I_PD": 0, "UCDIR_SCATTER_I_CE": 4,
                              "UCDIR_SCATTER_I_MODE": 0, "UCDIR_SCATTER_I_C": 3})
        return (ol, x_img, idx_scatter, (pool, x_tiles, tab, out_tiles, owner_y, owner_x, z, dummy_stats))

    def _plan(self, B, h, w):
        key = (B, h, w, self.mode)
        if key in self._plans:
            return self._plans[key]
        self._plans.clear()
        if self.mode == "tc":
            plan = self._plan_tc(B, h, w)
            self._plans[key] = plan
            return plan
        dev = self.device()
        ws = self.ws
        geo = geometry_direct(B, h, w)                            # model/ucdir.py:352-358: same pad rule
        TH, TW = geo.TH, geo.TW
        pool = Pool(dev)
        x_img = torch.empty((B, 3, h, w), dtype=F32, device=dev)
        x_tiles = torch.empty((B, TH, TW, 8), dtype=F32, device=dev)
        tab = torch.from_numpy(geo.table()).to(dev)
        out_tiles = torch.empty((B, TH, TW, 4), dtype=F32, device=dev)
        dummy_stats = torch.zeros((1, B, 2), dtype=torch.float64, device=dev)
        bld = _Builder(pool, B, dummy_stats)
        ol = bld.ops
        ol.add("UCDIR_OP_GATHER_TILES",
               {"UCDIR_GATHER_P_SRC_A": x_img.data_ptr(), "UCDIR_GATHER_P_TAB": tab.data_ptr(),
                "UCDIR_GATHER_P_DST": x_tiles.data_ptr()},
               {"UCDIR_GATHER_I_BT": B, "UCDIR_GATHER_I_TH": TH, "UCDIR_GATHER_I_TW": TW, "UCDIR_GATHER_I_IMG_H": h,
                "UCDIR_GATHER_I_IMG_W": w, "UCDIR_GATHER_I_PD": 0, "UCDIR_GATHER_I_CA": 3, "UCDIR_GATHER_I_CB": 0,
                "UCDIR_GATHER_I_CD": 8})

        def conv(name, x: Act, cout, skip: Optional[Act] = None, ks=3, act=2, dst: Optional[Act] = None) -> Act:
            y = dst or bld.new(cout, x.H, x.W, with_stats=False)
            _conv_op(ol, src0=x, src1=skip, w=ws.ptr(name + ".w"), bias=ws.ptr(name + ".b"), ks=ks, act=act, dst=y,
                     cout=cout, B=B)
            return y

        def pool2(x: Act) -> Act:
            y = bld.new(x.C, x.H // 2, x.W // 2, with_stats=False)
            ol.add("UCDIR_OP_MAXPOOL2", {"UCDIR_POOL_P_SRC": x.ptr, "UCDIR_POOL_P_DST": y.ptr},
                   {"UCDIR_POOL_I_B": B, "UCDIR_POOL_I_H": y.H, "UCDIR_POOL_I_W": y.W, "UCDIR_POOL_I_C": x.C})
            return y

        def upconv(name, x: Act, cout) -> Act:
            """ConvTranspose2d(k=2, s=2) as four 1x1 GEMMs writing interleaved output phases (ucdir.py:381-400)."""
            y = bld.new(cout, x.H * 2, x.W * 2, with_stats=False)
            for py in range(2):
                for px in range(2):
                    _conv_op(ol, src0=x, w=ws.ptr("%s.w%d%d" % (name, py, px)), bias=ws.ptr(name + ".b"), ks=1, dst=y,
                             cout=cout, B=B, dst_up=1, dst_py=py, dst_px=px)
            return y

        x = Act(x_tiles, 8, TH, TW, 0, keep=True)
        enc = []
        chans = [32, 64, 128, 256, 512]
        for lvl, c in enumerate(chans, start=1):
            a = conv("conv%d_1" % lvl, x, c)
            if lvl > 1:
                bld.release(x)
            x = conv("conv%d_2" % lvl, a, c)
            bld.release(a)
            if lvl < 5:
                x.keep = True
                enc.append(x)
                x = pool2(x)
        for lvl, c in zip(range(6, 10), [256, 128, 64, 32]):
            u = upconv("upv%d" % lvl, x, c)
            bld.release(x)
            skip = enc.pop()
            skip.keep = False
            a = conv("conv%d_1" % lvl, u, c, skip=skip)           # torch.cat([up, conv_k], 1) as a dual-source K loop
            bld.release(u); bld.release(skip)
            x = conv("conv%d_2" % lvl, a, c)
            bld.release(a)
        dst = Act(out_tiles, 4, TH, TW, 0, keep=True)
        conv("conv10_1", x, 3, ks=1, act=0, dst=dst)
        bld.release(x)
        owner_y = torch.from_numpy(geo.owner_y).to(dev); owner_x = torch.from_numpy(geo.owner_x).to(dev)
        z = torch.zeros(1, dtype=torch.int32, device=dev)
        idx_scatter = ol.add("UCDIR_OP_SCATTER",
                             {"UCDIR_SCATTER_P_EPS": out_tiles.data_ptr(), "UCDIR_SCATTER_P_OWNER_Y": owner_y.data_ptr(),
                              "UCDIR_SCATTER_P_OWNER_X": owner_x.data_ptr(), "UCDIR_SCATTER_P_Y0": z.data_ptr(),
                              "UCDIR_SCATTER_P_X0": z.data_ptr()},
                             {"UCDIR_SCATTER_I_BIMG": B, "UCDIR_SCATTER_I_IMG_H": h, "UCDIR_SCATTER_I_IMG_W": w,
                              "UCDIR_SCATTER_I_NTY": 1, "UCDIR_SCATTER_I_NTX": 1, "UCDIR_SCATTER_I_TH": TH,
                              "UCDIR_SCATTER_I_TW": TW, "UCDIR_SCATTER_I_PD": 0, "UCDIR_SCATTER_I_CE": 4,
                              "UCDIR_SCATTER_I_MODE": 0, "UCDIR_SCATTER_I_C": 3})
        plan = (ol, x_img, idx_scatter, (pool, x_tiles, tab, out_tiles, owner_y, owner_x, z, dummy_stats))
        self._plans[key] = plan
        return plan

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.ensure_weights()
        B, c, h, w = x.shape
        if c != 3:
            raise ValueError("predictor expects 3 input channels")
        ol, x_img, idx_scatter, _ = self._plan(B, h, w)
        x_img.copy_(x)
        out = torch.empty((B, 3, h, w), dtype=F32, device=x.device)
        ol.ops[idx_scatter].p[K_["UCDIR_SCATTER_P_OUT"]] = out.data_ptr()
        ol._arr = None
        with _device_guard(x.device):
            _run_ops(ol.array(), len(ol), _stream(x.device))
        return out


# ======================================================================================
# FiLM ResnetBlock (model/ucdir.py:86-100) as a module-level op on the fp32 kernels
# ======================================================================================
def run_film_block(mod, x: torch.Tensor, time_emb: torch.Tensor) -> torch.Tensor:
    """h = conv(Swish(GN_g(x))); h = FiLM(h, Linear(t)); h = conv(Swish(GN_g(h))); return h + res_conv(x).
    x: [B, C, H, W] fp32 NCHW, time_emb: [B, nl_emb_dim].  Channel counts must be multiples of 8."""
    dev = x.device
    _require_cuda(dev)
    B, C, H, W = x.shape
    Co, G = mod.dim_out, mod.norm_groups
    if C != mod.dim or C % 8 or Co % 8 or time_emb.shape[-1] % 8:
        raise ValueError("FiLM block: channel counts must match the module and be multiples of 8")
    f32 = lambda t: t.detach().to(device=dev, dtype=F32).contiguous()
    keep = []                                             # keep every buffer alive until the ops have been issued

    def buf(*shape, dtype=F32):
        t = torch.empty(shape, dtype=dtype, device=dev); keep.append(t); return t

    xs = f32(x); te = f32(time_emb.reshape(B, -1)); keep += [xs, te]
    E = te.shape[1]
    aff = mod.noise_func.use_affine_level
    lin = mod.noise_func.noise_func[0]
    w_lin = pack_conv_f32(lin.weight.view(lin.out_features, E, 1, 1)).to(dev); b_lin = f32(lin.bias)
    c1, c2 = mod.block1.block[3], mod.block2.block[3]
    w1, w2 = pack_conv_f32(c1.weight).to(dev), pack_conv_f32(c2.weight).to(dev)
    b1, b2 = f32(c1.bias), f32(c2.bias)
    g1, be1 = f32(mod.block1.block[0].weight), f32(mod.block1.block[0].bias)
    g2, be2 = f32(mod.block2.block[0].weight), f32(mod.block2.block[0].bias)
    keep += [w_lin, b_lin, w1, w2, b1, b2, g1, be1, g2, be2]
    x_nhwc, a1, h1, a2, h2 = buf(B, H, W, C), buf(B, H, W, C), buf(B, H, W, Co), buf(B, H, W, Co), buf(B, H, W, Co)
    film = buf(B, 1, 1, lin.out_features)
    st1, st2 = buf(B, G, 2, dtype=torch.float64), buf(B, G, 2, dtype=torch.float64)
    out = torch.empty((B, Co, H, W), dtype=F32, device=dev)
    ol = OpList()
    A = lambda t, c, h=H, w=W: Act(t, c, h, w, 0, True)
    ol.add("UCDIR_OP_LAYOUT", {0: xs.data_ptr(), 1: x_nhwc.data_ptr()}, {0: B, 1: C, 2: H * W, 3: 0})

    def gn(src, dst, gamma, beta, stats, ch):
        ol.add("UCDIR_OP_GN_STATS_F32", {"UCDIR_GNS_P_SRC": src.data_ptr(), "UCDIR_GNS_P_STATS": stats.data_ptr()},
               {"UCDIR_GNS_I_B": B, "UCDIR_GNS_I_HW": H * W, "UCDIR_GNS_I_C": ch, "UCDIR_GNS_I_G": G})
        ol.add("UCDIR_OP_GN_APPLY_F32",
               {"UCDIR_GNF_P_SRC": src.data_ptr(), "UCDIR_GNF_P_DST": dst.data_ptr(), "UCDIR_GNF_P_GAMMA": gamma.data_ptr(),
                "UCDIR_GNF_P_BETA": beta.data_ptr(), "UCDIR_GNF_P_STATS": stats.data_ptr()},
               {"UCDIR_GNS_I_B": B, "UCDIR_GNS_I_HW": H * W, "UCDIR_GNS_I_C": ch, "UCDIR_GNS_I_G": G, "UCDIR_GNS_I_SWISH": 1}, {0: 1e-5})

    gn(x_nhwc, a1, g1, be1, st1, C)
    _conv_op(ol, src0=A(te, E, 1, 1), w=w_lin.data_ptr(), bias=b_lin.data_ptr(), ks=1, dst=A(film, lin.out_features, 1, 1),
             cout=lin.out_features, B=B)                                     # FeatureWiseAffine's Linear as a 1x1 GEMM
    _conv_op(ol, src0=A(a1, C), w=w1.data_ptr(), bias=b1.data_ptr(), dst=A(h1, Co), cout=Co, B=B)
    o = ol.ops[-1]
    if aff:
        o.p[K_["UCDIR_CONV_P_FILM_G"]] = film.data_ptr()
        o.p[K_["UCDIR_CONV_P_FILM_B"]] = film.data_ptr() + Co * 4
    else:
        o.p[K_["UCDIR_CONV_P_FILM_B"]] = film.data_ptr()
    o.i[K_["UCDIR_CONV_I_FILM_STRIDE"]] = lin.out_features
    gn(h1, a2, g2, be2, st2, Co)
    if isinstance(mod.res_conv, torch.nn.Conv2d):
        wr, br = pack_conv_f32(mod.res_conv.weight).to(dev), f32(mod.res_conv.bias); keep += [wr, br]
        res = buf(B, H, W, Co)
        _conv_op(ol, src0=A(x_nhwc, C), w=wr.data_ptr(), bias=br.data_ptr(), ks=1, dst=A(res, Co), cout=Co, B=B)
    else:
        res = x_nhwc
    _conv_op(ol, src0=A(a2, Co), w=w2.data_ptr(), bias=b2.data_ptr(), res=A(res, Co), dst=A(h2, Co), cout=Co, B=B)
    ol.add("UCDIR_OP_LAYOUT", {0: h2.data_ptr(), 1: out.data_ptr()}, {0: B, 1: Co, 2: H * W, 3: 1})
    with _device_guard(dev):
        _run_ops(ol.array(), len(ol), _stream(dev))
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()      # buffers in `keep` are released when this returns
    return out
