"""CPU interpreter of `ucdir_op_t` arrays -- TEST INFRASTRUCTURE ONLY.

The build container has no GPU, so the engine's graph construction (ucdir_b200/engine.py: weight packing,
tile tables, buffer reuse, op wiring) is checked here by executing the op array it emits with this
interpreter (torch CPU, semantics restated from include/ucdir_b200.h) and comparing against the oracle.
The CUDA kernels themselves are checked on the B200 by the `-m gpu` tests.  Nothing under ucdir_b200/
imports this file; the product path has no CPU route.
"""
from __future__ import annotations

import ctypes
import gc
import math

import numpy as np
import torch
import torch.nn.functional as F

from ucdir_b200 import _lib
from ucdir_b200._lib import C as K


class Memory:
    """Resolve raw host pointers to views of live torch CPU storages."""

    def __init__(self):
        self.iv = []
        seen = set()
        import warnings
        warnings.simplefilter("ignore")
        for o in gc.get_objects():
            try:
                if isinstance(o, torch.Tensor) and o.device.type == "cpu" and o.numel() > 0:
                    st = o.untyped_storage()
                    p = st.data_ptr()
                    if p and p not in seen:
                        seen.add(p)
                        self.iv.append((p, st.nbytes(), st))
            except Exception:
                pass

    def view(self, ptr, shape, dtype=torch.float32):
        if not ptr:
            return None
        n = int(np.prod(shape))
        es = torch.empty(0, dtype=dtype).element_size()
        for p, nb, st in self.iv:
            if p <= ptr and ptr + n * es <= p + nb:
                off = ptr - p
                assert off % es == 0
                return torch.empty(0, dtype=dtype).set_(st, off // es, tuple(shape))
        raise KeyError("pointer %#x (+%d bytes) is not inside any live CPU tensor" % (ptr, n * es))


def _p(op, name):
    v = op.p[K[name]]
    return int(v) if v else 0


def _i(op, name):
    return int(op.i[K[name]])


def _f(op, name):
    return float(op.f[K[name]])


def swish(x):
    return x * torch.sigmoid(x)


def emu_conv(op, mem):
    B, H, W = _i(op, "UCDIR_CONV_I_B"), _i(op, "UCDIR_CONV_I_H"), _i(op, "UCDIR_CONV_I_W")
    C0, C1, Cout = _i(op, "UCDIR_CONV_I_C0"), _i(op, "UCDIR_CONV_I_C1"), _i(op, "UCDIR_CONV_I_COUT")
    ks, stride, up, groups = (_i(op, "UCDIR_CONV_I_KSIZE"), _i(op, "UCDIR_CONV_I_STRIDE"), _i(op, "UCDIR_CONV_I_UP"),
                              _i(op, "UCDIR_CONV_I_GROUPS"))
    pre, act, mode = _i(op, "UCDIR_CONV_I_PRE"), _i(op, "UCDIR_CONV_I_ACT"), _i(op, "UCDIR_CONV_I_MODE")
    sH, sW = _i(op, "UCDIR_CONV_I_SRC_H"), _i(op, "UCDIR_CONV_I_SRC_W")
    dstC, dstCoff = _i(op, "UCDIR_CONV_I_DST_C"), _i(op, "UCDIR_CONV_I_DST_COFF")
    dstUp, dpy, dpx = _i(op, "UCDIR_CONV_I_DST_UP"), _i(op, "UCDIR_CONV_I_DST_PY"), _i(op, "UCDIR_CONV_I_DST_PX")
    resC, attw_stride = _i(op, "UCDIR_CONV_I_RES_C"), _i(op, "UCDIR_CONV_I_ATTW_STRIDE")
    eps = _f(op, "UCDIR_CONV_F_EPS")
    Cin = C0 + C1
    x = mem.view(_p(op, "UCDIR_CONV_P_SRC0"), (B, sH, sW, C0))
    if C1:
        x = torch.cat([x, mem.view(_p(op, "UCDIR_CONV_P_SRC1"), (B, sH, sW, C1))], dim=-1)
    x = x.clone()
    if pre:
        s0 = mem.view(_p(op, "UCDIR_CONV_P_STATS0"), (B, 2), torch.float64).clone()
        if C1:
            s0 = s0 + mem.view(_p(op, "UCDIR_CONV_P_STATS1"), (B, 2), torch.float64)
        cnt = float(Cin * sH * sW)
        mean = s0[:, 0] / cnt
        var = (s0[:, 1] / cnt - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        gamma = mem.view(_p(op, "UCDIR_CONV_P_GAMMA"), (Cin,))
        beta = mem.view(_p(op, "UCDIR_CONV_P_BETA"), (Cin,))
        x = (x - mean.float().view(B, 1, 1, 1)) * rstd.float().view(B, 1, 1, 1) * gamma + beta
        if pre == 2:
            x = swish(x)
    xn = x.permute(0, 3, 1, 2)
    if up:
        xn = F.interpolate(xn, scale_factor=2, mode="nearest")
    Cg, Ng = Cin // groups, Cout // groups
    ldw = (Ng + 3) & ~3
    Kdim = ks * ks * Cg
    wp = mem.view(_p(op, "UCDIR_CONV_P_W"), (groups, Kdim, ldw))[:, :, :Ng]
    w = wp.view(groups, ks, ks, Cg, Ng).permute(0, 4, 3, 1, 2).reshape(Cout, Cg, ks, ks)
    y = F.conv2d(xn, w, None, stride=stride, padding=ks // 2, groups=groups)          # B, Cout, H, W
    assert y.shape[2] == H and y.shape[3] == W, (y.shape, H, W)
    bias_p = _p(op, "UCDIR_CONV_P_BIAS")
    if bias_p:
        y = y + mem.view(bias_p, (Cout,)).view(1, Cout, 1, 1)
    y = y.permute(0, 2, 3, 1)                                                           # B, H, W, Cout
    if mode == 1:
        att = mem.view(_p(op, "UCDIR_CONV_P_ATT"), (B, H, W, 8))
        aw_base = _p(op, "UCDIR_CONV_P_ATTW")
        aw = torch.stack([mem.view(aw_base + b * attw_stride * 4, (8,)) for b in range(B)])
        a = att * aw.view(B, 1, 1, 8)
        c = Cout // 8
        h = (y.reshape(B, H, W, c, 8) * a.unsqueeze(3)).sum(-1)
        res = mem.view(_p(op, "UCDIR_CONV_P_RES"), (B, H, W, resC))[..., :c]
        v = swish(h) + res
        nout = c
    else:
        fb = _p(op, "UCDIR_CONV_P_FILM_B")
        if fb:
            fs = _i(op, "UCDIR_CONV_I_FILM_STRIDE") or Cout
            row = lambda ptr: torch.stack([mem.view(ptr + b * fs * 4, (Cout,)) for b in range(B)]).view(B, 1, 1, Cout)
            b_ = row(fb)
            fg = _p(op, "UCDIR_CONV_P_FILM_G")
            y = (1 + row(fg)) * y + b_ if fg else y + b_
        if act == 1:
            y = swish(y)
        elif act == 2:
            y = torch.max(0.2 * y, y)
        rp = _p(op, "UCDIR_CONV_P_RES")
        if rp:
            y = y + mem.view(rp, (B, H, W, resC))[..., :Cout]
        v = y
        nout = Cout
    if dstUp:
        dst = mem.view(_p(op, "UCDIR_CONV_P_DST"), (B, 2 * H, 2 * W, dstC))
        dst[:, dpy::2, dpx::2, dstCoff:dstCoff + nout] = v
    else:
        dst = mem.view(_p(op, "UCDIR_CONV_P_DST"), (B, H, W, dstC))
        dst[..., dstCoff:dstCoff + nout] = v
    sp = _p(op, "UCDIR_CONV_P_DST_STATS")
    if sp:
        st = mem.view(sp, (B, 2), torch.float64)
        vd = v.double().reshape(B, -1)
        st[:, 0] += vd.sum(1)
        st[:, 1] += (vd * vd).sum(1)


def emu_sgemm(op, mem):
    batch, M, N, Kd = (_i(op, "UCDIR_SGEMM_I_BATCH"), _i(op, "UCDIR_SGEMM_I_M"), _i(op, "UCDIR_SGEMM_I_N"),
                       _i(op, "UCDIR_SGEMM_I_K"))
    lda, ldb, ldc = _i(op, "UCDIR_SGEMM_I_LDA"), _i(op, "UCDIR_SGEMM_I_LDB"), _i(op, "UCDIR_SGEMM_I_LDC")
    sa, sb, sc = _i(op, "UCDIR_SGEMM_I_SA"), _i(op, "UCDIR_SGEMM_I_SB"), _i(op, "UCDIR_SGEMM_I_SC")
    tb = _i(op, "UCDIR_SGEMM_I_TRANSB")
    alpha = _f(op, "UCDIR_SGEMM_F_ALPHA")
    ta = torch.bfloat16 if _i(op, "UCDIR_SGEMM_I_A_BF16") else torch.float32
    tbt = torch.bfloat16 if _i(op, "UCDIR_SGEMM_I_B_BF16") else torch.float32
    tc = torch.bfloat16 if _i(op, "UCDIR_SGEMM_I_C_BF16") else torch.float32
    ea, eb, ec = (2 if t == torch.bfloat16 else 4 for t in (ta, tbt, tc))
    for b in range(batch):
        A = mem.view(_p(op, "UCDIR_SGEMM_P_A") + b * sa * ea, ((M - 1) * lda + Kd,), ta)
        A = torch.as_strided(A, (M, Kd), (lda, 1)).float()
        if tb:
            Bm = mem.view(_p(op, "UCDIR_SGEMM_P_B") + b * sb * eb, ((N - 1) * ldb + Kd,), tbt)
            Bm = torch.as_strided(Bm, (N, Kd), (ldb, 1)).t().float()
        else:
            Bm = mem.view(_p(op, "UCDIR_SGEMM_P_B") + b * sb * eb, ((Kd - 1) * ldb + N,), tbt)
            Bm = torch.as_strided(Bm, (Kd, N), (ldb, 1)).float()
        Cm = mem.view(_p(op, "UCDIR_SGEMM_P_C") + b * sc * ec, ((M - 1) * ldc + N,), tc)
        torch.as_strided(Cm, (M, N), (ldc, 1)).copy_((alpha * (A @ Bm)).to(tc))


def emu_softmax(op, mem):
    rows, cols = _i(op, "UCDIR_SOFTMAX_I_ROWS"), _i(op, "UCDIR_SOFTMAX_I_COLS")
    in_ld = _i(op, "UCDIR_SOFTMAX_I_IN_LD") or cols
    X = mem.view(_p(op, "UCDIR_SOFTMAX_P_X"), (rows, in_ld))[:, :cols]
    o = _p(op, "UCDIR_SOFTMAX_P_OUT_BF16")
    if o:
        out_ld = _i(op, "UCDIR_SOFTMAX_I_OUT_LD")
        P = mem.view(o, (rows, out_ld), torch.bfloat16)
        P.zero_()
        P[:, :cols] = torch.softmax(X, dim=-1).to(torch.bfloat16)
    else:
        X.copy_(torch.softmax(X, dim=-1))


def emu_guidance(op, mem):
    B, GH, GW, H, W = (_i(op, "UCDIR_GUID_I_B"), _i(op, "UCDIR_GUID_I_GH"), _i(op, "UCDIR_GUID_I_GW"),
                       _i(op, "UCDIR_GUID_I_H"), _i(op, "UCDIR_GUID_I_W"))
    g = mem.view(_p(op, "UCDIR_GUID_P_GUIDE"), (B, GH, GW, 4))[..., :3].permute(0, 3, 1, 2)
    r = GW // W
    if r > 1:
        off = r // 2 - 1
        g = 0.25 * (g[:, :, off::r, off::r] + g[:, :, off::r, off + 1::r] + g[:, :, off + 1::r, off::r]
                    + g[:, :, off + 1::r, off + 1::r])
        assert g.shape[-2:] == (H, W)
    w0 = mem.view(_p(op, "UCDIR_GUID_P_W0"), (16, 3, 1, 1)); b0 = mem.view(_p(op, "UCDIR_GUID_P_B0"), (16,))
    w2 = mem.view(_p(op, "UCDIR_GUID_P_W2"), (8, 8, 3, 3)); b2 = mem.view(_p(op, "UCDIR_GUID_P_B2"), (8,))
    u = F.conv2d(g, w0, b0)
    gate = u[:, :8] * u[:, 8:]
    out = F.conv2d(gate, w2, b2, padding=1)
    mem.view(_p(op, "UCDIR_GUID_P_DST"), (B, H, W, 8)).copy_(out.permute(0, 2, 3, 1))


def emu_time_embed(op, mem):
    L, nblk, inner = _i(op, "UCDIR_TEMB_I_L"), _i(op, "UCDIR_TEMB_I_NBLK"), _i(op, "UCDIR_TEMB_I_INNER")
    lp = _p(op, "UCDIR_TEMB_P_LEVELS")
    levels = mem.view(lp, (L,)).clone() if lp else torch.full((L,), _f(op, "UCDIR_TEMB_F_LEVEL"), dtype=torch.float32)
    count = inner // 2
    step = torch.arange(count, dtype=torch.float32) / count
    e = levels.view(L, 1) * torch.exp(-math.log(1e4) * step.view(1, -1))
    enc = torch.cat([torch.sin(e), torch.cos(e)], dim=-1)
    w1 = mem.view(_p(op, "UCDIR_TEMB_P_W1"), (4 * inner, inner)); b1 = mem.view(_p(op, "UCDIR_TEMB_P_B1"), (4 * inner,))
    w2 = mem.view(_p(op, "UCDIR_TEMB_P_W2"), (inner, 4 * inner)); b2 = mem.view(_p(op, "UCDIR_TEMB_P_B2"), (inner,))
    t = F.linear(swish(F.linear(enc, w1, b1)), w2, b2)
    rec = 8 * inner + 8 + 64 + 8
    blk = mem.view(_p(op, "UCDIR_TEMB_P_BLK"), (nblk, rec))
    dst = mem.view(_p(op, "UCDIR_TEMB_P_DST"), (L, nblk, 8))
    for k in range(nblk):
        wa = blk[k, :8 * inner].view(8, inner); ba = blk[k, 8 * inner:8 * inner + 8]
        wb = blk[k, 8 * inner + 8:8 * inner + 72].view(8, 8); bb = blk[k, 8 * inner + 72:]
        dst[:, k] = F.linear(swish(F.linear(t, wa, ba)), wb, bb)


def _reflect(q, n):
    q = np.abs(q)
    return np.where(q >= n, 2 * (n - 1) - q, q)


def emu_gather(op, mem):
    BT, TH, TW = _i(op, "UCDIR_GATHER_I_BT"), _i(op, "UCDIR_GATHER_I_TH"), _i(op, "UCDIR_GATHER_I_TW")
    IH, IW, PD = _i(op, "UCDIR_GATHER_I_IMG_H"), _i(op, "UCDIR_GATHER_I_IMG_W"), _i(op, "UCDIR_GATHER_I_PD")
    CA, CB, CD = _i(op, "UCDIR_GATHER_I_CA"), _i(op, "UCDIR_GATHER_I_CB"), _i(op, "UCDIR_GATHER_I_CD")
    mode = _i(op, "UCDIR_GATHER_I_OUT_BF16")
    odt = torch.bfloat16 if mode else torch.float32
    tab = mem.view(_p(op, "UCDIR_GATHER_P_TAB"), (BT, 3), torch.int32).numpy()
    nimg = int(tab[:, 0].max()) + 1
    A = mem.view(_p(op, "UCDIR_GATHER_P_SRC_A"), (nimg, CA, IH, IW))
    Bs = mem.view(_p(op, "UCDIR_GATHER_P_SRC_B"), (nimg, CB, IH, IW)) if CB else None
    dst = mem.view(_p(op, "UCDIR_GATHER_P_DST"), (BT, TH, TW, 2 * CD if mode == 2 else CD), odt)
    for t in range(BT):
        img, y0, x0 = (int(v) for v in tab[t])
        sy = torch.from_numpy(_reflect(np.arange(TH) + y0 - PD, IH))
        sx = torch.from_numpy(_reflect(np.arange(TW) + x0 - PD, IW))
        dst[t].zero_()
        dst[t, :, :, :CA] = A[img][:, sy][:, :, sx].permute(1, 2, 0).to(odt)
        if CB:
            dst[t, :, :, CA:CA + CB] = Bs[img][:, sy][:, :, sx].permute(1, 2, 0).to(odt)
        if mode == 2:                                             # lo planes
            va = A[img][:, sy][:, :, sx].permute(1, 2, 0)
            dst[t, :, :, CD:CD + CA] = (va - va.to(odt).float()).to(odt)
            if CB:
                vb = Bs[img][:, sy][:, :, sx].permute(1, 2, 0)
                dst[t, :, :, CD + CA:CD + CA + CB] = (vb - vb.to(odt).float()).to(odt)


def emu_scatter(op, mem):
    BI, IH, IW = _i(op, "UCDIR_SCATTER_I_BIMG"), _i(op, "UCDIR_SCATTER_I_IMG_H"), _i(op, "UCDIR_SCATTER_I_IMG_W")
    NTY, NTX, TH, TW = (_i(op, "UCDIR_SCATTER_I_NTY"), _i(op, "UCDIR_SCATTER_I_NTX"), _i(op, "UCDIR_SCATTER_I_TH"),
                        _i(op, "UCDIR_SCATTER_I_TW"))
    PD, CE, mode, clip, Cc = (_i(op, "UCDIR_SCATTER_I_PD"), _i(op, "UCDIR_SCATTER_I_CE"), _i(op, "UCDIR_SCATTER_I_MODE"),
                              _i(op, "UCDIR_SCATTER_I_CLIP"), _i(op, "UCDIR_SCATTER_I_C"))
    eps_t = mem.view(_p(op, "UCDIR_SCATTER_P_EPS"), (BI * NTY * NTX, TH, TW, CE))
    oy = mem.view(_p(op, "UCDIR_SCATTER_P_OWNER_Y"), (IH,), torch.int32).long()
    ox = mem.view(_p(op, "UCDIR_SCATTER_P_OWNER_X"), (IW,), torch.int32).long()
    y0 = mem.view(_p(op, "UCDIR_SCATTER_P_Y0"), (NTY,), torch.int32).long()
    x0 = mem.view(_p(op, "UCDIR_SCATTER_P_X0"), (NTX,), torch.int32).long()
    out = mem.view(_p(op, "UCDIR_SCATTER_P_OUT"), (BI, Cc, IH, IW))
    yy = torch.arange(IH); xx = torch.arange(IW)
    py = yy + PD - y0[oy.clamp_min(0)]; px = xx + PD - x0[ox.clamp_min(0)]
    e = torch.zeros((BI, Cc, IH, IW))
    for img in range(BI):
        tile = (img * NTY + oy.clamp_min(0)).view(-1, 1) * NTX + ox.clamp_min(0).view(1, -1)
        v = eps_t[tile, py.view(-1, 1), px.view(1, -1)][..., :Cc]                    # IH, IW, C
        v = torch.where(((oy >= 0).view(-1, 1) & (ox >= 0).view(1, -1)).unsqueeze(-1), v, torch.zeros_like(v))
        e[img] = v.permute(2, 0, 1)
    if mode == 0:
        out.copy_(e)
        return
    xt = mem.view(_p(op, "UCDIR_SCATTER_P_XT"), (BI, Cc, IH, IW))
    pp = _p(op, "UCDIR_SCATTER_P_PARAMS")
    npz = _p(op, "UCDIR_SCATTER_P_NOISE")
    if pp:
        pv = mem.view(pp, (8,))
        names = ["A", "B", "C1", "C2", "SIGMA", "_clip", "_noise", "C3"]
        f = lambda n: pv[names.index(n)].clone()
        clip = bool(pv[5] != 0)
        if pv[6] == 0:
            npz = 0
    else:
        f = lambda n: torch.tensor(_f(op, "UCDIR_SCATTER_F_" + n), dtype=torch.float32)
    x0_ = f("A") * xt - f("B") * e
    if clip:
        x0_ = x0_.clamp(-1.0, 1.0)
    mean = f("C1") * x0_ + f("C2") * xt
    if float(f("C3")) != 0.0:
        mean = mean + f("C3") * e
    z = mem.view(npz, (BI, Cc, IH, IW)) if npz else torch.zeros_like(xt)
    out.copy_(mean + z * f("SIGMA"))


def emu_maxpool(op, mem):
    B, H, W, Cc = _i(op, "UCDIR_POOL_I_B"), _i(op, "UCDIR_POOL_I_H"), _i(op, "UCDIR_POOL_I_W"), _i(op, "UCDIR_POOL_I_C")
    if _i(op, "UCDIR_POOL_I_SPLIT"):                                # (hi, lo) bf16 plane pairs
        bf = torch.bfloat16
        s2 = mem.view(_p(op, "UCDIR_POOL_P_SRC"), (B, 2 * H, 2 * W, 2 * Cc), bf).float()
        v = F.max_pool2d((s2[..., :Cc] + s2[..., Cc:]).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
        d2 = mem.view(_p(op, "UCDIR_POOL_P_DST"), (B, H, W, 2 * Cc), bf)
        d2[..., :Cc] = v.to(bf)
        d2[..., Cc:] = (v - v.to(bf).float()).to(bf)
        return
    s = mem.view(_p(op, "UCDIR_POOL_P_SRC"), (B, 2 * H, 2 * W, Cc))
    d = mem.view(_p(op, "UCDIR_POOL_P_DST"), (B, H, W, Cc))
    d.copy_(F.max_pool2d(s.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))


def emu_memset(op, mem):
    n = int(op.i[0]) + (int(op.i[1]) << 31)
    mem.view(int(op.p[0]), (n,), torch.uint8).zero_()


def emu_tc_conv(op, mem):
    """UCDIR_OP_TC_CONV restated: bf16 operands, fp32 accumulation, epilogue exactly as documented in the header."""
    if _i(op, "UCDIR_TC_I_PHASES") == 4:                         # fused upsample phases = the four single-phase records
        ntot, kc = _i(op, "UCDIR_TC_I_NTOT"), _i(op, "UCDIR_TC_I_KC")
        ktot = 4 * (_i(op, "UCDIR_TC_I_C0") + _i(op, "UCDIR_TC_I_C1")) * (3 if _i(op, "UCDIR_TC_I_SPLIT") else 1)
        for ph in range(4):
            sub = _lib.Op.from_buffer_copy(op)
            sub.i[K["UCDIR_TC_I_PHASES"]] = 0
            sub.p[K["UCDIR_TC_P_W"]] = _p(op, "UCDIR_TC_P_W") + ph * ntot * ktot * 2
            sub.i[K["UCDIR_TC_I_OY0"]], sub.i[K["UCDIR_TC_I_OX0"]] = (ph >> 1) - 1, (ph & 1) - 1
            sub.i[K["UCDIR_TC_I_DST_PY"]], sub.i[K["UCDIR_TC_I_DST_PX"]] = ph >> 1, ph & 1
            emu_tc_conv(sub, mem)
        return
    g = lambda n: _i(op, "UCDIR_TC_I_" + n)
    B, H, W, sH, sW = g("B"), g("H"), g("W"), g("SRC_H"), g("SRC_W")
    C0, C1, Ntot = g("C0"), g("C1"), g("NTOT")
    ncv = g("NCOL_VALID") or Ntot
    nty, ntx, oy0, ox0, stride, groups, KC, NT = g("NTY"), g("NTX"), g("OY0"), g("OX0"), g("STRIDE"), g("GROUPS"), g("KC"), g("NT")
    gn, ncls, act, mode, dst_f32 = g("GN"), g("NCLS"), g("ACT"), g("MODE"), g("DST_F32")
    dstC, dstCoff, dstUp, dpy, dpx, resC, aws = g("DST_C"), g("DST_COFF"), g("DST_UP"), g("DST_PY"), g("DST_PX"), g("RES_C"), g("ATTW_STRIDE")
    eps = _f(op, "UCDIR_TC_F_EPS")
    bf = torch.bfloat16
    Cin = C0 + C1
    split = g("SPLIT")                                           # fp32_tc: (hi, lo) plane pairs, value = hi + lo
    cstride = g("SRC_CSTRIDE") or (2 * C0 if split else C0)
    a_lo = (g("SRC_LO_OFF") or C0) if split else 0
    x = mem.view(_p(op, "UCDIR_TC_P_SRC0"), ((B * sH * sW - 1) * cstride + a_lo + C0,), bf)
    xs = lambda off: torch.as_strided(x, (B, sH, sW, C0), (sH * sW * cstride, sW * cstride, cstride, 1), off).float()
    x = xs(0) + xs(a_lo) if split else xs(0)
    if g("SRC_GN_SWISH"):                                        # final_conv: Swish(GroupNorm(x)) on the source, rounded to bf16
        s0 = mem.view(_p(op, "UCDIR_TC_P_STATS0"), (B, 2), torch.float64)
        cnt = float(C0 * sH * sW)
        mean = s0[:, 0] / cnt
        var = (s0[:, 1] / cnt - mean * mean).clamp_min(0)
        rstd = (1.0 / torch.sqrt(var + eps)).float().view(B, 1, 1, 1)
        gam = mem.view(_p(op, "UCDIR_TC_P_SRC_GAMMA"), (C0,)); bet = mem.view(_p(op, "UCDIR_TC_P_SRC_BETA"), (C0,))
        a = rstd * gam.view(1, 1, 1, -1)
        x = swish(x * a + (bet.view(1, 1, 1, -1) - a * mean.float().view(B, 1, 1, 1))).to(bf).float()
    wb = g("W_BATCHED")
    if C1 and split:
        x1 = mem.view(_p(op, "UCDIR_TC_P_SRC1"), (B, sH, sW, 2 * C1), bf).float()
        x = torch.cat([x, x1[..., :C1] + x1[..., C1:]], dim=-1)
    elif C1:
        x = torch.cat([x, mem.view(_p(op, "UCDIR_TC_P_SRC1"), (B, sH, sW, C1), bf).float()], dim=-1)
    KB = g("KB") or KC
    if groups > 1:
        Cg, Ng = Cin // groups, Ntot // groups
        cg_eff = max(Cg, KB)                                   # weight-row K per tap (pack_tc_grouped)
    else:
        Cg, Ng, cg_eff = Cin, Ntot, Cin
    ktap = cg_eff
    Ktot = nty * ntx * ktap
    if wb:                                                       # per-image weights [B][Ntot][C0], strided
        rs = g("W_ROWSTRIDE"); bs = g("W_BATCHSTRIDE_LO") + (g("W_BATCHSTRIDE_HI") << 31)
        n_rows = g("W_ROWS") or Ntot                              # rows beyond the image's extent are TMA zero fill
        b_lo = g("W_LO_OFF") if split else 0
        wraw = mem.view(_p(op, "UCDIR_TC_P_W"), ((B - 1) * bs + (n_rows - 1) * rs + b_lo + C0,), bf)
        wbat = torch.zeros(B, Ntot, C0)
        wbat[:, :n_rows] = torch.as_strided(wraw, (B, n_rows, C0), (bs, rs, 1)).float()
        if split:
            wbat[:, :n_rows] += torch.as_strided(wraw, (B, n_rows, C0), (bs, rs, 1), b_lo).float()
    elif split:
        # per tap [W_hi(s0) | W_hi(s0) | W_hi(s1) | W_hi(s1) | W_lo(s0) | W_lo(s1)] (grouped: [W_hi | W_hi | W_lo])
        w3 = mem.view(_p(op, "UCDIR_TC_P_W"), (Ntot, nty * ntx, 3 * ktap), bf).float()
        c0k = ktap if groups > 1 else C0
        c1k = ktap - c0k
        h0, h0b, h1, h1b, l0, l1 = torch.split(w3, [c0k, c0k, c1k, c1k, c0k, c1k], dim=-1)
        assert torch.equal(h0, h0b) and torch.equal(h1, h1b), "SPLIT weight rows: the two W_hi copies differ"
        wp = torch.cat([h0 + l0, h1 + l1], dim=-1)
    else:
        wp = mem.view(_p(op, "UCDIR_TC_P_W"), (Ntot, nty * ntx, ktap), bf).float()
    acc = torch.zeros(B, H, W, Ntot)
    ys = torch.arange(H) * stride
    xs = torch.arange(W) * stride
    for ty in range(nty):
        for tx in range(ntx):
            sy, sx = ys + ty + oy0, xs + tx + ox0
            vy, vx = (sy >= 0) & (sy < sH), (sx >= 0) & (sx < sW)
            patch = x[:, sy.clamp(0, sH - 1)][:, :, sx.clamp(0, sW - 1)]
            patch = patch * (vy.view(1, -1, 1, 1) & vx.view(1, 1, -1, 1))
            if wb:
                acc += torch.einsum("bhwc,bnc->bhwn", patch, wbat)
                continue
            for gi in range(groups):
                cb = (gi * Cg) // cg_eff * cg_eff if groups > 1 else 0
                rows = slice(gi * Ng, (gi + 1) * Ng)
                acc[..., rows] += patch[..., cb:cb + cg_eff] @ wp[rows, ty * ntx + tx].t()
    if g("RES_FUSED"):                                           # the block's 1x1 res_conv of the same (un-normalised) input
        b2 = mem.view(_p(op, "UCDIR_TC_P_TB2"), (Ntot,))
        rC = g("DST_RES_C")
        if split:                                                # weights [s0 hi | s0 hi | s1 hi | s1 hi | s0 lo | s1 lo], (hi, lo) output planes
            w23 = mem.view(_p(op, "UCDIR_TC_P_W2"), (Ntot, 3 * Cin), bf).float()
            h0, _, h1, _, l0, l1 = torch.split(w23, [C0, C0, C1, C1, C0, C1], dim=-1)
            w2 = torch.cat([h0 + l0, h1 + l1], dim=-1)
            r32 = x @ w2.t() + b2.view(1, 1, 1, -1)
            dres = mem.view(_p(op, "UCDIR_TC_P_DST_RES"), (B, H, W, 2 * rC), bf)
            dres[..., :Ntot] = r32.to(bf)
            dres[..., rC:rC + Ntot] = (r32 - r32.to(bf).float()).to(bf)
        else:
            w2 = mem.view(_p(op, "UCDIR_TC_P_W2"), (Ntot, Cin), bf).float()
            dres = mem.view(_p(op, "UCDIR_TC_P_DST_RES"), (B, H, W, rC), bf)
            dres[..., :Ntot] = (x @ w2.t() + b2.view(1, 1, 1, -1)).to(bf)
    tb = mem.view(_p(op, "UCDIR_TC_P_TB"), (ncls if gn else 1, Ntot))
    if gn:
        s0 = mem.view(_p(op, "UCDIR_TC_P_STATS0"), (B, 2), torch.float64).clone()
        if C1:
            s0 = s0 + mem.view(_p(op, "UCDIR_TC_P_STATS1"), (B, 2), torch.float64)
        cnt = float(Cin * sH * sW)
        mean = s0[:, 0] / cnt
        var = (s0[:, 1] / cnt - mean * mean).clamp_min(0)
        rstd = (1.0 / torch.sqrt(var + eps)).float().view(B, 1, 1, 1)
        mr = mean.float().view(B, 1, 1, 1) * rstd
        tg = mem.view(_p(op, "UCDIR_TC_P_TG"), (ncls, Ntot))
        if ncls == 9:
            cy = torch.ones(H, dtype=torch.long); cy[0] = 0; cy[-1] = 2
            cx = torch.ones(W, dtype=torch.long); cx[0] = 0; cx[-1] = 2
            cls = cy.view(-1, 1) * 3 + cx.view(1, -1)
        else:
            cls = torch.zeros(H, W, dtype=torch.long)
        v = acc * rstd - mr * tg[cls].unsqueeze(0) + tb[cls].unsqueeze(0)
    else:
        alpha = float(op.f[K["UCDIR_TC_F_ALPHA"]]) or 1.0
        v = acc * alpha + tb[0].view(1, 1, 1, -1)
    if mode == 1:
        att = mem.view(_p(op, "UCDIR_TC_P_ATT"), (B, H, W, 8))
        base = _p(op, "UCDIR_TC_P_ATTW")
        aw = torch.stack([mem.view(base + b * aws * 4, (8,)) for b in range(B)])
        a = att * aw.view(B, 1, 1, 8)
        c = Ntot // 8
        h = (v.reshape(B, H, W, c, 8) * a.unsqueeze(3)).sum(-1)
        if split:
            r2 = mem.view(_p(op, "UCDIR_TC_P_RES"), (B, H, W, 2 * resC), bf).float()
            res = r2[..., :c] + r2[..., resC:resC + c]
        else:
            res = mem.view(_p(op, "UCDIR_TC_P_RES"), (B, H, W, resC), bf)[..., :c].float()
        out = swish(h) + res
        nout = c
    else:
        if act == 1:
            v = swish(v)
        elif act == 2:
            v = torch.maximum(0.2 * v, v)
        rp = _p(op, "UCDIR_TC_P_RES")
        if rp and split:
            r2 = mem.view(rp, (B, H, W, 2 * resC), bf).float()
            v = v + r2[..., :Ntot] + r2[..., resC:resC + Ntot]
        elif rp:
            v = v + mem.view(rp, (B, H, W, resC), bf)[..., :Ntot].float()
        out = v[..., :ncv]
        nout = ncv
    odt = torch.float32 if dst_f32 else bf
    pair = bool(split and not dst_f32)                           # (hi, lo) output planes
    lo_of = lambda t: (t - t.to(bf).float()).to(bf)
    out32 = out
    out = out.to(odt)
    d2 = _p(op, "UCDIR_TC_P_DST2")
    if d2:                                                       # columns >= T_COL0 go transposed to DST2[b][col][pixel]
        t0, tld = g("T_COL0"), g("T_LD")
        vt = mem.view(d2, (B, Ntot - t0, 2 * tld if pair else tld), bf)
        tr = out32[..., t0:].reshape(B, H * W, Ntot - t0).transpose(1, 2)
        vt[:, :, :H * W] = tr.to(bf)
        if pair:
            vt[:, :, tld:tld + H * W] = lo_of(tr)
        out, out32 = out[..., :t0], out32[..., :t0]
        nout = t0
    pitch = 2 * dstC if pair else dstC
    crop = g("DST_CROP")
    if crop:                                                     # fused final conv: only the interior, compact destination
        dst = mem.view(_p(op, "UCDIR_TC_P_DST"), (B, H - 2 * crop, W - 2 * crop, pitch), odt)
        dst[..., dstCoff:dstCoff + nout] = out[:, crop:H - crop, crop:W - crop]
        return
    if dstUp:
        dst = mem.view(_p(op, "UCDIR_TC_P_DST"), (B, 2 * H, 2 * W, pitch), odt)
        dst[:, dpy::2, dpx::2, dstCoff:dstCoff + nout] = out
        if pair:
            dst[:, dpy::2, dpx::2, dstC + dstCoff:dstC + dstCoff + nout] = lo_of(out32)
    else:
        dst = mem.view(_p(op, "UCDIR_TC_P_DST"), (B, H, W, pitch), odt)
        dst[..., dstCoff:dstCoff + nout] = out
        if pair:
            dst[..., dstC + dstCoff:dstC + dstCoff + nout] = lo_of(out32)
    if pair:
        out = out32                                              # statistics of the fp32 values
    sp = _p(op, "UCDIR_TC_P_DST_STATS")
    if sp:
        st = mem.view(sp, (B, 2), torch.float64)
        vd = out.double().reshape(B, -1)
        st[:, 0] += vd.sum(1)
        st[:, 1] += (vd * vd).sum(1)


def emu_tc_attn(op, mem):
    """UCDIR_OP_TC_ATTN restated: bf16 Q / K / V^T, fp32 scores and softmax, bf16 probabilities (unnormalised in the kernel,
    normalised here: same values up to rounding), fp32 accumulation, bf16 output."""
    g = lambda n: _i(op, "UCDIR_ATTN_I_" + n)
    B, N, Cc, qk_ld, vt_ld, o_ld = g("B"), g("N"), g("C"), g("QK_LD"), g("VT_LD"), g("O_LD")
    bf = torch.bfloat16
    qk = mem.view(_p(op, "UCDIR_ATTN_P_QK"), (B, N, qk_ld), bf).float()
    vt = mem.view(_p(op, "UCDIR_ATTN_P_VT"), (B, Cc, vt_ld), bf).float()[:, :, :N]
    scale = float(op.f[K["UCDIR_ATTN_F_SCALE"]]) or 1.0 / math.sqrt(Cc)
    q, k = qk[..., :Cc], qk[..., Cc:2 * Cc]
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) * scale, dim=-1).to(bf).float()
    o = torch.bmm(att, vt.transpose(1, 2))
    mem.view(_p(op, "UCDIR_ATTN_P_O"), (B, N, o_ld), bf)[..., :Cc] = o.to(bf)


def emu_gn_apply(op, mem):
    B, HW, Cc, sw = _i(op, "UCDIR_GNA_I_B"), _i(op, "UCDIR_GNA_I_HW"), _i(op, "UCDIR_GNA_I_C"), _i(op, "UCDIR_GNA_I_SWISH")
    bf = torch.bfloat16
    split = _i(op, "UCDIR_GNA_I_SPLIT")
    if split:
        x2 = mem.view(_p(op, "UCDIR_GNA_P_SRC"), (B, HW, 2 * Cc), bf).float()
        x = x2[..., :Cc] + x2[..., Cc:]
    else:
        x = mem.view(_p(op, "UCDIR_GNA_P_SRC"), (B, HW, Cc), bf).float()
    st = mem.view(_p(op, "UCDIR_GNA_P_STATS"), (B, 2), torch.float64)
    cnt = float(HW * Cc)
    mean = st[:, 0] / cnt
    var = (st[:, 1] / cnt - mean * mean).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + float(op.f[0]))).float().view(B, 1, 1)
    gamma = mem.view(_p(op, "UCDIR_GNA_P_GAMMA"), (Cc,)); beta = mem.view(_p(op, "UCDIR_GNA_P_BETA"), (Cc,))
    y = (x - mean.float().view(B, 1, 1)) * rstd * gamma + beta
    if sw:
        y = swish(y)
    if split:
        d = mem.view(_p(op, "UCDIR_GNA_P_DST"), (B, HW, 2 * Cc), bf)
        d[..., :Cc] = y.to(bf)
        d[..., Cc:] = (y - y.to(bf).float()).to(bf)
        return
    mem.view(_p(op, "UCDIR_GNA_P_DST"), (B, HW, Cc), bf).copy_(y.to(bf))


def emu_cast(op, mem):
    n = int(op.i[0]) + (int(op.i[1]) << 31)
    if int(op.i[2]) == 2:                                         # fp32 rows -> (hi, lo) bf16 plane pairs
        cols, in_ld, out_ld = int(op.i[3]), int(op.i[4]), int(op.i[5])
        src = mem.view(int(op.p[0]), (n, in_ld))[:, :cols]
        dst = mem.view(int(op.p[1]), (n, 2 * out_ld), torch.bfloat16)
        dst.zero_()
        dst[:, :cols] = src.to(torch.bfloat16)
        dst[:, out_ld:out_ld + cols] = (src - src.to(torch.bfloat16).float()).to(torch.bfloat16)
        return
    if int(op.i[2]) == 0:
        mem.view(int(op.p[1]), (n,), torch.bfloat16).copy_(mem.view(int(op.p[0]), (n,)).to(torch.bfloat16))
    else:
        mem.view(int(op.p[1]), (n,)).copy_(mem.view(int(op.p[0]), (n,), torch.bfloat16).float())


def emu_crop(op, mem):
    g = lambda n: _i(op, "UCDIR_CROP_I_" + n)
    BT, TH, TW, IH, IW, OY, OX = g("BT"), g("TH"), g("TW"), g("IH"), g("IW"), g("OY"), g("OX")
    src = mem.view(_p(op, "UCDIR_CROP_P_SRC"), (BT, TH, TW, 4))
    mem.view(_p(op, "UCDIR_CROP_P_DST"), (BT, IH, IW, 4)).copy_(src[:, OY:OY + IH, OX:OX + IW])


def emu_gn_stats(op, mem):
    g = lambda n: _i(op, "UCDIR_GNS_I_" + n)
    B, HW, Cc, G = g("B"), g("HW"), g("C"), g("G")
    x = mem.view(_p(op, "UCDIR_GNS_P_SRC"), (B, HW, G, Cc // G)).double()
    st = mem.view(_p(op, "UCDIR_GNS_P_STATS"), (B, G, 2), torch.float64)
    st[..., 0] = x.sum((1, 3)); st[..., 1] = (x * x).sum((1, 3))


def emu_gn_apply_f32(op, mem):
    g = lambda n: _i(op, "UCDIR_GNS_I_" + n)
    B, HW, Cc, G, sw = g("B"), g("HW"), g("C"), g("G"), g("SWISH")
    x = mem.view(_p(op, "UCDIR_GNF_P_SRC"), (B, HW, G, Cc // G))
    st = mem.view(_p(op, "UCDIR_GNF_P_STATS"), (B, G, 2), torch.float64)
    cnt = float(HW * (Cc // G))
    mean = st[..., 0] / cnt
    var = (st[..., 1] / cnt - mean * mean).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + float(op.f[0]))).float().view(B, 1, G, 1)
    y = ((x - mean.float().view(B, 1, G, 1)) * rstd).reshape(B, HW, Cc)
    y = y * mem.view(_p(op, "UCDIR_GNF_P_GAMMA"), (Cc,)) + mem.view(_p(op, "UCDIR_GNF_P_BETA"), (Cc,))
    mem.view(_p(op, "UCDIR_GNF_P_DST"), (B, HW, Cc)).copy_(swish(y) if sw else y)


def emu_layout(op, mem):
    B, Cc, HW, d = int(op.i[0]), int(op.i[1]), int(op.i[2]), int(op.i[3])
    if d == 0:
        mem.view(int(op.p[1]), (B, HW, Cc)).copy_(mem.view(int(op.p[0]), (B, Cc, HW)).transpose(1, 2))
    else:
        mem.view(int(op.p[1]), (B, Cc, HW)).copy_(mem.view(int(op.p[0]), (B, HW, Cc)).transpose(1, 2))


def emu_to_image(op, mem):
    g = lambda n: _i(op, "UCDIR_IMG_I_" + n)
    B, C, H, W, PD = g("B"), g("C"), g("H"), g("W"), g("PD")
    lo, hi = np.float32(_f(op, "UCDIR_IMG_F_MIN")), np.float32(_f(op, "UCDIR_IMG_F_MAX"))
    src = mem.view(_p(op, "UCDIR_IMG_P_SRC"), (B, C, H, W))[:, :, PD:H - PD, PD:W - PD]
    v = ((src.clamp(float(lo), float(hi)) - float(lo)) / float(hi - lo)).numpy()            # fp32, as core/metrics.py:14-16
    img = (v * np.float32(255.0)).round().astype(np.uint8)                                  # :29-31 (numpy rounds half to even)
    mem.view(_p(op, "UCDIR_IMG_P_DST"), (B, H - 2 * PD, W - 2 * PD, C), torch.uint8).copy_(torch.from_numpy(img).permute(0, 2, 3, 1))


DISPATCH = {
    K["UCDIR_OP_CONV_F32"]: emu_conv, K["UCDIR_OP_SGEMM_F32"]: emu_sgemm, K["UCDIR_OP_SOFTMAX_F32"]: emu_softmax,
    K["UCDIR_OP_GUIDANCE"]: emu_guidance, K["UCDIR_OP_TIME_EMBED"]: emu_time_embed,
    K["UCDIR_OP_GATHER_TILES"]: emu_gather, K["UCDIR_OP_SCATTER"]: emu_scatter, K["UCDIR_OP_MAXPOOL2"]: emu_maxpool,
    K["UCDIR_OP_MEMSET"]: emu_memset, K["UCDIR_OP_TC_CONV"]: emu_tc_conv, K["UCDIR_OP_GN_APPLY_BF16"]: emu_gn_apply,
    K["UCDIR_OP_CAST"]: emu_cast, K["UCDIR_OP_CROP_TILES"]: emu_crop, K["UCDIR_OP_GN_STATS_F32"]: emu_gn_stats,
    K["UCDIR_OP_GN_APPLY_F32"]: emu_gn_apply_f32, K["UCDIR_OP_LAYOUT"]: emu_layout, K["UCDIR_OP_TO_IMAGE_U8"]: emu_to_image,
    K["UCDIR_OP_TC_ATTN"]: emu_tc_attn,
}

LAUNCHED = []


def run_ops(ops, n, stream=0):
    """Drop-in for ucdir_b200._lib.run_ops (tests monkeypatch the engine's runner with this)."""
    mem = Memory()
    with torch.no_grad():
        for k in range(n):
            op = ops[k]
            LAUNCHED.append(int(op.kind))
            DISPATCH[int(op.kind)](op, mem)
