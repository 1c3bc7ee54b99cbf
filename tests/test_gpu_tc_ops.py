"""Per-op GPU tests of the tcgen05 / TMA kernels: the same `ucdir_op_t` record is executed by the CUDA library
on device copies and by the CPU interpreter (tests/op_emulator.py) on host copies of identical bf16 inputs.
Differences are then only accumulation order (fp32) and one bf16 rounding of the result."""
import numpy as np
import pytest
import torch

from tests import op_emulator
from ucdir_b200 import _lib, engine as E

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


class Case:
    """Named host tensors + a function that builds the op list from Act/pointer views on a given device."""

    def __init__(self):
        self.t = {}

    def add(self, name, tensor):
        self.t[name] = tensor.contiguous()
        return self

    def on(self, dev):
        return {k: v.clone().to(dev).contiguous() for k, v in self.t.items()}


def run_both(case, build):
    host = case.on("cpu")
    ol = build(host)
    op_emulator.run_ops(ol.array(), len(ol))
    devt = case.on("cuda")
    ol2 = build(devt)
    _lib.run_ops(ol2.array(), len(ol2), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return host, {k: v.cpu() for k, v in devt.items()}


def assert_close(got, want, what, rtol=2e-2, atol=2e-2):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    scale = want.abs().max().item() + 1e-6
    mask = err > atol * scale + rtol * want.abs()
    bad = mask.float().mean().item()
    where = mask.nonzero()[:12].tolist() if bad else []
    assert bad == 0.0, (f"{what}: max err {err.max().item():.4e} (scale {scale:.3e}), {bad:.3%} elements out of tolerance; "
                        f"first bad indices {where}, got {[round(got[tuple(i)].item(), 4) for i in where[:6]]} "
                        f"want {[round(want[tuple(i)].item(), 4) for i in where[:6]]}")


def act(t, C, H, W, stats=None):
    return E.Act(t, C, H, W, stats.data_ptr() if stats is not None else 0, True)


def rnd(g, *shape, scale=1.0):
    return torch.randn(*shape, generator=g) * scale


def stats_of(x_bf16):
    B = x_bf16.shape[0]
    v = x_bf16.double().reshape(B, -1)
    return torch.stack([v.sum(1), (v * v).sum(1)], dim=1).contiguous()


def dense_case(seed, B, H, W, C0, C1, Cout, ks, gn, act_, res, stride=1, row3=0, halo=0, kc=64, fuse_res=False):
    g = torch.Generator().manual_seed(seed)
    sH, sW = H * stride, W * stride
    c = Case()
    c.add("x0", (rnd(g, B, sH, sW, C0) + 0.3).to(BF))
    if C1:
        c.add("x1", (rnd(g, B, sH, sW, C1) * 0.7 - 0.2).to(BF))
    Cin = C0 + C1
    w = rnd(g, Cout, Cin, ks, ks, scale=1.0 / np.sqrt(Cin * ks * ks))
    bias = rnd(g, Cout, scale=0.1)
    gamma = 1 + 0.3 * rnd(g, Cin) if gn else None
    beta = 0.2 * rnd(g, Cin) if gn else None
    nt = E._tc_nt(Cout)
    wp, tb, tg = E.pack_tc_dense(w, bias, nt, gamma, beta)
    c.add("w", wp).add("tb", tb)
    if tg is not None:
        c.add("tg", tg)
    if gn:
        c.add("s0", stats_of(c.t["x0"]))
        if C1:
            c.add("s1", stats_of(c.t["x1"]))
    if res:
        c.add("res", rnd(g, B, H, W, Cout).to(BF))
    c.add("dst", torch.zeros(B, H, W, Cout, dtype=BF)).add("dstats", torch.zeros(B, 2, dtype=torch.float64))
    if fuse_res:                                   # the block's 1x1 res_conv of the same input, computed by the same launch
        w2p, tb2, _ = E.pack_tc_dense(rnd(g, Cout, Cin, 1, 1, scale=1.0 / np.sqrt(Cin)), rnd(g, Cout, scale=0.1), nt)
        c.add("w2", w2p).add("tb2", tb2).add("dres", torch.zeros(B, H, W, Cout, dtype=BF))

    def build(t):
        ol = E.OpList()
        extra = dict(w2=t["w2"].data_ptr(), tb2=t["tb2"].data_ptr(), dst_res=act(t["dres"], Cout, H, W)) if fuse_res else {}
        E._tc_op(ol, **extra, src0=act(t["x0"], C0, sH, sW, t.get("s0")), src1=act(t["x1"], C1, sH, sW, t.get("s1")) if C1 else None,
                 w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), tg=t["tg"].data_ptr() if "tg" in t else 0, gn=1 if gn else 0,
                 ncls=9 if (gn and ks == 3) else 1, nty=ks, ntx=ks, oy0=-(ks // 2), ox0=-(ks // 2), stride=stride, act=act_,
                 res=act(t["res"], Cout, H, W) if res else None, dst=act(t["dst"], Cout, H, W, t["dstats"]), ntot=Cout, B=B, nt=nt,
                 row3=row3, halo=halo, kc=kc)
        return ol

    def torch_ref():
        """The same layer through torch.nn.functional in the reference's own form (GroupNorm(1, C) -> conv -> Swish -> + residual,
        model/ucdir.py:109-111) on the same bf16 inputs, fp32 weights: independent of the op interpreter and of the weight packer."""
        F = torch.nn.functional
        x = c.t["x0"].float()
        if C1:
            x = torch.cat([x, c.t["x1"].float()], dim=-1)
        x = x.permute(0, 3, 1, 2)
        if gn:
            x = F.group_norm(x, 1, gamma, beta, eps=1e-5)
        y = F.conv2d(x, w, bias, stride=stride, padding=ks // 2)
        if act_ == 1:
            y = y * torch.sigmoid(y)
        y = y.permute(0, 2, 3, 1)
        if res:
            y = y + c.t["res"].float()
        return y
    build.torch_ref = torch_ref
    return c, build


@pytest.mark.parametrize("cfg", [
    dict(seed=1, B=2, H=16, W=16, C0=64, C1=0, Cout=64, ks=1, gn=False, act_=0, res=False),
    dict(seed=2, B=3, H=24, W=20, C0=128, C1=0, Cout=128, ks=3, gn=True, act_=1, res=False),
    dict(seed=3, B=2, H=16, W=16, C0=128, C1=64, Cout=64, ks=3, gn=True, act_=1, res=False),
    dict(seed=4, B=1, H=32, W=32, C0=256, C1=0, Cout=256, ks=3, gn=False, act_=0, res=True),
    dict(seed=5, B=5, H=8, W=8, C0=512, C1=512, Cout=512, ks=3, gn=True, act_=1, res=False),
    dict(seed=6, B=2, H=16, W=16, C0=64, C1=0, Cout=64, ks=3, gn=False, act_=0, res=False, stride=2),
    dict(seed=7, B=2, H=18, W=18, C0=512, C1=0, Cout=1536, ks=1, gn=True, act_=0, res=False),
    dict(seed=8, B=1, H=128, W=128, C0=64, C1=0, Cout=64, ks=3, gn=True, act_=1, res=False),
], ids=lambda c: "C%d+%d_%d_k%d_s%d_%dx%d" % (c["C0"], c["C1"], c["Cout"], c["ks"], c.get("stride", 1), c["H"], c["W"]))
def test_tc_dense(cfg):
    c, build = dense_case(**cfg)
    host, dev = run_both(c, build)
    assert_close(dev["dst"], host["dst"], "dst")
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=2e-3, atol=2e-3)
    assert_close(dev["dst"], build.torch_ref(), "dst vs torch.nn.functional", rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("cfg", [
    dict(seed=21, B=2, H=16, W=16, C0=64, C1=0, Cout=64),
    dict(seed=22, B=1, H=128, W=128, C0=64, C1=0, Cout=64),
    dict(seed=23, B=2, H=40, W=72, C0=128, C1=64, Cout=64),       # partial super tiles in both directions, concat input
    dict(seed=24, B=3, H=64, W=64, C0=256, C1=128, Cout=128),
    dict(seed=25, B=2, H=24, W=20, C0=128, C1=0, Cout=128),
    dict(seed=26, B=5, H=48, W=96, C0=64, C1=0, Cout=64, act_=0),   # 90 items: most CTAs get one, images change inside a CTA's range
    dict(seed=27, B=2, H=40, W=72, C0=16, C1=0, Cout=64, gn=False, act_=0, kc=16),     # the in-conv: 32-byte pixel rows, no GroupNorm
    dict(seed=28, B=1, H=128, W=128, C0=16, C1=0, Cout=64, gn=False, act_=0, kc=16),
    dict(seed=29, B=2, H=40, W=72, C0=64, C1=64, Cout=64, fuse_res=True),      # + fused 1x1 res_conv, 2 tiles per item
    dict(seed=30, B=1, H=128, W=128, C0=128, C1=64, Cout=64, fuse_res=True),
    dict(seed=31, B=3, H=64, W=64, C0=256, C1=128, Cout=128, fuse_res=True),    # 1 tile per item
    dict(seed=32, B=2, H=24, W=20, C0=64, C1=0, Cout=128, fuse_res=True),
], ids=lambda c: "C%d+%d_%d_%dx%dx%d%s" % (c["C0"], c["C1"], c["Cout"], c["B"], c["H"], c["W"], "_res" if c.get("fuse_res") else ""))
def test_tc_dense_halo(cfg):
    """csrc/ucdir_dhalo.cu: super tiles of 256/Cout 8x16-pixel tiles, one halo box per 64-channel chunk serves all nine taps."""
    cfg = dict(dict(ks=3, gn=True, act_=1, res=False, halo=1), **cfg)
    c, build = dense_case(**cfg)
    assert _lib.tc_schedule(build(c.on("cpu")).array()[0]) == 2
    host, dev = run_both(c, build)
    assert_close(dev["dst"], host["dst"], "dst")
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=2e-3, atol=2e-3)
    assert_close(dev["dst"], build.torch_ref(), "dst vs torch.nn.functional", rtol=3e-2, atol=3e-2)
    if cfg.get("fuse_res"):
        assert float(host["dres"].abs().max()) > 0.1
        assert_close(dev["dres"], host["dres"], "fused res_conv")


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 64), (1, 128, 128, 64), (3, 40, 72, 64), (2, 24, 36, 128)])
def test_tc_final_conv_fused_gn_swish(B, H, W, C):
    """csrc/ucdir_fhalo.cu: final_conv (GroupNorm -> Swish -> conv3x3 to 3 fp32 channels, model/ucdir.py:266-268) with the
    norm + activation applied to the landed halo box in shared memory; out-of-image pixels must stay zero padding."""
    g = torch.Generator().manual_seed(B * H + C)
    c = Case()
    c.add("x", (rnd(g, B, H, W, C) * 0.9 + 0.2).to(BF))
    w = rnd(g, 3, C, 3, 3, scale=1.0 / np.sqrt(C * 9))
    bias = rnd(g, 3, scale=0.1)
    wp, tb, _ = E.pack_tc_dense(w, bias, 16)
    c.add("w", wp).add("tb", tb).add("gamma", 1 + 0.3 * rnd(g, C)).add("beta", 0.2 * rnd(g, C)).add("s0", stats_of(c.t["x"]))
    c.add("dst", torch.zeros(B, H, W, 4))

    def build(t):
        ol = E.OpList()
        E._tc_op(ol, src0=act(t["x"], C, H, W, t["s0"]), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), dst=act(t["dst"], 4, H, W),
                 ntot=16, B=B, nt=16, dst_f32=1, ncol_valid=3, src_gn_swish=1, src_gamma=t["gamma"].data_ptr(),
                 src_beta=t["beta"].data_ptr())
        return ol
    assert _lib.tc_schedule(build(c.on("cpu")).array()[0]) == 3
    host, dev = run_both(c, build)
    assert_close(dev["dst"][..., :3], host["dst"][..., :3], "eps")
    assert float(dev["dst"][..., 3].abs().max()) == 0.0          # the pad column is not written


@pytest.mark.parametrize("halo", [0, 1], ids=["streamed", "halo"])
@pytest.mark.parametrize("C,B,H,W", [(64, 2, 16, 16), (128, 1, 24, 16), (256, 3, 8, 8), (512, 3, 8, 8), (64, 1, 128, 128),
                                     (128, 3, 64, 64), (256, 3, 32, 32), (64, 2, 36, 20), (256, 1, 18, 18)])
def test_tc_grouped_mix(C, B, H, W, halo):
    """halo=1: csrc/ucdir_mix.cu (one halo box per 8 x 16 pixel tile, resident weights) where it applies (C <= 256);
    the larger shapes make CTA unit ranges cross image and column-set boundaries."""
    if halo and C > 256:
        pytest.skip("halo schedule covers C = 64 / 128 / 256")
    g = torch.Generator().manual_seed(C + H)
    c = Case()
    c.add("h1", (torch.nn.functional.silu(rnd(g, B, H, W, C))).to(BF))
    w = rnd(g, 8 * C, C // 8, 3, 3, scale=1.0 / np.sqrt(C // 8 * 9))
    bias = rnd(g, 8 * C, scale=0.1)
    gamma, beta = 1 + 0.3 * rnd(g, C), 0.2 * rnd(g, C)
    kc, kb, nt, nsplit = E.tc_mix_tiling(C)
    wp, tb, tg = E.pack_tc_grouped(w, bias, 8, kb, gamma, beta)
    c.add("w", wp).add("tb", tb).add("tg", tg).add("s0", stats_of(c.t["h1"]))
    c.add("att", rnd(g, B, H, W, 8)).add("attw", rnd(g, B, 8)).add("res", rnd(g, B, H, W, C).to(BF))
    c.add("dst", torch.zeros(B, H, W, C, dtype=BF)).add("dstats", torch.zeros(B, 2, dtype=torch.float64))

    def build(t):
        ol = E.OpList()
        E._tc_op(ol, src0=act(t["h1"], C, H, W, t["s0"]), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), tg=t["tg"].data_ptr(),
                 gn=1, ncls=9, groups=8, kc=kc, kb=kb, nsplit=nsplit, nt=nt, mode=1, att=t["att"].data_ptr(), attw=t["attw"].data_ptr(),
                 attw_stride=8, res=act(t["res"], C, H, W), dst=act(t["dst"], C, H, W, t["dstats"]), ntot=8 * C, B=B, halo=halo)
        return ol
    assert _lib.tc_schedule(build(c.on("cpu")).array()[0]) == halo
    host, dev = run_both(c, build)
    assert_close(dev["dst"], host["dst"], "mix dst")
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=2e-3, atol=2e-3)
    # the integration module through torch.nn.functional in the reference's own form (model/ucdir.py:112,135-140): independent
    # of the op interpreter and of the weight packer
    F = torch.nn.functional
    h = F.group_norm(c.t["h1"].float().permute(0, 3, 1, 2), 1, gamma, beta, eps=1e-5)
    hset = F.conv2d(h, w, bias, padding=1, groups=8).view(B, C, 8, H, W)
    att_sp = c.t["att"].permute(0, 3, 1, 2) * c.t["attw"].view(B, 8, 1, 1)
    hh = torch.sum(hset * att_sp.unsqueeze(1), dim=2)
    want = (hh * torch.sigmoid(hh)).permute(0, 2, 3, 1) + c.t["res"].float()
    assert_close(dev["dst"], want, "mix dst vs torch.nn.functional", rtol=3e-2, atol=3e-2)


def split_planes(t32):
    """fp32 [..., C] -> bf16 [..., 2C] = [hi | lo] (UCDIR_TC_I_SPLIT activation layout)."""
    hi, lo = E.split_hi_lo(t32)
    return torch.cat([hi, lo], dim=-1).contiguous()


def join_planes(t, C):
    return t[..., :C].float() + t[..., C:2 * C].float()


@pytest.mark.parametrize("halo", [0, 1], ids=["streamed", "halo"])
@pytest.mark.parametrize("C,B,H,W", [(64, 2, 16, 16), (128, 1, 24, 16), (64, 1, 128, 128), (128, 3, 64, 64), (64, 2, 36, 20), (128, 2, 18, 18),
                                     (256, 3, 32, 32), (256, 1, 18, 18), (256, 3, 8, 8)])
def test_tc_grouped_mix_split(C, B, H, W, halo):
    """fp32_tc form of the integration-module conv: (hi, lo) plane pairs, three passes per tap, fp32 epilogue.  halo=1 is
    mix_halo_kernel<CG, SPLIT> (hi and lo halo boxes in a three-stage ring, both weight planes resident); both schedules are held
    to the fp32-class tolerance against torch.nn.functional in the reference's own form (model/ucdir.py:112,135-140)."""
    g = torch.Generator().manual_seed(7 * C + H)
    c = Case()
    h1 = torch.nn.functional.silu(rnd(g, B, H, W, C))
    c.add("h1", split_planes(h1))
    w = rnd(g, 8 * C, C // 8, 3, 3, scale=1.0 / np.sqrt(C // 8 * 9))
    bias = rnd(g, 8 * C, scale=0.1)
    gamma, beta = 1 + 0.3 * rnd(g, C), 0.2 * rnd(g, C)
    kc, kb, nt, nsplit = E.tc_mix_tiling(C)
    wp, tb, tg = E.pack_tc_grouped(w, bias, 8, kb, gamma, beta, split=True)
    h1v = join_planes(c.t["h1"], C)                       # the value the kernels see (16 mantissa bits)
    c.add("w", wp).add("tb", tb).add("tg", tg).add("s0", stats_of(h1v))
    res = rnd(g, B, H, W, C)
    c.add("att", rnd(g, B, H, W, 8)).add("attw", rnd(g, B, 8)).add("res", split_planes(res))
    c.add("dst", torch.zeros(B, H, W, 2 * C, dtype=BF)).add("dstats", torch.zeros(B, 2, dtype=torch.float64))

    def sact(t, stats=None):
        return E.Act(t, C, H, W, stats.data_ptr() if stats is not None else 0, True, True)

    def build(t):
        ol = E.OpList()
        E._tc_op(ol, split=1, src0=sact(t["h1"], t["s0"]), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), tg=t["tg"].data_ptr(),
                 gn=1, ncls=9, groups=8, kc=kc, kb=kb, nsplit=nsplit, nt=nt, mode=1, att=t["att"].data_ptr(), attw=t["attw"].data_ptr(),
                 attw_stride=8, res=sact(t["res"]), dst=sact(t["dst"], t["dstats"]), ntot=8 * C, B=B, halo=halo)
        return ol
    assert _lib.tc_schedule(build(c.on("cpu")).array()[0]) == halo
    host, dev = run_both(c, build)
    got = join_planes(dev["dst"], C)
    assert_close(got, join_planes(host["dst"], C), "split mix dst vs interpreter", rtol=1e-3, atol=1e-4)
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=1e-5, atol=1e-5)
    F = torch.nn.functional
    h = F.group_norm(h1v.permute(0, 3, 1, 2), 1, gamma, beta, eps=1e-5)
    hset = F.conv2d(h, w, bias, padding=1, groups=8).view(B, C, 8, H, W)
    att_sp = c.t["att"].permute(0, 3, 1, 2) * c.t["attw"].view(B, 8, 1, 1)
    hh = torch.sum(hset * att_sp.unsqueeze(1), dim=2)
    want = (hh * torch.sigmoid(hh)).permute(0, 2, 3, 1) + join_planes(c.t["res"], C)
    assert_close(got, want, "split mix dst vs torch.nn.functional", rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("cfg", [
    dict(seed=41, B=2, H=40, W=72, C0=16, C1=0, Cout=64, gn=False, act_=0, kc=16),          # the in-conv
    dict(seed=42, B=1, H=128, W=128, C0=16, C1=0, Cout=64, gn=False, act_=0, kc=16),
    dict(seed=43, B=2, H=40, W=72, C0=64, C1=64, Cout=64, fuse_res=True),                     # + fused 1x1 res_conv, concat input
    dict(seed=44, B=1, H=64, W=64, C0=128, C1=64, Cout=64, fuse_res=True),
    dict(seed=45, B=2, H=32, W=48, C0=256, C1=128, Cout=128, fuse_res=True),
    dict(seed=46, B=2, H=24, W=20, C0=64, C1=0, Cout=128, fuse_res=True),
], ids=lambda c: "C%d+%d_%d_%dx%dx%d%s" % (c["C0"], c["C1"], c["Cout"], c["B"], c["H"], c["W"], "_res" if c.get("fuse_res") else ""))
def test_tc_dense_halo_split(cfg):
    """fp32_tc form of csrc/ucdir_dhalo.cu: (hi, lo) plane pairs, per channel chunk a lo box (x W_hi) and a hi box (x W_hi, x W_lo);
    used for the 16-channel in-conv and for conv1 with the block's 1x1 res_conv riding along.  fp32-class tolerance against the op
    interpreter and against torch.nn.functional in the reference's form (model/ucdir.py:109-111,120)."""
    cfg = dict(dict(ks=3, gn=True, act_=1, kc=64, fuse_res=False), **cfg)
    g = torch.Generator().manual_seed(cfg["seed"])
    B, H, W, C0, C1, Cout, gn, act_, kc, fuse = (cfg[k] for k in ("B", "H", "W", "C0", "C1", "Cout", "gn", "act_", "kc", "fuse_res"))
    Cin = C0 + C1
    c = Case()
    x0 = rnd(g, B, H, W, C0) + 0.3
    c.add("x0", split_planes(x0))
    x0v = join_planes(c.t["x0"], C0)
    xv = x0v
    if C1:
        x1 = rnd(g, B, H, W, C1) * 0.7 - 0.2
        c.add("x1", split_planes(x1))
        xv = torch.cat([x0v, join_planes(c.t["x1"], C1)], dim=-1)
    w = rnd(g, Cout, Cin, 3, 3, scale=1.0 / np.sqrt(Cin * 9))
    bias = rnd(g, Cout, scale=0.1)
    gamma = 1 + 0.3 * rnd(g, Cin) if gn else None
    beta = 0.2 * rnd(g, Cin) if gn else None
    nt = E._tc_nt(Cout)
    wp, tb, tg = E.pack_tc_dense(w, bias, nt, gamma, beta, split=True, c0=C0)
    c.add("w", wp).add("tb", tb)
    if tg is not None:
        c.add("tg", tg)
    if gn:
        c.add("s0", stats_of(x0v))
        if C1:
            c.add("s1", stats_of(join_planes(c.t["x1"], C1)))
    c.add("dst", torch.zeros(B, H, W, 2 * Cout, dtype=BF)).add("dstats", torch.zeros(B, 2, dtype=torch.float64))
    if fuse:
        w2 = rnd(g, Cout, Cin, 1, 1, scale=1.0 / np.sqrt(Cin))
        b2 = rnd(g, Cout, scale=0.1)
        w2p, tb2, _ = E.pack_tc_dense(w2, b2, nt, split=True, c0=C0)
        c.add("w2", w2p).add("tb2", tb2).add("dres", torch.zeros(B, H, W, 2 * Cout, dtype=BF))

    def sact(t, C, stats=None):
        return E.Act(t, C, H, W, stats.data_ptr() if stats is not None else 0, True, True)

    def build(t):
        ol = E.OpList()
        extra = dict(w2=t["w2"].data_ptr(), tb2=t["tb2"].data_ptr(), dst_res=sact(t["dres"], Cout)) if fuse else {}
        E._tc_op(ol, **extra, split=1, src0=sact(t["x0"], C0, t.get("s0")), src1=sact(t["x1"], C1, t.get("s1")) if C1 else None,
                 w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), tg=t["tg"].data_ptr() if "tg" in t else 0, gn=1 if gn else 0,
                 ncls=9 if gn else 1, act=act_, dst=sact(t["dst"], Cout, t["dstats"]), ntot=Cout, B=B, nt=nt, halo=1, kc=kc)
        return ol
    assert _lib.tc_schedule(build(c.on("cpu")).array()[0]) == 2
    host, dev = run_both(c, build)
    got = join_planes(dev["dst"], Cout)
    assert_close(got, join_planes(host["dst"], Cout), "split dense dst vs interpreter", rtol=1e-3, atol=1e-4)
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=1e-4, atol=1e-4)      # per-thread fp32 partial sums, double across threads
    F = torch.nn.functional
    xn = xv.permute(0, 3, 1, 2)
    y = F.conv2d(F.group_norm(xn, 1, gamma, beta, eps=1e-5) if gn else xn, w, bias, padding=1)
    if act_ == 1:
        y = y * torch.sigmoid(y)
    assert_close(got, y.permute(0, 2, 3, 1), "split dense dst vs torch.nn.functional", rtol=1e-3, atol=1e-4)
    if fuse:
        want = F.conv2d(xn, w2, b2).permute(0, 2, 3, 1)
        assert_close(join_planes(dev["dres"], Cout), want, "fused res_conv vs torch.nn.functional", rtol=1e-3, atol=1e-4)
        assert_close(join_planes(dev["dres"], Cout), join_planes(host["dres"], Cout), "fused res_conv vs interpreter", rtol=1e-3, atol=1e-4)


def test_tc_upsample_phases_equal_upsample_then_conv():
    """Four 2x2-tap phase convolutions == nearest-2x + conv3x3 (model/ucdir.py:53-60), checked against torch."""
    g = torch.Generator().manual_seed(9)
    B, H, W, C = 2, 8, 12, 128
    x = rnd(g, B, H, W, C).to(BF)
    w = rnd(g, C, C, 3, 3, scale=1.0 / np.sqrt(9 * C))
    bias = rnd(g, C, scale=0.1)
    c = Case().add("x", x).add("dst", torch.zeros(B, 2 * H, 2 * W, C, dtype=BF)).add("dstats", torch.zeros(B, 2, dtype=torch.float64))
    for py in range(2):
        for px in range(2):
            wp, tb = E.pack_tc_up_phase(w, bias, py, px, 128)
            c.add("w%d%d" % (py, px), wp).add("tb%d%d" % (py, px), tb)

    def build(t):
        ol = E.OpList()
        for py in range(2):
            for px in range(2):
                E._tc_op(ol, src0=act(t["x"], C, H, W), w=t["w%d%d" % (py, px)].data_ptr(), tb=t["tb%d%d" % (py, px)].data_ptr(),
                         nty=2, ntx=2, oy0=py - 1, ox0=px - 1, dst=act(t["dst"], C, 2 * H, 2 * W, t["dstats"]), ntot=C, B=B,
                         nt=128, dst_up=1, dst_py=py, dst_px=px)
        return ol
    host, dev = run_both(c, build)
    assert_close(dev["dst"], host["dst"], "phase dst vs interpreter")
    xn = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    want = torch.nn.functional.conv2d(xn, w, bias, padding=1).permute(0, 2, 3, 1)
    assert_close(dev["dst"], want, "phase dst vs upsample+conv", rtol=3e-2, atol=3e-2)


def test_tc_final_conv_fp32_out_and_gn_apply():
    g = torch.Generator().manual_seed(10)
    B, H, W, C = 2, 16, 16, 64
    x = (rnd(g, B, H, W, C) * 1.3 + 0.4).to(BF)
    w = rnd(g, 3, C, 3, 3, scale=1.0 / np.sqrt(9 * C))
    bias = rnd(g, 3, scale=0.1)
    wp, tb, _ = E.pack_tc_dense(w, bias, 16)
    c = Case().add("x", x).add("s0", stats_of(x)).add("gamma", 1 + 0.3 * rnd(g, C)).add("beta", 0.2 * rnd(g, C))
    c.add("xn", torch.zeros(B, H, W, C, dtype=BF)).add("w", wp).add("tb", tb).add("eps", torch.zeros(B, H, W, 4))

    def build(t):
        ol = E.OpList()
        ol.add("UCDIR_OP_GN_APPLY_BF16", {"UCDIR_GNA_P_SRC": t["x"].data_ptr(), "UCDIR_GNA_P_DST": t["xn"].data_ptr(),
                                          "UCDIR_GNA_P_GAMMA": t["gamma"].data_ptr(), "UCDIR_GNA_P_BETA": t["beta"].data_ptr(),
                                          "UCDIR_GNA_P_STATS": t["s0"].data_ptr()},
               {"UCDIR_GNA_I_B": B, "UCDIR_GNA_I_HW": H * W, "UCDIR_GNA_I_C": C, "UCDIR_GNA_I_SWISH": 1}, {0: 1e-5})
        E._tc_op(ol, src0=act(t["xn"], C, H, W), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), dst=act(t["eps"], 4, H, W), ntot=16,
                 B=B, nt=16, dst_f32=1, ncol_valid=3)
        return ol
    host, dev = run_both(c, build)
    assert_close(dev["xn"], host["xn"], "gn_apply")
    assert_close(dev["eps"][..., :3], host["eps"][..., :3], "final conv eps")


def test_sgemm_bf16_operands():
    g = torch.Generator().manual_seed(12)
    Bt, N, C = 3, 80, 64
    qkv = rnd(g, Bt, N, 3 * C).to(BF)
    c = Case().add("qkv", qkv).add("S", torch.zeros(Bt, N, N)).add("O", torch.zeros(Bt, N, C, dtype=BF))

    def build(t):
        ol = E.OpList()
        q = t["qkv"].data_ptr()
        ol.add("UCDIR_OP_SGEMM_F32", {"UCDIR_SGEMM_P_A": q, "UCDIR_SGEMM_P_B": q + C * 2, "UCDIR_SGEMM_P_C": t["S"].data_ptr()},
               {"UCDIR_SGEMM_I_BATCH": Bt, "UCDIR_SGEMM_I_M": N, "UCDIR_SGEMM_I_N": N, "UCDIR_SGEMM_I_K": C, "UCDIR_SGEMM_I_LDA": 3 * C,
                "UCDIR_SGEMM_I_LDB": 3 * C, "UCDIR_SGEMM_I_LDC": N, "UCDIR_SGEMM_I_SA": N * 3 * C, "UCDIR_SGEMM_I_SB": N * 3 * C,
                "UCDIR_SGEMM_I_SC": N * N, "UCDIR_SGEMM_I_TRANSB": 1, "UCDIR_SGEMM_I_A_BF16": 1, "UCDIR_SGEMM_I_B_BF16": 1},
               {"UCDIR_SGEMM_F_ALPHA": 0.125})
        ol.add("UCDIR_OP_SOFTMAX_F32", {"UCDIR_SOFTMAX_P_X": t["S"].data_ptr()}, {"UCDIR_SOFTMAX_I_ROWS": Bt * N, "UCDIR_SOFTMAX_I_COLS": N})
        ol.add("UCDIR_OP_SGEMM_F32", {"UCDIR_SGEMM_P_A": t["S"].data_ptr(), "UCDIR_SGEMM_P_B": q + 2 * C * 2, "UCDIR_SGEMM_P_C": t["O"].data_ptr()},
               {"UCDIR_SGEMM_I_BATCH": Bt, "UCDIR_SGEMM_I_M": N, "UCDIR_SGEMM_I_N": C, "UCDIR_SGEMM_I_K": N, "UCDIR_SGEMM_I_LDA": N,
                "UCDIR_SGEMM_I_LDB": 3 * C, "UCDIR_SGEMM_I_LDC": C, "UCDIR_SGEMM_I_SA": N * N, "UCDIR_SGEMM_I_SB": N * 3 * C,
                "UCDIR_SGEMM_I_SC": N * C, "UCDIR_SGEMM_I_TRANSB": 0, "UCDIR_SGEMM_I_B_BF16": 1, "UCDIR_SGEMM_I_C_BF16": 1},
               {"UCDIR_SGEMM_F_ALPHA": 1.0})
        return ol
    host, dev = run_both(c, build)
    assert_close(dev["S"], host["S"], "softmax(QK^T)", rtol=1e-3, atol=1e-4)
    assert_close(dev["O"], host["O"], "PV")


@pytest.mark.parametrize("B,H,W,C", [(3, 16, 16, 512), (2, 8, 8, 512), (2, 18, 18, 128)])
def test_tc_attention_chain(B, H, W, C):
    """qkv conv (V^T written transposed) -> S = alpha Q K^T (batched weights) -> softmax -> O = P V, all on the
    tensor-core kernel, against the interpreter and against torch's attention on the same bf16 qkv."""
    import math
    g = torch.Generator().manual_seed(77 + H)
    N = H * W
    NP = (N + 7) & ~7
    nt_s = 256 if N % 256 == 0 else (128 if N % 128 == 0 else 64)
    NS = (N + nt_s - 1) // nt_s * nt_s
    x = rnd(g, B, H, W, C).to(BF)
    wq = rnd(g, 3 * C, C, 1, 1, scale=1.5 / np.sqrt(C))
    wp, tb, _ = E.pack_tc_dense(wq, None, 256 if (3 * C) % 256 == 0 else 128)
    c = Case().add("x", x).add("w", wp).add("tb", tb).add("zeros", torch.zeros(4096))
    c.add("qk", torch.zeros(B, H, W, 2 * C, dtype=BF)).add("vt", torch.zeros(B, C, NP, dtype=BF))
    c.add("S", torch.zeros(B, N, NS)).add("P", torch.zeros(B, N, NP, dtype=BF)).add("O", torch.zeros(B, H, W, C, dtype=BF))
    nt_qkv = 256 if (3 * C) % 256 == 0 else 128

    def build(t):
        ol = E.OpList()
        qk = act(t["qk"], 2 * C, H, W)
        E._tc_op(ol, src0=act(t["x"], C, H, W), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), nty=1, ntx=1, oy0=0, ox0=0, dst=qk,
                 ntot=3 * C, B=B, nt=nt_qkv, dst2=t["vt"].data_ptr(), t_col0=2 * C, t_ld=NP)
        E._tc_op(ol, src0=act(t["qk"], C, H, W), src_cstride=2 * C, w=t["qk"].data_ptr() + C * 2, w_batched=1, w_rowstride=2 * C,
                 w_batchstride=N * 2 * C, w_rows=N, tb=t["zeros"].data_ptr(), nty=1, ntx=1, oy0=0, ox0=0, dst=act(t["S"], NS, H, W),
                 ntot=NS, B=B, nt=nt_s, dst_f32=1, ncol_valid=NS, alpha=1.0 / math.sqrt(C))
        ol.add("UCDIR_OP_SOFTMAX_F32", {"UCDIR_SOFTMAX_P_X": t["S"].data_ptr(), "UCDIR_SOFTMAX_P_OUT_BF16": t["P"].data_ptr()},
               {"UCDIR_SOFTMAX_I_ROWS": B * N, "UCDIR_SOFTMAX_I_COLS": N, "UCDIR_SOFTMAX_I_IN_LD": NS, "UCDIR_SOFTMAX_I_OUT_LD": NP})
        E._tc_op(ol, src0=act(t["P"], N, H, W), src_cstride=NP, w=t["vt"].data_ptr(), w_batched=1, w_rowstride=NP,
                 w_batchstride=C * NP, tb=t["zeros"].data_ptr(), nty=1, ntx=1, oy0=0, ox0=0, dst=act(t["O"], C, H, W), ntot=C, B=B,
                 nt=E._tc_nt(C))
        return ol
    host, dev = run_both(c, build)
    assert_close(dev["qk"], host["qk"], "q,k")
    assert_close(dev["vt"][..., :N], host["vt"][..., :N], "V^T")
    # S is scratch after the softmax (the kernel keeps the exponentials in place), so P is what is compared
    assert_close(dev["P"], host["P"], "softmax")
    assert_close(dev["O"], host["O"], "O = P V")
    q, k = dev["qk"].float().reshape(B, N, 2 * C).split(C, dim=-1)
    v = dev["vt"].float()[..., :N].transpose(1, 2)
    want = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ v
    assert_close(dev["O"].reshape(B, N, C), want, "O vs torch attention", rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("rows,cols,in_ld,bf16_out", [(37, 4096, 4096, True), (5, 16384, 16384, True), (9, 2500, 2560, True), (7, 3000, 3000, False)])
def test_softmax_long_rows(rows, cols, in_ld, bf16_out):
    """One-pass shared-memory softmax used for rows >= 2048 columns (attention over large tiles)."""
    g = torch.Generator().manual_seed(rows + cols)
    out_ld = (cols + 7) & ~7
    c = Case().add("S", rnd(g, rows, in_ld, scale=3.0)).add("P", torch.zeros(rows, out_ld, dtype=BF))

    def build(t):
        ol = E.OpList()
        p = {"UCDIR_SOFTMAX_P_X": t["S"].data_ptr()}
        if bf16_out:
            p["UCDIR_SOFTMAX_P_OUT_BF16"] = t["P"].data_ptr()
        ol.add("UCDIR_OP_SOFTMAX_F32", p, {"UCDIR_SOFTMAX_I_ROWS": rows, "UCDIR_SOFTMAX_I_COLS": cols, "UCDIR_SOFTMAX_I_IN_LD": in_ld,
                                          "UCDIR_SOFTMAX_I_OUT_LD": out_ld})
        return ol
    host, dev = run_both(c, build)
    if bf16_out:
        assert_close(dev["P"], host["P"], "softmax bf16 out", rtol=1e-2, atol=1e-3)
    else:
        assert_close(dev["S"][:, :cols], host["S"][:, :cols], "softmax in place", rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("cfg", [
    dict(seed=31, B=2, H=128, W=128, C0=64, C1=0, Cout=64, ks=3, gn=True, act_=1, res=False, row3=1),
    dict(seed=32, B=1, H=128, W=128, C0=128, C1=64, Cout=64, ks=3, gn=True, act_=1, res=False, row3=1),
    dict(seed=33, B=1, H=6, W=256, C0=64, C1=0, Cout=128, ks=3, gn=False, act_=0, res=True, row3=1),
], ids=["64_64", "128+64_64", "W256_64_128"])
def test_tc_dense_row3(cfg):
    """ROW3 schedule: one 130-pixel activation row per filter row, the three horizontal taps read it through
    descriptors shifted by 0 / 1 / 2 rows (swizzle base offset).  Same results as the per-tap schedule."""
    c, build = dense_case(**cfg)
    host, dev = run_both(c, build)
    assert_close(dev["dst"], host["dst"], "dst")
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=2e-3, atol=2e-3)



# --------------------------------------------------------------------------------------------------------------------
# fused attention core (csrc/ucdir_attn.cu): UCDIR_OP_TC_ATTN against torch on the same bf16 operands
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N", [(3, 256), (2, 400), (1, 1296), (2, 64), (1, 16), (1, 2176)])
def test_tc_flash_attention(B, N):
    """N = 256 / 400 / 1296: the token counts of 128 / 160 / 288-pixel samples; 64 and 16: less than one key block /
    one MMA K step; 2176 = 17 key blocks (lazy-rescale path across many blocks, both score buffers, ring wrap-around).
    Scores are scaled up so that the softmax is peaked (row maxima move between blocks) for half of the cases."""
    C = 512
    g = torch.Generator().manual_seed(N * 7 + B)
    NP = (N + 7) & ~7
    for peak in (1.0, 6.0):
        qk = (rnd(g, B, N, 2 * C) * peak).to(BF)
        v = rnd(g, B, N, C).to(BF)
        vt = torch.zeros(B, C, NP, dtype=BF)
        vt[:, :, :N] = v.transpose(1, 2)
        vt[:, :, N:] = float("nan")                                  # the pitch padding must never be read
        c = Case().add("qk", qk).add("vt", vt).add("o", torch.zeros(B, N, C, dtype=BF))

        def build(t):
            ol = E.OpList()
            ol.add("UCDIR_OP_TC_ATTN", {"UCDIR_ATTN_P_QK": t["qk"].data_ptr(), "UCDIR_ATTN_P_VT": t["vt"].data_ptr(), "UCDIR_ATTN_P_O": t["o"].data_ptr()},
                   {"UCDIR_ATTN_I_B": B, "UCDIR_ATTN_I_N": N, "UCDIR_ATTN_I_C": C, "UCDIR_ATTN_I_QK_LD": 2 * C, "UCDIR_ATTN_I_VT_LD": NP,
                    "UCDIR_ATTN_I_O_LD": C}, {"UCDIR_ATTN_F_SCALE": 1.0 / np.sqrt(C)})
            return ol

        ol = build(c.on("cuda"))                                      # record validation (no launch)
        _lib.check_ops(ol.array(), len(ol))
        devt = c.on("cuda")
        ol = build(devt)
        _lib.run_ops(ol.array(), len(ol), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got = devt["o"].cpu().float()
        q, k = qk[..., :C].float(), qk[..., C:].float()
        att = torch.softmax(torch.bmm(q, k.transpose(1, 2)) / np.sqrt(C), dim=-1)
        want = torch.bmm(att, v.float())
        assert torch.isfinite(got).all(), "non-finite output (N=%d peak=%g)" % (N, peak)
        assert_close(got, want, "flash attention B=%d N=%d peak=%g" % (B, N, peak), rtol=2e-2, atol=1e-2)


def test_tc_flash_attention_16k_tokens_vs_materialised_path():
    """N = 16 384 (a 1024x1024 tile at 1/8 resolution -- what `sr.py -p val` runs): the fused kernel against the three-launch
    form (score GEMM to fp32 in HBM, softmax, PV GEMM) of the same library on identical operands."""
    C, N, B = 512, 16384, 1
    g = torch.Generator().manual_seed(4)
    x = (rnd(g, B, 128, 128, C) * 0.5).to(BF).cuda()
    wq = (rnd(g, 3 * C, C) * (1.0 / np.sqrt(C))).to(BF).cuda()
    tb = torch.zeros(3 * C).cuda()
    qk = torch.empty(B, N, 2 * C, dtype=BF, device="cuda")
    vt = torch.empty(B, C, N, dtype=BF, device="cuda")
    o1 = torch.empty(B, N, C, dtype=BF, device="cuda"); o2 = torch.empty_like(o1)
    S = torch.empty(B, N, N, dtype=torch.float32, device="cuda"); P = torch.empty(B, N, N, dtype=BF, device="cuda")
    zeros = torch.zeros(N, device="cuda")
    ol = E.OpList()
    xa = act(x, C, 128, 128)
    E._tc_op(ol, src0=xa, w=wq.data_ptr(), tb=tb.data_ptr(), nty=1, ntx=1, oy0=0, ox0=0, dst=act(qk, 2 * C, 128, 128), ntot=3 * C, B=B, nt=256,
             dst2=vt.data_ptr(), t_col0=2 * C, t_ld=N)
    ol.add("UCDIR_OP_TC_ATTN", {"UCDIR_ATTN_P_QK": qk.data_ptr(), "UCDIR_ATTN_P_VT": vt.data_ptr(), "UCDIR_ATTN_P_O": o1.data_ptr()},
           {"UCDIR_ATTN_I_B": B, "UCDIR_ATTN_I_N": N, "UCDIR_ATTN_I_C": C, "UCDIR_ATTN_I_QK_LD": 2 * C, "UCDIR_ATTN_I_VT_LD": N, "UCDIR_ATTN_I_O_LD": C})
    E._tc_op(ol, src0=act(qk, C, 128, 128), src_cstride=2 * C, w=qk.data_ptr() + C * 2, w_batched=1, w_rowstride=2 * C, w_batchstride=N * 2 * C,
             w_rows=N, tb=zeros.data_ptr(), nty=1, ntx=1, oy0=0, ox0=0, dst=act(S, N, 128, 128), ntot=N, B=B, nt=256, dst_f32=1, ncol_valid=N,
             alpha=1.0 / np.sqrt(C))
    ol.add("UCDIR_OP_SOFTMAX_F32", {"UCDIR_SOFTMAX_P_X": S.data_ptr(), "UCDIR_SOFTMAX_P_OUT_BF16": P.data_ptr()},
           {"UCDIR_SOFTMAX_I_ROWS": B * N, "UCDIR_SOFTMAX_I_COLS": N, "UCDIR_SOFTMAX_I_IN_LD": N, "UCDIR_SOFTMAX_I_OUT_LD": N})
    E._tc_op(ol, src0=act(P, N, 128, 128), src_cstride=N, w=vt.data_ptr(), w_batched=1, w_rowstride=N, w_batchstride=C * N, tb=zeros.data_ptr(),
             nty=1, ntx=1, oy0=0, ox0=0, dst=act(o2, C, 128, 128), ntot=C, B=B, nt=256)
    _lib.run_ops(ol.array(), len(ol), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    a, b = o1.float().cpu(), o2.float().cpu()
    assert torch.isfinite(a).all()
    assert_close(a, b, "flash vs materialised attention, 16384 tokens", rtol=2e-2, atol=1e-2)


@pytest.mark.parametrize("B,H,W,C", [(2, 8, 12, 128), (5, 8, 8, 512), (3, 32, 32, 256)])
def test_tc_upsample_fused_phases(B, H, W, C):
    """UCDIR_TC_I_PHASES = 4: the four phase convolutions of nearest-2x + conv3x3 in ONE launch (phase = extra work-item dimension,
    weight blocks stacked along N) against the interpreter and against torch's upsample + conv on the same bf16 input; statistics of
    the output accumulate over all four phases."""
    g = torch.Generator().manual_seed(C + H)
    x = rnd(g, B, H, W, C).to(BF)
    w = rnd(g, C, C, 3, 3, scale=1.0 / np.sqrt(9 * C))
    bias = rnd(g, C, scale=0.1)
    nt = E._tc_nt(C)
    blocks = []
    for py in range(2):
        for px in range(2):
            wp, tb = E.pack_tc_up_phase(w, bias, py, px, nt)
            blocks.append(wp)
    c = Case().add("x", x).add("dst", torch.zeros(B, 2 * H, 2 * W, C, dtype=BF)).add("dstats", torch.zeros(B, 2, dtype=torch.float64))
    c.add("w", torch.cat(blocks, 0)).add("tb", tb)

    def build(t):
        ol = E.OpList()
        E._tc_op(ol, src0=act(t["x"], C, H, W), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), nty=2, ntx=2, oy0=-1, ox0=-1,
                 dst=act(t["dst"], C, 2 * H, 2 * W, t["dstats"]), ntot=C, B=B, nt=nt, dst_up=1, phases=4)
        return ol
    _lib.check_ops(build(c.on("cpu")).array(), 1)
    host, dev = run_both(c, build)
    assert_close(dev["dst"], host["dst"], "fused phases vs interpreter")
    assert_close(dev["dstats"], host["dstats"], "stats", rtol=2e-3, atol=2e-3)
    xn = torch.nn.functional.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    want = torch.nn.functional.conv2d(xn, w, bias, padding=1).permute(0, 2, 3, 1)
    assert_close(dev["dst"], want, "fused phases vs upsample+conv", rtol=3e-2, atol=3e-2)
