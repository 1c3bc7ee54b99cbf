"""End-to-end parity of the bf16 / tcgen05 path against the same golden vectors and oracle as the fp32 path.

Stated bf16 tolerance (BASELINE.json north_star "stated bf16 tolerance"; SURVEY 8d measured 7e-3 max / 1.2e-3
mean on eps for bf16 operand rounding alone, this path also stores activations in bf16):
    per UNet evaluation (eps, |eps| <= ~1.7):   max abs err <= 2e-2,  mean abs err <= 2.5e-3
    per denoising step and final image in [-1,1]: max abs err <= 2e-2, mean abs err <= 2e-3
(about 2x the measured values of profiles/r01_bf16_errors.json: eps 8.6e-3 / 1.45e-3, images 9.0e-3 / 5.4e-4 -- a kernel
regression that doubles the error fails.)
The measured errors are written to gpurun_out/bf16_errors.json (copied to profiles/ for DESIGN.md)."""
import json
import os

import numpy as np
import pytest
import torch

import ucdir_b200
from oracle import ucdir_oracle as O

pytestmark = pytest.mark.gpu
T = lambda a: torch.from_numpy(np.asarray(a))
EPS_MAX, EPS_MEAN = 2e-2, 2.5e-3
IMG_MAX, IMG_MEAN = 2e-2, 2e-3
REPORT = {}


def check(got, want, what, mx, mn):
    got, want = torch.as_tensor(got).float().cpu(), torch.as_tensor(want).float().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    err = (got - want).abs()
    REPORT[what] = {"max_abs_err": err.max().item(), "mean_abs_err": err.mean().item(), "ref_absmax": want.abs().max().item()}
    assert torch.isfinite(got).all(), what + ": non-finite output"
    assert err.max().item() <= mx and err.mean().item() <= mn, "%s: max %.3e mean %.3e (limits %.1e / %.1e)" % (
        what, err.max().item(), err.mean().item(), mx, mn)


@pytest.fixture(scope="module")
def net(sid_weights):
    from ucdir_b200 import _lib
    _lib.load()
    n, _ = sid_weights
    n = n.to("cuda")
    n.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, torch.device("cuda"))
    n.denoise_fn.engine().set_precision("bf16")
    yield n
    n.denoise_fn.engine().set_precision("fp32")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(REPORT, open("gpurun_out/bf16_errors.json", "w"), indent=1)


def test_unet_golden(net, golden):
    g = golden("unet")
    eps = net.denoise_fn(T(g["x6"]).cuda(), T(g["level"]).cuda(), T(g["guide"]).cuda())
    check(eps, g["eps"], "eps DY3h.forward 64->96", EPS_MAX, EPS_MEAN)
    eps2 = net.denoise_fn.naiveforward(T(g["xs"]).cuda(), T(g["lv2"]).cuda(), T(g["gs"]).cuda())
    check(eps2, g["eps2"], "eps naiveforward B=2", EPS_MAX, EPS_MEAN)


def test_tiler_golden(net, golden, monkeypatch):
    g = golden("tiler")
    skip, padding = (int(v) for v in g["geom"])
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", skip); monkeypatch.setattr(unet, "tile_padding", padding)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    out = unet(T(g["x"]).cuda(), T(g["level"]).cuda(), T(g["guide"]).cuda())
    check(out, g["out"], "eps tiler (64,16) 96x80", EPS_MAX, EPS_MEAN)


def test_super_resolution_golden_e2e(net, golden):
    g = golden("sr_e2e")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cuda"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    try:
        out = net.super_resolution(T(g["x_in"]).cuda(), True)
    finally:
        net._noise_source = None
        net.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, torch.device("cuda"))
    check(out, g["out"], "super_resolution T=4 (all snapshots)", IMG_MAX, IMG_MEAN)


def test_50_step_sampler_vs_oracle_teacher_forced_every_10th(net, sid_weights):
    """sid val schedule (T=50) on a 64x64 image: full 50-step run on the GPU; the CPU oracle replays steps
    t = 49, 40, 30, 20, 10, 0 from the GPU's own x_t (teacher forcing), so each compared step isolates one
    UNet evaluation + posterior update.  Also checks the end-to-end run stays finite and in range."""
    _, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    sched = O.schedule_buffers(ucdir_b200.SID_VAL_SCHEDULE)
    g = torch.Generator().manual_seed(21)
    x_in = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        initx = O.predictor_forward(sd, "predictor.", x_in)
    den = lambda xc, lvl, gd: O.unet_forward(sd, "denoise_fn.", lay, xc, lvl, gd)
    x = torch.randn(1, 3, 64, 64, generator=g)
    worst = 0.0
    try:
        for t in reversed(range(50)):
            z = torch.randn(1, 3, 64, 64, generator=g)
            net._noise_source = lambda shape: z
            got = net.p_sample(x.cuda(), t, condition_x=x_in.cuda(), kwargs={"guide": initx.cuda()}).cpu()
            if t % 10 == 0 or t == 49:
                with torch.no_grad():
                    want = O.p_sample(sched, den, x, t, x_in, initx, z)
                check(got, want, "p_sample t=%d" % t, IMG_MAX, IMG_MEAN)
            assert torch.isfinite(got).all()
            x = got
    finally:
        net._noise_source = None
    assert x.abs().max().item() <= 1.0 + 1e-3


def test_full_size_tile_locality_vs_oracle(net, sid_weights, monkeypatch):
    _, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", 128); monkeypatch.setattr(unet, "tile_padding", 16)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    g = torch.Generator().manual_seed(3)
    low = torch.nn.functional.interpolate(torch.rand(1, 3, 64, 64, generator=g), size=(1024, 1024), mode="bilinear")
    x_in = (low * 0.1 + 0.05 * torch.randn(1, 3, 1024, 1024, generator=g)).clamp(0, 1) * 2 - 1
    guide = (low * 2 - 1).contiguous()
    x_t = torch.randn(1, 3, 1024, 1024, generator=g)
    lvl = torch.full((1, 1), 0.5)
    eps = unet(torch.cat([x_in, x_t], 1).cuda(), lvl.cuda(), guide.cuda()).cpu()
    geo = unet.engine().default_geometry(1, 1024, 1024)
    xp = torch.nn.functional.pad(torch.cat([x_in, x_t], 1), (16,) * 4, mode="reflect")
    gp = torch.nn.functional.pad(guide, (16,) * 4, mode="reflect")
    for (ty, tx) in [(0, 0), (5, 7), (10, 10)]:
        y0, x0 = geo.ys[ty], geo.xs[tx]
        with torch.no_grad():
            want = O.unet_naiveforward(sd, "denoise_fn.", lay, xp[..., y0:y0 + 128, x0:x0 + 128], lvl, gp[..., y0:y0 + 128, x0:x0 + 128])
        check(eps[..., y0:y0 + 96, x0:x0 + 96], want[..., 16:112, 16:112], "eps 1024 tile (%d,%d)" % (ty, tx), EPS_MAX, EPS_MEAN)


@pytest.mark.parametrize("shape", [(2, 33, 47), (1, 48, 80), (2, 256, 256)])
def test_ragged_and_batched_shapes_vs_oracle(net, sid_weights, shape):
    """Non-aligned, non-square and batched inputs (incl. BASELINE config C2's 256x256 -> S = 288, whose levels are
    288/144/72/36/18 pixels wide: tile rectangles that do not divide 128) on the tensor-core path."""
    _, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    b, h, w = shape
    g = torch.Generator().manual_seed(h * 1000 + w)
    x6 = torch.rand(b, 6, h, w, generator=g) * 2 - 1
    guide = torch.rand(b, 3, h, w, generator=g) * 2 - 1
    lvl = torch.rand(b, 1, generator=g)
    with torch.no_grad():
        want = O.unet_forward(sd, "denoise_fn.", lay, x6, lvl, guide)
    check(net.denoise_fn(x6.cuda(), lvl.cuda(), guide.cuda()), want, "eps forward %s" % (shape,), EPS_MAX, EPS_MEAN)


def test_reference_default_tiling_1152_bf16_vs_fp32_path(net):
    """What `sr.py -p val` produces for a 1024x1024 image: DDPM.test pads it to 1152x1152 (model/model.py:127-128), which
    is above the 1 Mpx trigger, so DY3h.forward tiles it with the reference defaults (skip 1024, padding 64) into four
    1024x1024 tiles, each with 16384-token attention.  The CPU oracle needs minutes per tile at this size, so the check
    is between the two independent CUDA paths: the bf16 tensor-core path against the fp32 SIMT parity path."""
    unet = net.denoise_fn
    eng = unet.engine()
    assert unet.tile_skip == 1024 and unet.tile_padding == 64
    g = torch.Generator().manual_seed(8)
    low = torch.nn.functional.interpolate(torch.rand(1, 6, 36, 36, generator=g), size=(1152, 1152), mode="bilinear")
    x6 = (low * 2 - 1 + 0.1 * torch.randn(1, 6, 1152, 1152, generator=g)).clamp(-1, 1).cuda()
    guide = (low[:, :3] * 2 - 1).contiguous().cuda()
    lvl = torch.full((1, 1), 0.6).cuda()
    geo = eng.default_geometry(1, 1152, 1152)
    assert geo.kind == "tiled" and geo.n_tiles == 4 and geo.TH == 1024
    a = unet(x6, lvl, guide)
    b = unet(x6, lvl, guide)
    check(a, b, "1152 default tiling: idempotence", 1e-3, 1e-4)
    eng.set_precision("fp32")
    try:
        ref = unet(x6, lvl, guide)
    finally:
        eng.set_precision("bf16")
    check(a, ref, "1152 default tiling: bf16 vs fp32 path", EPS_MAX, EPS_MEAN)
