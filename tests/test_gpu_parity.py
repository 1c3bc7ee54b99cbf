"""Parity tests proper (run on the B200 with `-m gpu`): the CUDA path, called through the reference-facing
module API (define_G / super_resolution / DY3h.forward -> C ABI -> kernels), against
  (a) the committed golden vectors produced by the reference itself (tests/golden/make_golden.py),
  (b) the CPU oracle (oracle/ucdir_oracle.py) on seeded inputs at sizes it finishes in seconds,
  (c) size-independent properties at BASELINE.json's full size (1024x1024, 128-px tiles).
fp32 tolerance (BASELINE.json north_star): rtol 1e-3 / atol 1e-4, met by BOTH fp32-class precisions: the split-operand
tensor-core mode "fp32_tc" and the SIMT mode "fp32" (every `net` test runs once per mode).  Nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

import ucdir_b200
from oracle import ucdir_oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
T = lambda a: torch.from_numpy(np.asarray(a))


REPORT, MODE = {}, ["-"]


def close(a, b, rtol=RTOL, atol=ATOL, what=""):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    bad = (err > atol + rtol * b.abs()).float().mean().item()
    REPORT.setdefault(MODE[0], {})[what] = {"max_abs_err": err.max().item(), "mean_abs_err": err.mean().item(),
                                            "ref_absmax": b.abs().max().item(), "frac_outside": bad}
    assert bad == 0.0, f"{what}: max abs err {err.max().item():.3e}, {bad:.2%} of elements outside rtol={rtol} atol={atol}"


@pytest.fixture(scope="module", params=["fp32_tc", "fp32"])
def net(sid_weights, request):
    """Both paths that claim the reference's fp32 tolerance run every test below: "fp32_tc" = tcgen05 with split (hi + lo
    bf16) operands, three MMA passes per filter tap, fp32 accumulation and epilogue (the product's parity mode), and "fp32" =
    the SIMT FFMA kernels (debug path, independent arithmetic)."""
    from ucdir_b200 import _lib
    _lib.load()                                   # raises if the .so is missing or the device is not sm_100
    n, _ = sid_weights
    n = n.to("cuda")
    n.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, torch.device("cuda"))
    n.denoise_fn.engine().set_precision(request.param)
    MODE[0] = request.param
    yield n
    n.denoise_fn.engine().set_precision("fp32")
    MODE[0] = "-"
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(REPORT, open("gpurun_out/fp32_errors.json", "w"), indent=1)      # measured errors per mode, copied to profiles/


@pytest.fixture(scope="module")
def sd(sid_weights):
    return sid_weights[1]


@pytest.fixture(scope="module")
def layout():
    return O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])


def test_library_is_native_and_counts_launches(net):
    from ucdir_b200 import _lib
    before = _lib.launch_count()
    net.predictor(torch.zeros(1, 3, 48, 48, device="cuda"))
    torch.cuda.synchronize()
    assert _lib.launch_count() - before >= 30


def test_unet_golden_direct_and_naive(net, golden):
    g = golden("unet")
    eps = net.denoise_fn(T(g["x6"]).cuda(), T(g["level"]).cuda(), T(g["guide"]).cuda())
    close(eps, g["eps"], what="DY3h.forward 64->96")
    eps2 = net.denoise_fn.naiveforward(T(g["xs"]).cuda(), T(g["lv2"]).cuda(), T(g["gs"]).cuda())
    close(eps2, g["eps2"], what="naiveforward B=2 per-sample levels")


def test_predictor_golden(net, golden):
    g = golden("unet")
    close(net.predictor(T(g["xp"]).cuda()), g["pred"], what="UNetSeeInDark 40x56")


def test_tiler_golden(net, golden, monkeypatch):
    g = golden("tiler")
    skip, padding = (int(v) for v in g["geom"])
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", skip); monkeypatch.setattr(unet, "tile_padding", padding)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    out = unet(T(g["x"]).cuda(), T(g["level"]).cuda(), T(g["guide"]).cuda())
    close(out, g["out"], what="patch_forward_guide (64,16) on 96x80")


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
    close(net.pre_initx, g["initx"], what="initx")
    close(out, g["out"], what="super_resolution T=4 continous")


def test_degraded_guidance_wrapper_golden(golden):
    """ResiGaussianGuideDY_de.super_resolution (model/diffusion.py:518-523; the degraded input guides) on the CUDA path."""
    from ucdir_b200.model.networks import define_G
    torch.manual_seed(1234)
    net = define_G({"model": dict(ucdir_b200.SID_MODEL_OPT, diffusion_name="ResiGaussianGuideDY_de")}).cuda().eval()
    g = golden("sr_de")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cuda"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    out = net.super_resolution(T(g["x_in"]).cuda(), False)
    close(net.pre_initx, g["initx"], what="initx")
    close(out, g["out"], what="ResiGaussianGuideDY_de.super_resolution T=4")


def test_teacher_forced_steps_vs_oracle_c1(net, sd, layout):
    """BASELINE config C1 (1x3x128x128, T=4): every step is fed the oracle's x_t so that errors cannot
    compound; eps-equivalent check via x_{t-1} of the same inputs."""
    sched_opt = dict(schedule="linear", n_timestep=4, linear_start=1e-6, linear_end=0.4)
    sched = O.schedule_buffers(sched_opt)
    net.set_new_noise_schedule(sched_opt, torch.device("cuda"))
    g = torch.Generator().manual_seed(5)
    x_in = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        initx = O.predictor_forward(sd, "predictor.", x_in)
    close(net.predictor(x_in.cuda()), initx, what="predictor 128")
    den = lambda xc, lvl, gd: O.unet_forward(sd, "denoise_fn.", layout, xc, lvl, gd)
    x = torch.randn(1, 3, 128, 128, generator=g)
    try:
        for t in reversed(range(4)):
            z = torch.randn(1, 3, 128, 128, generator=g)
            with torch.no_grad():
                want = O.p_sample(sched, den, x, t, x_in, initx, z)
            net._noise_source = lambda shape: z
            got = net.p_sample(x.cuda(), t, condition_x=x_in.cuda(), kwargs={"guide": initx.cuda()})
            close(got, want, what=f"p_sample t={t}")
            x = want
    finally:
        net._noise_source = None
        net.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, torch.device("cuda"))


@pytest.mark.parametrize("shape", [(1, 48, 80), (2, 33, 47), (1, 96, 96)])
def test_ragged_shapes_vs_oracle(net, sd, layout, shape):
    """Non-square, non-aligned and batch>1 inputs through DY3h.forward's pad-to-(h//32+1)*32 rule."""
    b, h, w = shape
    g = torch.Generator().manual_seed(h * 1000 + w)
    x6 = torch.rand(b, 6, h, w, generator=g) * 2 - 1
    guide = torch.rand(b, 3, h, w, generator=g) * 2 - 1
    lvl = torch.rand(b, 1, generator=g)
    with torch.no_grad():
        want = O.unet_forward(sd, "denoise_fn.", layout, x6, lvl, guide)
    close(net.denoise_fn(x6.cuda(), lvl.cuda(), guide.cuda()), want, what=f"forward {shape}")


def test_small_image_under_tile_size(net, sd, layout, monkeypatch):
    """utils/util.py:114-115: an image smaller than the tile gets pd = skip - min(h,w) + padding."""
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", 64); monkeypatch.setattr(unet, "tile_padding", 8)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    g = torch.Generator().manual_seed(11)
    x6 = torch.rand(1, 6, 48, 56, generator=g) * 2 - 1
    guide = torch.rand(1, 3, 48, 56, generator=g) * 2 - 1
    lvl = torch.full((1, 1), 0.7)
    with torch.no_grad():
        want = O.unet_forward(sd, "denoise_fn.", layout, x6, lvl, guide, skip=64, padding=8, force_tiler=True)
    close(unet(x6.cuda(), lvl.cuda(), guide.cuda()), want, what="tiler on 48x56 with 64-px tiles")


def test_reference_error_behaviour(net):
    """Same failures as the reference: reflect pad >= dim (32x32 direct) raises; CPU tensors raise."""
    with pytest.raises(Exception):
        net.denoise_fn(torch.zeros(1, 6, 32, 32, device="cuda"), torch.zeros(1, 1, device="cuda"),
                       torch.zeros(1, 3, 32, 32, device="cuda"))


def test_full_size_properties_1024_tiled(net, sd, layout, monkeypatch):
    """BASELINE config C3: 1x3x1024x1024, tiler (128,16) -> 121 tiles.  Properties that do not need the
    oracle at full size: (1) idempotence (same inputs -> same output up to fp64-atomic order),
    (2) locality: a pixel's eps depends only on its owning tile -> three tiles recomputed alone through
    naiveforward on the CPU oracle match the stitched result, (3) the posterior step is linear in the
    injected noise: x(z) - x(0) == sigma_t * z."""
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", 128); monkeypatch.setattr(unet, "tile_padding", 16)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    g = torch.Generator().manual_seed(3)
    low = torch.nn.functional.interpolate(torch.rand(1, 3, 64, 64, generator=g), size=(1024, 1024), mode="bilinear")
    x_in = (low * 0.1 + 0.05 * torch.randn(1, 3, 1024, 1024, generator=g)).clamp(0, 1) * 2 - 1
    guide = (low * 2 - 1).contiguous()
    x_t = torch.randn(1, 3, 1024, 1024, generator=g)
    lvl = torch.full((1, 1), 0.5)
    x6 = torch.cat([x_in, x_t], 1).cuda()
    eps1 = unet(x6, lvl.cuda(), guide.cuda())
    eps2 = unet(x6, lvl.cuda(), guide.cuda())
    assert eps1.shape == (1, 3, 1024, 1024)
    close(eps1, eps2, rtol=1e-5, atol=1e-6, what="idempotence")
    geo = unet.engine().default_geometry(1, 1024, 1024)
    assert geo.n_tiles == 121 and geo.PD == 16
    xp = torch.nn.functional.pad(torch.cat([x_in, x_t], 1), (16,) * 4, mode="reflect")
    gp = torch.nn.functional.pad(guide, (16,) * 4, mode="reflect")
    e = eps1.cpu()
    for (ty, tx) in [(0, 0), (5, 7), (10, 10)]:
        y0, x0 = geo.ys[ty], geo.xs[tx]
        with torch.no_grad():
            want = O.unet_naiveforward(sd, "denoise_fn.", layout, xp[..., y0:y0 + 128, x0:x0 + 128], lvl,
                                       gp[..., y0:y0 + 128, x0:x0 + 128])
        ys, xs = y0 + 16 - 16, x0 + 16 - 16                   # interior of the window in unpadded coordinates
        close(e[..., ys:ys + 96, xs:xs + 96], want[..., 16:112, 16:112], what=f"tile ({ty},{tx}) vs oracle")
    # posterior linearity in z at t = 10
    z = torch.randn(1, 3, 1024, 1024, generator=g).cuda()
    xt = x_t.cuda()
    net._noise_source = lambda shape: z
    a = net.p_sample(xt, 10, condition_x=x_in.cuda(), kwargs={"guide": guide.cuda()})
    net._noise_source = lambda shape: torch.zeros_like(z)
    b = net.p_sample(xt, 10, condition_x=x_in.cuda(), kwargs={"guide": guide.cuda()})
    net._noise_source = None
    sigma = float(np.exp(np.float32(0.5) * net._sched_host["posterior_log_variance_clipped"][10]))
    close(a - b, sigma * z, rtol=1e-4, atol=2e-5, what="posterior linear in z")


def test_film_resnet_block_golden(golden):
    """SURVEY 8 a14 on the GPU: FiLM ResnetBlock (8-group GroupNorm, additive and affine FiLM) vs the reference's output."""
    from ucdir_b200.model.ucdir import ResnetBlock
    g = golden("modules")
    for tag, aff in (("film", False), ("filmaff", True)):
        m = ResnetBlock(16, 32, nl_emb_dim=64, use_affine_level=aff, norm_groups=8)
        m.load_state_dict({k[len(tag) + 3:]: T(g[k]) for k in g.files if k.startswith(tag + ".w.")}, strict=True)
        y = m.cuda()(T(g[tag + ".x"]).cuda(), T(g[tag + ".t"]).cuda())
        close(y, g[tag + ".y"], what="FiLM ResnetBlock " + tag)


def test_ddim_sample_golden(net, golden):
    """ddim_sample (5 strided steps, eta=1) vs the reference's trajectory with injected noise."""
    g = golden("ddim")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cuda"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    try:
        traj = net.ddim_sample(T(g["x_in"]).cuda(), True, kwargs={"guide": T(g["initx"]).cuda()})
    finally:
        net._noise_source = None
        net.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, torch.device("cuda"))
    close(traj, g["traj"], what="ddim trajectory")


def test_tensor2img_kernel_bit_exact(golden):
    """UCDIR_OP_TO_IMAGE_U8 (crop + core/metrics.py:8-34 on the device) against the reference's own uint8 output."""
    from ucdir_b200.utils.image import tensor2img
    g = golden("image")
    x = T(g["x"]).cuda()
    assert np.array_equal(tensor2img(x), g["img"])
    assert np.array_equal(tensor2img(x, crop=int(g["crop"])), g["img_crop"])
    with pytest.raises(Exception):
        tensor2img(T(g["x"]))                                  # CPU tensor: no fallback


@pytest.mark.parametrize("shape", [(1, 40, 56), (2, 72, 88), (1, 130, 250)])
def test_predictor_tensor_core_plan(sid_weights, sd, golden, shape):
    """UNetSeeInDark on the split-operand tensor-core plan (what the samplers use whenever the denoiser runs on the tensor cores):
    the reference's golden vector and ragged / batched shapes against the oracle, fp32 tolerance."""
    net, _ = sid_weights
    net = net.to("cuda")
    eng = net.predictor.engine()
    eng.set_mode("tc")
    try:
        if shape == (1, 40, 56):
            g = golden("unet")
            close(net.predictor(T(g["xp"]).cuda()), g["pred"], what="predictor tc plan, golden 40x56")
        b, h, w = shape
        x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(h + w)) * 2 - 1
        with torch.no_grad():
            want = O.predictor_forward(sd, "predictor.", x)
        close(net.predictor(x.cuda()), want, what="predictor tc plan %s" % (shape,))
    finally:
        eng.set_mode("fp32")
