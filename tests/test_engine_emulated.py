"""Host logic on CPU: the op arrays the engine emits (weight packing, tile tables, buffer reuse, op wiring) are
executed by the CPU interpreter in tests/op_emulator.py and compared with the reference's golden vectors.
The CUDA kernels themselves are covered by the `-m gpu` tests; this file catches graph / plumbing errors in
the build container, where there is no GPU."""
import numpy as np
import pytest
import torch

import ucdir_b200
from ucdir_b200 import _lib, engine
from tests import op_emulator

T = lambda a: torch.from_numpy(np.asarray(a))


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(engine, "_RUNNER", op_emulator.run_ops)
    monkeypatch.setattr(engine, "_TEST_CPU_PLAN", True)
    op_emulator.LAUNCHED.clear()
    yield


def close(a, b, rtol=1e-3, atol=1e-4):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {err}"


def test_product_path_refuses_cpu(sid_weights):
    net, _ = sid_weights
    with pytest.raises(_lib.UcdirLibraryError):
        net.denoise_fn(torch.zeros(1, 6, 64, 64), torch.zeros(1, 1), torch.zeros(1, 3, 64, 64))
    with pytest.raises(_lib.UcdirLibraryError):
        net.predictor(torch.zeros(1, 3, 64, 64))


def test_unet_forward_direct_and_naive(emulated, golden, sid_weights):
    net, _ = sid_weights
    g = golden("unet")
    eps = net.denoise_fn(T(g["x6"]), T(g["level"]), T(g["guide"]))          # pads 64 -> 96
    close(eps, g["eps"])
    eps2 = net.denoise_fn.naiveforward(T(g["xs"]), T(g["lv2"]), T(g["gs"]))  # batch 2, per-sample levels
    close(eps2, g["eps2"])
    # every op the engine emitted validates against the C ABI's own argument checks (no GPU needed)
    sess = next(iter(net.denoise_fn.engine()._sessions.values()))
    _lib.check_ops(sess.step_ops.array(), len(sess.step_ops))
    _lib.check_ops(sess.static_ops.array(), len(sess.static_ops))


def test_predictor(emulated, golden, sid_weights):
    net, _ = sid_weights
    g = golden("unet")
    close(net.predictor(T(g["xp"])), g["pred"], rtol=1e-4, atol=1e-5)


def test_tiler(emulated, golden, sid_weights, monkeypatch):
    net, _ = sid_weights
    g = golden("tiler")
    skip, padding = (int(v) for v in g["geom"])
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", skip)
    monkeypatch.setattr(unet, "tile_padding", padding)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    unet.engine()._sessions.clear()
    out = unet(T(g["x"]), T(g["level"]), T(g["guide"]))
    close(out, g["out"])


def test_tiler_chunked_equals_unchunked(emulated, golden, sid_weights, monkeypatch):
    net, _ = sid_weights
    g = golden("tiler")
    unet = net.denoise_fn
    monkeypatch.setattr(unet, "tile_skip", 64); monkeypatch.setattr(unet, "tile_padding", 16)
    monkeypatch.setattr(unet, "tile_trigger", 0)
    monkeypatch.setenv("UCDIR_CHUNK_PIXELS", str(64 * 64 * 2))              # 2 tiles per chunk, 3x3 = 9 tiles
    unet.engine()._sessions.clear()
    out = unet(T(g["x"]), T(g["level"]), T(g["guide"]))
    sess = next(iter(unet.engine()._sessions.values()))
    assert len(sess.chunks) == 5
    close(out, g["out"])


def test_super_resolution_e2e(emulated, golden, sid_weights):
    net, _ = sid_weights
    g = golden("sr_e2e")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cpu"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    try:
        out = net.super_resolution(T(g["x_in"]), True)
    finally:
        net._noise_source = None
    close(net.pre_initx, g["initx"], rtol=1e-4, atol=1e-5)
    close(out, g["out"])
    # launches per step are what bench.py reports as gpu_launches
    sess = next(iter(net.denoise_fn.engine()._sessions.values()))
    assert sess.launches_per_step() > 100


def _variant_net(name):
    """A diffusion wrapper variant with the seeded sid weights (same constructor order => same state_dict)."""
    from ucdir_b200.model.networks import define_G
    opt = dict(ucdir_b200.SID_MODEL_OPT, diffusion_name=name)
    torch.manual_seed(1234)
    return define_G({"model": opt})


def test_degraded_guidance_wrapper(emulated, golden, sid_weights):
    """ResiGaussianGuideDY_de.super_resolution (model/diffusion.py:518-523) against the reference's own output."""
    _, sd = sid_weights
    net = _variant_net("ResiGaussianGuideDY_de")
    assert list(net.state_dict().keys()) == list(sd.keys())
    g = golden("sr_de")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cpu"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    out = net.super_resolution(T(g["x_in"]), False)
    close(net.pre_initx, g["initx"], rtol=1e-4, atol=1e-5)
    close(out, g["out"])


def test_initxloss_wrapper_is_guide_dy_at_inference(emulated, golden, sid_weights):
    """ResiGaussianGuideDY_initxloss (model/diffusion.py:528-571) differs only in its training loss."""
    net = _variant_net("ResiGaussianGuideDY_initxloss")
    g = golden("sr_e2e")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cpu"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    out = net.super_resolution(T(g["x_in"]), True)
    close(out, g["out"])


def test_bf16_graph_matches_reference_within_bf16_tolerance(emulated, golden, sid_weights):
    """tcgen05-path graph (GroupNorm folded into weights + border-class tables, grouped K chunks with zero
    filled foreign channels, 4-phase upsample convs, bf16 storage) executed by the CPU interpreter."""
    net, _ = sid_weights
    unet = net.denoise_fn
    eng = unet.engine()
    eng.set_precision("bf16")
    try:
        g = golden("unet")
        eps = unet(T(g["x6"]), T(g["level"]), T(g["guide"]))
        sess = next(iter(eng._sessions.values()))
        _lib.check_ops(sess.step_ops.array(), len(sess.step_ops))          # C-side argument validation of every TC op
        err = (eps - T(g["eps"])).abs()
        ref = T(g["eps"]).abs().max().item()
        assert err.max().item() < 0.03 * ref and err.mean().item() < 0.004 * ref, (err.max().item(), err.mean().item(), ref)
        eps2 = unet.naiveforward(T(g["xs"]), T(g["lv2"]), T(g["gs"]))
        err2 = (eps2 - T(g["eps2"])).abs()
        assert err2.max().item() < 0.03 * ref and err2.mean().item() < 0.004 * ref, (err2.max().item(), err2.mean().item())
    finally:
        eng.set_precision("fp32")


def test_film_resnet_block_module_op(emulated, golden):
    """SR3-style FiLM ResnetBlock (GroupNorm with 8 groups, both FiLM variants) as a module-level op,
    weights loaded through the reference's state_dict keys."""
    from ucdir_b200.model.ucdir import ResnetBlock
    g = golden("modules")
    for tag, aff in (("film", False), ("filmaff", True)):
        m = ResnetBlock(16, 32, nl_emb_dim=64, use_affine_level=aff, norm_groups=8)
        sd = {k[len(tag) + 3:]: T(g[k]) for k in g.files if k.startswith(tag + ".w.")}
        m.load_state_dict(sd, strict=True)
        y = m(T(g[tag + ".x"]), T(g[tag + ".t"]))
        close(y, g[tag + ".y"], rtol=1e-4, atol=1e-5)


def test_ddim_sample(emulated, golden, sid_weights):
    net, _ = sid_weights
    g = golden("ddim")
    n, ls, le = g["sched"]
    net.set_new_noise_schedule(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)),
                               torch.device("cpu"))
    noises = iter([T(z) for z in g["noises"]])
    net._noise_source = lambda shape: next(noises)
    try:
        traj = net.ddim_sample(T(g["x_in"]), True, kwargs={"guide": T(g["initx"])})
    finally:
        net._noise_source = None
    close(traj, g["traj"])


def test_boundary_contract_deepcopy_load_state_dict_and_reschedule(emulated, golden, sid_weights):
    """What model/model.py does to the object define_G returns: deep-copies it for EMA (:50), loads a checkpoint into it
    (:236-239, strict=False) and calls set_new_noise_schedule repeatedly (:59-61, sr.py:399)."""
    import copy
    net, sd = sid_weights
    g = golden("unet")
    args = (T(g["x6"]), T(g["level"]), T(g["guide"]))
    base = net.denoise_fn(*args)
    ema = copy.deepcopy(net)                                       # engines are per module; the copy builds its own
    assert ema.denoise_fn._engine is None and set(ema.state_dict()) == set(net.state_dict())
    close(ema.denoise_fn(*args), base, rtol=0, atol=0)
    # loading different weights must invalidate the packed copies
    sd2 = {k: (v * 0.5 if k.endswith("final_conv.3.weight") else v) for k, v in ema.state_dict().items()}
    missing = ema.load_state_dict({k: v for k, v in sd2.items() if not k.startswith("predictor.conv10")}, strict=False)
    assert all(k.startswith("predictor.conv10") for k in missing.missing_keys)
    out2 = ema.denoise_fn(*args)
    bias = sd["denoise_fn.final_conv.3.bias"].view(1, -1, 1, 1)
    close(out2 - bias, (base - bias) * 0.5, rtol=1e-4, atol=1e-5)
    # schedules can be replaced at will; buffers keep the reference's names, dtypes and lengths
    for T_ in (2000, 50, 7):
        ema.set_new_noise_schedule(dict(schedule="linear", n_timestep=T_, linear_start=1e-6, linear_end=0.01), torch.device("cpu"))
        assert ema.num_timesteps == T_ and ema.betas.shape == (T_,) and ema.betas.dtype == torch.float32
        assert ema.sqrt_alphas_cumprod_prev.shape == (T_ + 1,) and ema.sqrt_alphas_cumprod_prev.dtype == np.float64
    with pytest.raises(NotImplementedError):
        ema(T(g["x6"]))                                            # training forward (p_losses) is out of scope


def test_tensor2img_op(emulated, golden):
    """UCDIR_OP_TO_IMAGE_U8 wiring (ucdir_b200.utils.image.tensor2img) against the reference's tensor2img output, bit exact."""
    from ucdir_b200.utils.image import tensor2img
    g = golden("image")
    x = T(g["x"])
    assert np.array_equal(tensor2img(x), g["img"])
    assert np.array_equal(tensor2img(x, crop=int(g["crop"])), g["img_crop"])
    assert np.array_equal(tensor2img(x[0]), g["img"])                      # (C,H,W) input
    with pytest.raises(NotImplementedError):
        tensor2img(x, out_type=np.float32)


def test_ddpm_inference_wrapper(emulated, sid_weights, monkeypatch):
    """DDPMInference.feed_data / test / get_current_visuals (model/model.py:124-138,167-179): reflect pad by 64, sample, crop --
    against the oracle's restatement, T=1 schedule, injected noise."""
    from oracle import ucdir_oracle as O
    from ucdir_b200.model import model as M
    net, sd = sid_weights
    monkeypatch.setattr(M.networks, "define_G", lambda opt: net)
    dd = M.DDPMInference({"model": ucdir_b200.SID_MODEL_OPT}, device="cpu")
    so = dict(schedule="linear", n_timestep=1, linear_start=1e-6, linear_end=0.4)
    dd.set_new_noise_schedule(so, schedule_phase="val")
    gen = torch.Generator().manual_seed(3)
    sr = torch.rand(1, 3, 72, 80, generator=gen) * 2 - 1
    noise = torch.randn(1, 3, 72 + 128, 80 + 128, generator=gen)
    net._noise_source = lambda shape: noise
    try:
        dd.feed_data({"SR": sr.clone(), "HR": sr.clone(), "Index": 0})
        dd.test(continous=False)
    finally:
        net._noise_source = None
    vis = dd.get_current_visuals()
    assert vis["SR"].shape == sr.shape and torch.equal(vis["INF"], sr) and net.training
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    with torch.no_grad():
        want = O.ddpm_test(sd, lay, O.schedule_buffers(so), sr, [noise], continous=False)
    close(vis["SR"], want)
    assert np.array_equal(dd.current_image(), O.tensor2img(vis["SR"])) or \
        np.abs(dd.current_image().astype(int) - O.tensor2img(want).astype(int)).max() <= 1


def test_fused_res_conv_graph_equals_separate_ops(emulated, golden, sid_weights, monkeypatch):
    """bf16 graph with the 1x1 res_conv folded into the conv1 record (UCDIR_TC_I_RES_FUSED, csrc/ucdir_dhalo.cu) against the same
    graph with separate records: fewer ops, identical result in the CPU interpreter, every record valid for the C ABI."""
    net, _ = sid_weights
    unet = net.denoise_fn
    eng = unet.engine()
    eng.set_precision("bf16")
    g = golden("unet")
    outs, n_ops = {}, {}
    try:
        for fuse in (1, 0):
            monkeypatch.setattr(engine, "_TC_FUSE_RES", fuse)
            eng._sessions.clear()
            outs[fuse] = unet(T(g["x6"]), T(g["level"]), T(g["guide"])).clone()
            sess = next(iter(eng._sessions.values()))
            _lib.check_ops(sess.step_ops.array(), len(sess.step_ops))
            n_ops[fuse] = len(sess.step_ops)
            fused = [o for o in sess.step_ops.ops if o.kind == _lib.C["UCDIR_OP_TC_CONV"] and o.i[_lib.C["UCDIR_TC_I_RES_FUSED"]]]
            assert (len(fused) == 7) == bool(fuse)
            assert all(_lib.tc_schedule(o) == 2 for o in fused)
    finally:
        eng.set_precision("fp32")
    assert n_ops[1] == n_ops[0] - 7
    assert torch.equal(outs[1], outs[0])


# --------------------------------------------------------------------------------------------------------------------
# round 2: binding / invalidation regressions
# --------------------------------------------------------------------------------------------------------------------
def _sched(net, T_=2):
    so = dict(schedule="linear", n_timestep=T_, linear_start=1e-6, linear_end=0.4)
    net.set_new_noise_schedule(so, torch.device("cpu"))
    return so


@pytest.mark.parametrize("variant", ["ResiGaussianGuideDY", "ResiGaussianGuideDY_de", "ResiGaussianGuideDY_initxloss"])
def test_same_shape_images_back_to_back_do_not_share_guidance(emulated, sid_weights, variant):
    """A validation loop over same-shape images (sr.py:518-525): every image must get its own guidance maps even when its
    tensors land on the allocation the previous image just freed.  Round 1 keyed the bind on (data_ptr, _version, shape) and
    returned the previous image's result.  Three different images in a loop, previous tensors deleted, each equal to the
    oracle; recycled addresses are forced by reusing one storage through fresh tensor objects."""
    from oracle import ucdir_oracle as O
    _, sd = sid_weights
    net = _variant_net(variant)
    so = _sched(net, 1)
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    gen = torch.Generator().manual_seed(21)
    storage = torch.empty(1, 3, 40, 48)
    noise = torch.randn(1, 3, 40, 48, generator=gen)
    net._noise_source = lambda shape: noise
    outs = []
    for k in range(3):
        img = torch.rand(1, 3, 40, 48, generator=gen) * 2 - 1
        # a NEW tensor object at the SAME address with version 0: what the caching allocator hands out after `del`
        x = torch.frombuffer(memoryview(storage.numpy()), dtype=torch.float32).view(1, 3, 40, 48)
        storage.copy_(img)
        assert x.data_ptr() == storage.data_ptr() and x._version == 0
        out = net.super_resolution(x, False)
        with torch.no_grad():
            want, _ = O.super_resolution(sd, lay, O.schedule_buffers(so), img, [noise], continous=False,
                                         guide_from="input" if variant == "ResiGaussianGuideDY_de" else "initx")
        close(out, want)
        outs.append(out.clone())
        del x, out
    assert not torch.allclose(outs[0], outs[1]) and not torch.allclose(outs[1], outs[2])


def test_reference_style_denoise_fn_loop_rebinds_cond_every_call(emulated, sid_weights):
    """The generic drop-in forward under the REFERENCE's own sampler / dpm_solver wrapper (model/diffusion.py:166,
    sr.py:203-205): `denoise_fn(torch.cat([cond, x_t], 1), level, guide=initx)` with a fresh cat tensor per step (usually at
    the address of the previous one).  cond must be re-copied on every call; the guidance maps (a function of `guide` only)
    are computed once for as long as the caller passes the same guide object."""
    from oracle import ucdir_oracle as O
    net, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    gen = torch.Generator().manual_seed(5)
    cond = torch.rand(1, 3, 40, 48, generator=gen) * 2 - 1
    guide = torch.rand(1, 3, 40, 48, generator=gen) * 2 - 1
    lvl = torch.full((1, 1), 0.7)
    eng = net.denoise_fn.engine()
    eng._sessions.clear()
    storage = torch.empty(1, 6, 40, 48)
    binds = []
    for k in range(3):
        xt = torch.randn(1, 3, 40, 48, generator=gen)
        storage.copy_(torch.cat([cond, xt], 1))
        x6 = torch.frombuffer(memoryview(storage.numpy()), dtype=torch.float32).view(1, 6, 40, 48)   # same address, version 0
        eps = net.denoise_fn(x6, lvl, guide)
        with torch.no_grad():
            want = O.unet_forward(sd, "denoise_fn.", lay, torch.cat([cond, xt], 1), lvl, guide)
        close(eps, want)
        binds.append(next(iter(eng._sessions.values())).n_binds)
    assert binds == [1, 1, 1], binds                   # guidance maps computed once
    guide2 = guide.clone()
    net.denoise_fn(x6, lvl, guide2)
    assert next(iter(eng._sessions.values())).n_binds == 2      # a different guide object recomputes them
    guide2.mul_(0.5)                                            # ... and so does an in-place edit of the bound one
    eps = net.denoise_fn(x6, lvl, guide2)
    assert next(iter(eng._sessions.values())).n_binds == 3
    with torch.no_grad():
        close(eps, O.unet_forward(sd, "denoise_fn.", lay, x6, lvl, guide2))


def test_in_place_parameter_update_repacks_weights(emulated, golden, sid_weights):
    """Optimizers (and any `p.mul_()` / `p.copy_()` under no_grad) write parameters in place; packed kernel weights and plans
    must follow (round 1 only invalidated on load_state_dict / .to()).  Detected through the parameters' version counters;
    writes through a `.data` view bypass those by design and need `engine().invalidate_weights()` (also checked)."""
    import copy
    net, _ = sid_weights
    net = copy.deepcopy(net)
    g = golden("unet")
    args = (T(g["x6"]), T(g["level"]), T(g["guide"]))
    base = net.denoise_fn(*args).clone()
    with torch.no_grad():
        net.denoise_fn.final_conv[3].weight.mul_(0.5)
    bias = net.denoise_fn.final_conv[3].bias.detach().view(1, -1, 1, 1)
    close(net.denoise_fn(*args) - bias, (base - bias) * 0.5, rtol=1e-4, atol=1e-5)
    p0 = net.predictor(T(g["xp"])).clone()
    with torch.no_grad():
        net.predictor.conv10_1.weight.mul_(2.0)
    b10 = net.predictor.conv10_1.bias.detach().view(1, -1, 1, 1)
    close(net.predictor(T(g["xp"])) - b10, (p0 - b10) * 2.0, rtol=1e-4, atol=1e-5)
    net.predictor.conv10_1.weight.data.mul_(0.5)                 # invisible to the version counter
    net.predictor.engine().invalidate_weights()
    close(net.predictor(T(g["xp"])), p0, rtol=1e-4, atol=1e-5)


def test_guideless_wrappers_fail_like_the_reference(emulated, golden, sid_weights):
    """ResiGaussianDiffusion / ResiPercepGaussianDiffusion / NoDiffusion / GaussianDiffusion.super_resolution call denoise_fn
    without `guide` (model/diffusion.py:302-304,428-432,620-622,650-662); with the only UNet the reference ships that is a
    TypeError there (recorded in tests/golden/wrappers.json by make_golden.py) and here.  Constructors keep the state_dict."""
    import json, os
    _, sd = sid_weights
    rec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wrappers.json")))
    x = T(golden("unet")["xp"])
    for name in ("ResiGaussianDiffusion", "ResiPercepGaussianDiffusion", "NoDiffusion"):
        net = _variant_net(name)
        assert list(net.state_dict().keys()) == list(sd.keys())
        for k in ("denoise_fn.final_conv.3.weight", "predictor.conv10_1.weight"):
            assert torch.equal(net.state_dict()[k], sd[k])
        _sched(net, 2)
        assert rec[name]["error"] == "TypeError"
        with pytest.raises(TypeError, match="guide"):
            net.super_resolution(x)
    base = _variant_net("ResiGaussianGuideDY")
    _sched(base, 2)
    assert rec["GaussianDiffusion.super_resolution"]["error"] == "TypeError"
    with pytest.raises(TypeError, match="guide"):
        ucdir_b200.model.diffusion.GaussianDiffusion.super_resolution(base, x)


def test_product_schedule_buffers_bit_exact(golden, sid_weights):
    """The PRODUCT's set_new_noise_schedule (not the oracle's) against the reference's buffers: 12 fp32 vectors + the float64
    numpy attribute, bit for bit, for every schedule in tests/golden/schedule.npz."""
    import ucdir_b200.model.diffusion  # noqa: F401
    net, _ = sid_weights
    g = golden("schedule")
    kinds = {"sidval": "linear", "yamlval": "linear", "train": "linear", "quad": "quad", "warm": "warmup10"}
    for tag, kind in kinds.items():
        n, ls, le = g[f"{tag}.opt"]
        net.set_new_noise_schedule(dict(schedule=kind, n_timestep=int(n), linear_start=float(ls), linear_end=float(le)), torch.device("cpu"))
        checked = 0
        for k in g.files:
            if not k.startswith(tag + ".") or k.endswith(".opt"):
                continue
            field = k.split(".", 1)[1]
            want = g[k]
            have = net.sqrt_alphas_cumprod_prev if field == "sqrt_alphas_cumprod_prev_f64" else getattr(net, field).numpy()
            assert have.dtype == want.dtype and np.array_equal(have, want), (tag, field)
            checked += 1
        assert checked == 13, (tag, checked)
    # schedule arrived through load_state_dict only (model/model.py:236-239): samplers rebuild their host copies lazily
    import copy
    other = copy.deepcopy(net)
    other._sched_host = None
    assert other._step_scalars(1) == net._step_scalars(1)


@pytest.mark.parametrize("precision", ["bf16", "fp32_tc"])
def test_graph_branches_bracket_the_res_conv(emulated, golden, sid_weights, precision):
    """UCDIR_OP_FLAG_BRANCH / JOIN (step graph with branches, include/ucdir_b200.h): every BRANCH op is a block's separate 1x1
    res_conv, directly followed by the conv1 that reads the same input(s) and writes something else, and the next flagged op is
    the JOIN on the integration conv that reads the res_conv's output as its residual -- nothing in between touches it."""
    net, _ = sid_weights
    unet = net.denoise_fn
    eng = unet.engine()
    eng.set_precision(precision)
    try:
        g = golden("unet")
        unet(T(g["x6"]), T(g["level"]), T(g["guide"]))
        sess = next(iter(eng._sessions.values()))
        ops = sess.step_ops.ops
        P = lambda o, k: o.p[_lib.C["UCDIR_TC_P_" + k]]
        I = lambda o, k: o.i[_lib.C["UCDIR_TC_I_" + k]]
        BR, JN = _lib.C["UCDIR_OP_FLAG_BRANCH"], _lib.C["UCDIR_OP_FLAG_JOIN"]
        branches = [k for k, o in enumerate(ops) if o.flags & BR]
        assert branches, "no branch in the step graph"
        for k in branches:
            rc, c1 = ops[k], ops[k + 1]
            assert rc.kind == c1.kind == _lib.C["UCDIR_OP_TC_CONV"]
            assert I(rc, "NTY") == 1 and I(rc, "NTX") == 1 and I(rc, "MODE") == 0 and not I(rc, "GN")          # the 1x1 res_conv
            assert I(c1, "NTY") == 3 and c1.flags == 0 and not I(c1, "RES_FUSED")
            assert P(c1, "SRC0") == P(rc, "SRC0") and P(c1, "SRC1") == P(rc, "SRC1")                            # same input(s)
            assert P(c1, "DST") != P(rc, "DST") and P(rc, "DST") not in (P(c1, "SRC0"), P(c1, "SRC1"))
            mix = ops[k + 2]
            assert (mix.flags & JN) and I(mix, "MODE") == 1 and P(mix, "RES") == P(rc, "DST") and P(mix, "SRC0") == P(c1, "DST")
        # a JOIN without a BRANCH before it does not occur, and blocks with a fused res_conv carry no flags
        assert sum(1 for o in ops if o.flags & JN) == len(branches)
        assert all(o.flags == 0 for o in ops if o.kind == _lib.C["UCDIR_OP_TC_CONV"] and I(o, "RES_FUSED"))
    finally:
        eng.set_precision("fp32")


def test_fp32_tc_graph_meets_fp32_tolerance(emulated, golden, sid_weights):
    """precision "fp32_tc": split-operand (hi + lo bf16 pairs) tensor-core graph -- three K passes per tap, plane-pair
    activations, exact epilogue math -- executed by the CPU interpreter must meet the reference's fp32 tolerance
    (rtol 1e-3 / atol 1e-4), and every record must pass the C ABI's argument checks and route to the kernel that has a split form for it."""
    net, _ = sid_weights
    unet = net.denoise_fn
    eng = unet.engine()
    eng.set_precision("fp32_tc")
    try:
        g = golden("unet")
        eps = unet(T(g["x6"]), T(g["level"]), T(g["guide"]))
        sess = next(iter(eng._sessions.values()))
        _lib.check_ops(sess.step_ops.array(), len(sess.step_ops))
        tc = [o for o in sess.step_ops.ops if o.kind == _lib.C["UCDIR_OP_TC_CONV"]]
        # split records run the streamed kernel, except the integration-module convs with C = 64 / 128 / 256 (halo mix kernel, SPLIT form), the in-conv and conv1 + res_conv (halo dense kernel)
        I = lambda o, k: o.i[_lib.C["UCDIR_TC_I_" + k]]
        assert tc and all(I(o, "SPLIT") == 1 for o in tc)
        for o in tc:
            want = 1 if (I(o, "MODE") == 1 and I(o, "C0") in (64, 128, 256) and I(o, "H") >= 2 and I(o, "W") >= 2) else 0
            if I(o, "MODE") == 0 and I(o, "NTY") == 3 and (I(o, "KC") == 16 or I(o, "RES_FUSED")):
                want = 2                                         # the in-conv and conv1 + fused res_conv: halo dense kernel, SPLIT form
            assert _lib.tc_schedule(o) == want, (I(o, "MODE"), I(o, "C0"), _lib.tc_schedule(o))
        assert any(_lib.tc_schedule(o) == 1 for o in tc)
        close(eps, g["eps"])
        eps2 = unet.naiveforward(T(g["xs"]), T(g["lv2"]), T(g["gs"]))
        close(eps2, g["eps2"])
    finally:
        eng.set_precision("fp32")


def test_reference_ddpm_methods_drive_our_module(emulated, sid_weights):
    """The reference's own `DDPM.feed_data / test / get_current_visuals / set_new_noise_schedule` (model/model.py:101-179,
    unmodified, imported through tests/shims) operating on the module ucdir_b200.define_G returns, kernels emulated on CPU.
    `DDPM.__init__` itself needs CUDA + a process group (it hard-codes torch.device('cuda') and wraps in DDP): that part runs on
    the GPU box in tests/test_gpu_reference_caller.py; here the object is built without it."""
    from oracle import ucdir_oracle as O
    from tests import shims
    if shims.install() is None:
        pytest.skip("reference tree not available")
    import model.model as refmodel
    net, sd = sid_weights
    dd = refmodel.DDPM.__new__(refmodel.DDPM)
    dd.opt, dd.device, dd.netG, dd.schedule_phase = {"phase": "val"}, torch.device("cpu"), net, None
    so = dict(schedule="linear", n_timestep=1, linear_start=1e-6, linear_end=0.4)
    dd.set_new_noise_schedule(so, schedule_phase="val")
    gen = torch.Generator().manual_seed(3)
    sr = torch.rand(1, 3, 72, 80, generator=gen) * 2 - 1
    noise = torch.randn(1, 3, 72 + 128, 80 + 128, generator=gen)
    net._noise_source = lambda shape: noise
    try:
        dd.feed_data({"SR": sr.clone(), "HR": sr.clone(), "Index": torch.tensor([0])})
        dd.test(continous=True)
    finally:
        net._noise_source = None
    vis = dd.get_current_visuals()
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    with torch.no_grad():
        want = O.ddpm_test(sd, lay, O.schedule_buffers(so), sr, [noise], continous=True)
    close(vis["SR"], want)
    assert torch.equal(vis["INF"], sr) and net.training


def test_predictor_tensor_core_plan(emulated, golden, sid_weights):
    """UNetSeeInDark on the tensor-core plan (split operands, channels padded to 64, LeakyReLU epilogue, ConvTranspose as four 1x1
    phase GEMMs, max-pool on plane pairs) against the reference's golden output, fp32 tolerance; records valid for the C ABI."""
    net, _ = sid_weights
    eng = net.predictor.engine()
    eng.set_mode("tc")
    try:
        g = golden("unet")
        y = net.predictor(T(g["xp"]))
        plan = next(iter(eng._plans.values()))
        _lib.check_ops(plan[0].array(), len(plan[0]))
        assert any(o.kind == _lib.C["UCDIR_OP_TC_CONV"] and o.i[_lib.C["UCDIR_TC_I_SPLIT"]] for o in plan[0].ops)
        close(y, g["pred"], rtol=1e-4, atol=2e-5)
    finally:
        eng.set_mode("fp32")
