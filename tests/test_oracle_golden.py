"""Pin the CPU oracle (oracle/ucdir_oracle.py) against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  Oracle and reference both run torch-CPU fp32 here, so the tolerances are
tight: differences come only from op-order restatements (bmm vs einsum, summation order)."""
import numpy as np
import torch

from oracle import ucdir_oracle as O

T = lambda a: torch.from_numpy(np.asarray(a))


def sub_sd(g, prefix):
    return {k[len(prefix):]: T(g[k]) for k in g.files if k.startswith(prefix)}


def close(a, b, rtol=1e-5, atol=1e-6):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), f"max abs err {err}"


def test_schedule_bit_exact(golden):
    g = golden("schedule")
    tags = sorted({k.split(".")[0] for k in g.files})
    kinds = {"sidval": "linear", "yamlval": "linear", "train": "linear", "quad": "quad", "warm": "warmup10"}
    assert set(tags) == set(kinds)
    for tag in tags:
        n, ls, le = g[f"{tag}.opt"]
        out = O.schedule_buffers(dict(schedule=kinds[tag], n_timestep=int(n), linear_start=float(ls), linear_end=float(le)))
        for k, v in out.items():
            ref = g[f"{tag}.{k}"]
            assert ref.dtype == v.dtype and np.array_equal(ref, v), (tag, k)


def test_resblock_dy3h(golden):
    g = golden("modules")
    for tag in ("rb", "rbid"):
        sd = sub_sd(g, f"{tag}.w.")
        y = O.resblock_dy3h(sd, "", T(g[f"{tag}.x"]), T(g[f"{tag}.t"]), T(g[f"{tag}.guide"]))
        close(y, g[f"{tag}.y"])


def test_attention(golden):
    g = golden("modules")
    y = O.self_attention(sub_sd(g, "attn.w."), "", T(g["attn.x"]))
    close(y, g["attn.y"], rtol=1e-5, atol=2e-6)


def test_film_resblock(golden):
    g = golden("modules")
    for tag, aff in (("film", False), ("filmaff", True)):
        y = O.resblock_film(sub_sd(g, f"{tag}.w."), "", T(g[f"{tag}.x"]), T(g[f"{tag}.t"]), norm_groups=8,
                            use_affine_level=aff)
        close(y, g[f"{tag}.y"])


def test_unet_forward_and_predictor(golden, sid_weights):
    import ucdir_b200
    _, sd = sid_weights
    g = golden("unet")
    lay = O.UNetLayout(**{k: v for k, v in ucdir_b200.SID_MODEL_OPT["unet"].items()})
    with torch.no_grad():
        eps = O.unet_forward(sd, "denoise_fn.", lay, T(g["x6"]), T(g["level"]), T(g["guide"]))
        close(eps, g["eps"], rtol=1e-4, atol=1e-5)
        eps2 = O.unet_naiveforward(sd, "denoise_fn.", lay, T(g["xs"]), T(g["lv2"]), T(g["gs"]))
        close(eps2, g["eps2"], rtol=1e-4, atol=1e-5)
        close(O.predictor_forward(sd, "predictor.", T(g["xp"])), g["pred"], rtol=1e-5, atol=1e-6)


def test_super_resolution_e2e(golden, sid_weights):
    import ucdir_b200
    _, sd = sid_weights
    g = golden("sr_e2e")
    n, ls, le = g["sched"]
    sched = O.schedule_buffers(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)))
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    noises = [T(z) for z in g["noises"]]
    with torch.no_grad():
        out, initx = O.super_resolution(sd, lay, sched, T(g["x_in"]), noises, continous=True)
    close(initx, g["initx"], rtol=1e-5, atol=1e-6)
    close(out, g["out"], rtol=1e-4, atol=2e-5)


def test_super_resolution_degraded_guidance(golden, sid_weights):
    """ResiGaussianGuideDY_de (model/diffusion.py:481-523): the degraded input, not initx, guides the integration modules."""
    import ucdir_b200
    _, sd = sid_weights
    g = golden("sr_de")
    n, ls, le = g["sched"]
    sched = O.schedule_buffers(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)))
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    with torch.no_grad():
        out, initx = O.super_resolution(sd, lay, sched, T(g["x_in"]), [T(z) for z in g["noises"]], continous=False, guide_from="input")
        out_wrong, _ = O.super_resolution(sd, lay, sched, T(g["x_in"]), [T(z) for z in g["noises"]], continous=False)
    close(initx, g["initx"], rtol=1e-5, atol=1e-6)
    close(out, g["out"], rtol=1e-4, atol=2e-5)
    assert (out_wrong - T(g["out"])).abs().max().item() > 1e-3       # the guide source matters


def test_tiler(golden, sid_weights):
    import ucdir_b200
    _, sd = sid_weights
    g = golden("tiler")
    skip, padding = (int(v) for v in g["geom"])
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    with torch.no_grad():
        out = O.unet_forward(sd, "denoise_fn.", lay, T(g["x"]), T(g["level"]), T(g["guide"]), skip=skip,
                             padding=padding, force_tiler=True)
    close(out, g["out"], rtol=1e-4, atol=1e-5)


def test_tile_windows_match_reference_rule():
    # utils/util.py:122-134: i in arange(0, L, skip-2*pad), clamped so the window ends at L
    assert O.tile_windows(1280, 1024, 64) == [0, 256]
    assert O.tile_windows(1056, 128, 16) == [min(i, 1056 - 128) for i in range(0, 1056, 96)]
    assert len(O.tile_windows(1056, 128, 16)) == 11
    assert O.tiler_pad(1024, 1024, 128, 16) == 16 and O.tiler_pad(96, 80, 128, 16) == 128 - 80 + 16


def test_ddim_sample(golden, sid_weights):
    """the strided sampler (model/diffusion.py:246-294) restated in the oracle vs the reference's trajectory."""
    import ucdir_b200
    _, sd = sid_weights
    g = golden("ddim")
    n, ls, le = g["sched"]
    sched = O.schedule_buffers(dict(schedule="linear", n_timestep=int(n), linear_start=float(ls), linear_end=float(le)))
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    den = lambda xc, lvl, gd: O.unet_forward(sd, "denoise_fn.", lay, xc, lvl, gd)
    with torch.no_grad():
        traj = O.ddim_sample(sched, den, T(g["x_in"]), T(g["initx"]), [T(z) for z in g["noises"]])
    close(traj, g["traj"], rtol=1e-4, atol=2e-5)


def test_tensor2img(golden):
    """core/metrics.py:8-34 incl. clamping and round-half-to-even ties, against the reference's own output."""
    g = golden("image")
    x = T(g["x"])
    assert np.array_equal(O.tensor2img(x), g["img"])
    c = int(g["crop"])
    assert np.array_equal(O.tensor2img(x[..., c:-c, c:-c]), g["img_crop"])
