"""Generate the committed golden vectors by running the *reference itself* (read-only import).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The GPU box never runs this; tests read the .npz files.

How the reference is imported: sys.path gets /root/reference, `lpips` is stubbed
(model/diffusion.py:12 imports it, only PerceptualGaussianDiffusion uses it), config/sid.yaml is
read with PyYAML.  The tiler's `assert noisy.is_cuda` (utils/util.py:113) is dropped by exec'ing
the function's own source minus that line; the reference file is untouched.
Noise is injected by patching torch.randn / torch.randn_like inside model.diffusion.
Big weights are not stored: they are re-drawn from torch.manual_seed (bit-identical through
ucdir_b200's mirror constructors, guarded by the stored sha256).
"""
import hashlib
import inspect
import os
import sys
import types

import numpy as np
import torch
import yaml

REF = os.environ.get("UCDIR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.modules["lpips"] = types.ModuleType("lpips")

import model.networks as refnet  # noqa: E402
import model.ucdir as refucdir  # noqa: E402
import model.diffusion as refdiff  # noqa: E402
import utils.util as refutil  # noqa: E402

torch.set_num_threads(os.cpu_count())
WEIGHT_SEED, INPUT_SEED, NOISE_SEED = 1234, 0, 42


def sd_digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def npsd(sd, prefix=""):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    opt = yaml.safe_load(open(os.path.join(REF, "config/sid.yaml")))
    g = torch.Generator().manual_seed(INPUT_SEED)
    rnd = lambda *s: torch.randn(*s, generator=g)
    uni = lambda *s: torch.rand(*s, generator=g) * 2 - 1

    # ---- 1. schedules (model/diffusion.py:101-148) ----
    arrs = {}
    for tag, so in {"sidval": dict(schedule="linear", n_timestep=50, linear_start=1e-6, linear_end=0.4),
                    "yamlval": opt["model"]["beta_schedule"]["val"],
                    "train": opt["model"]["beta_schedule"]["train"],
                    "quad": dict(schedule="quad", n_timestep=20, linear_start=1e-4, linear_end=2e-2),
                    "warm": dict(schedule="warmup10", n_timestep=40, linear_start=1e-4, linear_end=2e-2)}.items():
        gd = refdiff.GaussianDiffusion(None, 128)
        gd.set_new_noise_schedule(so, torch.device("cpu"))
        for k, v in gd.state_dict().items():
            arrs[f"{tag}.{k}"] = v.numpy()
        arrs[f"{tag}.sqrt_alphas_cumprod_prev_f64"] = gd.sqrt_alphas_cumprod_prev
        arrs[f"{tag}.opt"] = np.array([so["n_timestep"], so["linear_start"], so["linear_end"]], dtype=np.float64)
    save("schedule", **arrs)

    # ---- 2. small standalone modules, weights stored ----
    torch.manual_seed(7)
    arrs = {}
    for tag, (cin, cout) in {"rb": (16, 32), "rbid": (16, 16)}.items():
        m = refucdir.ResnetBlockDY3h(cin, cout, nl_emb_dim=64).eval()
        with torch.no_grad():
            m.norm1.weight.copy_(1 + 0.3 * rnd(cin)); m.norm1.bias.copy_(0.2 * rnd(cin))
            m.norm2.weight.copy_(1 + 0.3 * rnd(cout)); m.norm2.bias.copy_(0.2 * rnd(cout))
        x, t, gi = rnd(2, cin, 16, 24), rnd(2, 1, 64), uni(2, 3, 32, 48)
        with torch.no_grad():
            y = m(x, t, gi)
        arrs.update(npsd(m.state_dict(), f"{tag}.w."))
        arrs.update({f"{tag}.x": x.numpy(), f"{tag}.t": t.numpy(), f"{tag}.guide": gi.numpy(), f"{tag}.y": y.numpy()})
    m = refucdir.SelfAttention(32, norm_groups=1).eval()
    with torch.no_grad():
        m.norm.weight.copy_(1 + 0.3 * rnd(32)); m.norm.bias.copy_(0.2 * rnd(32))
        x = 2 * rnd(2, 32, 8, 12)
        y = m(x)
    arrs.update(npsd(m.state_dict(), "attn.w."))
    arrs.update({"attn.x": x.numpy(), "attn.y": y.numpy()})
    for tag, aff in {"film": False, "filmaff": True}.items():
        m = refucdir.ResnetBlock(16, 32, nl_emb_dim=64, use_affine_level=aff, norm_groups=8).eval()
        x, t = rnd(2, 16, 12, 20), rnd(2, 64)
        with torch.no_grad():
            y = m(x, t)
        arrs.update(npsd(m.state_dict(), f"{tag}.w."))
        arrs.update({f"{tag}.x": x.numpy(), f"{tag}.t": t.numpy(), f"{tag}.y": y.numpy()})
    save("modules", **arrs)

    # ---- 3. full sid model, seeded weights ----
    torch.manual_seed(WEIGHT_SEED)
    net = refnet.define_G(opt).eval()
    digest = sd_digest(net.state_dict())
    print("weights sha256", digest)
    unet = net.denoise_fn
    x6, guide = uni(1, 6, 64, 64), uni(1, 3, 64, 64)
    lvl = torch.full((1, 1), 0.62)
    with torch.no_grad():
        eps = unet(x6, lvl, guide)                             # pads to 96 (ucdir.py:303-307)
        xs = uni(2, 6, 32, 64)
        gs = uni(2, 3, 32, 64)
        lv2 = torch.tensor([[0.9], [0.3]])
        eps2 = unet.naiveforward(xs, lv2, gs)                  # batch 2, per-sample levels, non-square
        xp = uni(1, 3, 40, 56)
        pred = net.predictor(xp)
    save("unet", digest=np.array(digest), x6=x6.numpy(), guide=guide.numpy(), level=lvl.numpy(), eps=eps.numpy(),
         xs=xs.numpy(), gs=gs.numpy(), lv2=lv2.numpy(), eps2=eps2.numpy(), xp=xp.numpy(), pred=pred.numpy())

    # ---- 4. end to end: super_resolution, T=4, injected noise (diffusion.py:473-478) ----
    so = dict(schedule="linear", n_timestep=4, linear_start=1e-6, linear_end=0.4)
    net.set_new_noise_schedule(so, torch.device("cpu"))
    x_in = uni(1, 3, 64, 64)
    ng = torch.Generator().manual_seed(NOISE_SEED)
    noises = [torch.randn(1, 3, 64, 64, generator=ng) for _ in range(4)]
    it = iter(noises)
    steps = []
    orig_ps = net.p_sample

    def rec_ps(*a, **k):
        out = orig_ps(*a, **k)
        steps.append(out.clone())
        return out

    net.p_sample = rec_ps
    refdiff.torch.randn, refdiff.torch.randn_like = (lambda *a, **k: next(it)), (lambda *a, **k: next(it))
    try:
        with torch.no_grad():
            out = net.super_resolution(x_in, True)
    finally:
        refdiff.torch.randn, refdiff.torch.randn_like = torch.randn, torch.randn_like
        net.p_sample = orig_ps
    save("sr_e2e", x_in=x_in.numpy(), noises=torch.stack(noises).numpy(), out=out.numpy(),
         steps=torch.stack(steps).numpy(), initx=net.pre_initx.numpy(),
         sched=np.array([4, 1e-6, 0.4], dtype=np.float64))

    # ---- 5. tiler (utils/util.py:108-146) with the full net, (skip,padding)=(64,16) on 96x80 ----
    src = inspect.getsource(refutil.patch_forward_guide).replace("    assert noisy.is_cuda\n", "")
    ns = {"F": torch.nn.functional, "np": np, "torch": torch}
    exec(src, ns)
    xt, gt = uni(1, 6, 96, 80), uni(1, 3, 96, 80)
    lv = torch.full((1, 1), 0.8)
    with torch.no_grad():
        tiled = ns["patch_forward_guide"](xt, unet.naiveforward, params={"time": lv, "guide": gt}, skip=64, padding=16)
    save("tiler", x=xt.numpy(), guide=gt.numpy(), level=lv.numpy(), out=tiled.numpy(),
         geom=np.array([64, 16], dtype=np.int64))


def main_ddim():
    """6. ddim_sample (model/diffusion.py:246-294) on the full sid model: T=10 schedule, 5 sampling steps, eta=1."""
    opt = yaml.safe_load(open(os.path.join(REF, "config/sid.yaml")))
    g = torch.Generator().manual_seed(INPUT_SEED + 5)
    torch.manual_seed(WEIGHT_SEED)
    net = refnet.define_G(opt).eval()
    assert sd_digest(net.state_dict()) == str(np.load(os.path.join(OUT, "unet.npz"))["digest"])
    so = dict(schedule="linear", n_timestep=10, linear_start=1e-6, linear_end=0.4)
    net.set_new_noise_schedule(so, torch.device("cpu"))
    x_in = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        initx = net.predictor(x_in)
    ng = torch.Generator().manual_seed(NOISE_SEED + 1)
    noises = [torch.randn(1, 3, 64, 64, generator=ng) for _ in range(5)]
    it = iter(noises)
    refdiff.torch.randn, refdiff.torch.randn_like = (lambda *a, **k: next(it)), (lambda *a, **k: next(it))
    try:
        with torch.no_grad():
            traj = net.ddim_sample(x_in, continous=True, kwargs={"guide": initx})
    finally:
        refdiff.torch.randn, refdiff.torch.randn_like = torch.randn, torch.randn_like
    save("ddim", x_in=x_in.numpy(), initx=initx.numpy(), noises=torch.stack(noises).numpy(), traj=traj.numpy(),
         sched=np.array([10, 1e-6, 0.4], dtype=np.float64))


def main_variants():
    """7. ResiGaussianGuideDY_de.super_resolution (model/diffusion.py:481-523; guidance = the degraded input): T=4, injected
    noise, continous=False.  Same seed and constructor order as ResiGaussianGuideDY, hence the same weights (sha256 checked)."""
    opt = yaml.safe_load(open(os.path.join(REF, "config/sid.yaml")))
    opt["model"]["diffusion_name"] = "ResiGaussianGuideDY_de"
    g = torch.Generator().manual_seed(INPUT_SEED + 7)
    torch.manual_seed(WEIGHT_SEED)
    net = refnet.define_G(opt).eval()
    assert sd_digest(net.state_dict()) == str(np.load(os.path.join(OUT, "unet.npz"))["digest"])
    so = dict(schedule="linear", n_timestep=4, linear_start=1e-6, linear_end=0.4)
    net.set_new_noise_schedule(so, torch.device("cpu"))
    x_in = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    ng = torch.Generator().manual_seed(NOISE_SEED + 2)
    noises = [torch.randn(1, 3, 64, 64, generator=ng) for _ in range(4)]
    it = iter(noises)
    refdiff.torch.randn, refdiff.torch.randn_like = (lambda *a, **k: next(it)), (lambda *a, **k: next(it))
    try:
        with torch.no_grad():
            out = net.super_resolution(x_in, False)
    finally:
        refdiff.torch.randn, refdiff.torch.randn_like = torch.randn, torch.randn_like
    save("sr_de", x_in=x_in.numpy(), noises=torch.stack(noises).numpy(), out=out.numpy(), initx=net.pre_initx.numpy(),
         sched=np.array([4, 1e-6, 0.4], dtype=np.float64))


def main_image():
    """8. core/metrics.py:8-34 tensor2img on a tensor with out-of-range values and exact rounding ties."""
    import core.metrics as refmetrics
    g = torch.Generator().manual_seed(INPUT_SEED + 8)
    x = torch.randn(1, 3, 40, 56, generator=g) * 0.8
    ties = (torch.arange(0, 256, dtype=torch.float32) + 0.5) / 255.0 * 2 - 1          # values that land on k + 0.5 after * 255
    x[0, 0, 0, :56] = ties[:56]; x[0, 1, 1, :56] = ties[100:156]; x[0, 2, 2, :56] = ties[199:255]
    img = refmetrics.tensor2img(x.clone())
    img_crop = refmetrics.tensor2img(x[..., 8:-8, 8:-8].clone())
    save("image", x=x.numpy(), img=img, img_crop=img_crop, crop=np.array(8))


def main_wrappers():
    """9. What the reference's guide-less wrappers do with the only UNet it ships (DY3h needs `guide`): record the exception
    type each one raises (model/diffusion.py:302-304,428-432,620-622,650-662) -> tests/golden/wrappers.json."""
    import json
    opt = yaml.safe_load(open(os.path.join(REF, "config/sid.yaml")))
    rec = {}
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(INPUT_SEED + 9)) * 2 - 1
    so = dict(schedule="linear", n_timestep=2, linear_start=1e-6, linear_end=0.4)

    def attempt(fn):
        try:
            with torch.no_grad():
                fn()
            return {"error": None}
        except Exception as e:                      # noqa: BLE001 -- recording whatever the reference raises
            return {"error": type(e).__name__, "message": str(e)[:200]}

    for name in ("ResiGaussianDiffusion", "ResiPercepGaussianDiffusion", "NoDiffusion"):
        opt["model"]["diffusion_name"] = name
        torch.manual_seed(WEIGHT_SEED)
        net = refnet.define_G(opt).eval()
        same = sd_digest(net.state_dict()) == str(np.load(os.path.join(OUT, "unet.npz"))["digest"])
        rec[name] = dict(attempt(lambda: (net.set_new_noise_schedule(so, torch.device("cpu")), net.super_resolution(x))),
                         digest_matches_guide_dy=same)
    opt["model"]["diffusion_name"] = "ResiGaussianGuideDY"
    torch.manual_seed(WEIGHT_SEED)
    net = refnet.define_G(opt).eval()
    net.set_new_noise_schedule(so, torch.device("cpu"))
    rec["GaussianDiffusion.super_resolution"] = attempt(lambda: refdiff.GaussianDiffusion.super_resolution(net, x))
    json.dump(rec, open(os.path.join(OUT, "wrappers.json"), "w"), indent=1, sort_keys=True)
    print(rec)


if __name__ == "__main__":
    if "--only-wrappers" in sys.argv:
        main_wrappers()
    elif "--only-image" in sys.argv:
        main_image()
    elif "--only-ddim" in sys.argv:
        main_ddim()
    elif "--only-variants" in sys.argv:
        main_variants()
    else:
        main()
        main_ddim()
        main_variants()
        main_image()
        main_wrappers()
