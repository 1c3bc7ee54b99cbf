"""CPU oracle for the UCDIR iterative-denoising hot path.

TEST INFRASTRUCTURE ONLY.  This file is the *checker*: a functional, CPU, fp32
restatement (torch.nn.functional on host tensors) of the reference algorithm.  It
is imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  Nothing under ucdir_b200/ imports it, and the product path
never falls back to it.

Parity pinning: the reference ships no tests / golden vectors,
so this restatement is pinned against outputs of the reference itself, generated
in the build container by importing /root/reference (tests/golden/make_golden.py)
and committed under tests/golden/*.npz.  tests/test_oracle_golden.py replays them.

Every function cites the reference file:line it follows (paths relative to the
reference root).  Weights are addressed by the reference's state_dict key names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# noise schedule  (model/diffusion.py:23-54, 101-148)
# --------------------------------------------------------------------------------------
def beta_schedule(schedule: str, n_timestep: int, linear_start: float, linear_end: float) -> np.ndarray:
    """float64 betas; model/diffusion.py:23-54 ('cosine' omitted: never selected by a shipped config)."""
    if schedule == "linear":
        return np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    if schedule == "quad":
        return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    if schedule == "const":
        return linear_end * np.ones(n_timestep, dtype=np.float64)
    if schedule == "jsd":
        return 1.0 / np.linspace(n_timestep, 1, n_timestep, dtype=np.float64)
    if schedule in ("warmup10", "warmup50"):
        frac = 0.1 if schedule == "warmup10" else 0.5
        b = linear_end * np.ones(n_timestep, dtype=np.float64)
        n = int(n_timestep * frac)
        b[:n] = np.linspace(linear_start, linear_end, n, dtype=np.float64)
        return b
    raise NotImplementedError(schedule)


def schedule_buffers(schedule_opt: dict) -> Dict[str, np.ndarray]:
    """The 12 fp32 buffers + the float64 sqrt_alphas_cumprod_prev attribute.

    model/diffusion.py:101-148.  All derived quantities are formed in float64 and
    cast to fp32 last, in exactly the reference's expression order.
    """
    betas = beta_schedule(schedule_opt["schedule"], schedule_opt["n_timestep"],
                          schedule_opt["linear_start"], schedule_opt["linear_end"])
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    out = {"sqrt_alphas_cumprod_prev_f64": np.sqrt(np.append(1.0, ac))}
    f32 = lambda a: np.asarray(a, dtype=np.float64).astype(np.float32)
    out["betas"] = f32(betas)
    out["alphas_cumprod"] = f32(ac)
    out["alphas_cumprod_prev"] = f32(ac_prev)
    out["sqrt_alphas_cumprod"] = f32(np.sqrt(ac))
    out["sqrt_one_minus_alphas_cumprod"] = f32(np.sqrt(1.0 - ac))
    out["log_one_minus_alphas_cumprod"] = f32(np.log(1.0 - ac))
    out["sqrt_recip_alphas_cumprod"] = f32(np.sqrt(1.0 / (ac + 1e-10)))
    out["sqrt_recipm1_alphas_cumprod"] = f32(np.sqrt(1.0 / (ac + 1e-10) - 1))
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    out["posterior_variance"] = f32(pv)
    out["posterior_log_variance_clipped"] = f32(np.log(np.maximum(pv, 1e-20)))
    out["posterior_mean_coef1"] = f32(betas * np.sqrt(ac_prev) / (1.0 - ac))
    out["posterior_mean_coef2"] = f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac))
    return out


# --------------------------------------------------------------------------------------
# UNet pieces (model/ucdir.py)
# --------------------------------------------------------------------------------------
def swish(x: Tensor) -> Tensor:
    """model/ucdir.py:48-50."""
    return x * torch.sigmoid(x)


def positional_encoding(noise_level: Tensor, dim: int) -> Tensor:
    """model/ucdir.py:24-29.  noise_level (B,1) -> (B,1,dim)."""
    count = dim // 2
    step = torch.arange(count, dtype=noise_level.dtype, device=noise_level.device) / count
    enc = noise_level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    return torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)


def noise_level_mlp(sd: SD, pre: str, noise_level: Tensor, inner: int) -> Tensor:
    """model/ucdir.py:212-214."""
    e = positional_encoding(noise_level, inner)
    h = F.linear(e, sd[pre + "noise_level_mlp.1.weight"], sd[pre + "noise_level_mlp.1.bias"])
    return F.linear(swish(h), sd[pre + "noise_level_mlp.3.weight"], sd[pre + "noise_level_mlp.3.bias"])


def block_attw(sd: SD, pre: str, t_emb: Tensor) -> Tensor:
    """Per-block timestep weights attw (B,8): model/ucdir.py:106,125."""
    b = t_emb.shape[0]
    h = F.linear(t_emb, sd[pre + "noise_func.0.weight"], sd[pre + "noise_func.0.bias"])
    h = F.linear(swish(h), sd[pre + "noise_func.2.weight"], sd[pre + "noise_func.2.bias"])
    return h.view(b, -1)


def guidance_map(sd: SD, pre: str, guide: Tensor, width: int) -> Tensor:
    """Step-invariant half of att_sp: model/ucdir.py:133-135 without the attw factor."""
    ratio = width / guide.shape[-1]
    g = F.interpolate(guide, scale_factor=ratio, mode="bilinear", align_corners=False)
    g = F.conv2d(g, sd[pre + "conv2.0.weight"], sd[pre + "conv2.0.bias"])
    a, b = g.chunk(2, dim=1)                                   # SimpleGate, ucdir.py:149-152
    return F.conv2d(a * b, sd[pre + "conv2.2.weight"], sd[pre + "conv2.2.bias"], padding=1)


def resblock_dy3h(sd: SD, pre: str, x: Tensor, t_emb: Tensor, guide: Tensor, nset: int = 8) -> Tensor:
    """ResnetBlockDY3h.forward, model/ucdir.py:122-140.  pre ends with 'res_block.'"""
    b, _, hh, ww = x.shape
    attw = block_attw(sd, pre, t_emb)
    h = F.group_norm(x, 1, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps=1e-5)
    h = F.conv2d(h, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    h = swish(h)
    h = F.group_norm(h, 1, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps=1e-5)
    att_sp = guidance_map(sd, pre, guide, ww) * attw.view(b, nset, 1, 1)
    cout = sd[pre + "conv1.weight"].shape[0]
    hset = F.conv2d(h, sd[pre + "spdyconv.weight"], sd[pre + "spdyconv.bias"], padding=1, groups=nset)
    hset = hset.view(b, cout, nset, hh, ww)
    h = torch.sum(hset * att_sp.unsqueeze(1), dim=2)
    h = swish(h)
    if (pre + "res_conv.weight") in sd:
        res = F.conv2d(x, sd[pre + "res_conv.weight"], sd[pre + "res_conv.bias"])
    else:
        res = x
    return h + res


def self_attention(sd: SD, pre: str, x: Tensor, norm_groups: int = 1) -> Tensor:
    """SelfAttention.forward (n_head=1), model/ucdir.py:165-182.  pre ends with 'attn.'"""
    b, c, hh, ww = x.shape
    n = F.group_norm(x, norm_groups, sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps=1e-5)
    qkv = F.conv2d(n, sd[pre + "qkv.weight"])
    q, k, v = qkv.view(b, 3, c, hh * ww).unbind(1)              # (b, c, N) each
    attn = torch.bmm(q.transpose(1, 2), k) / math.sqrt(c)       # (b, Nq, Nk)
    attn = torch.softmax(attn, dim=-1)
    out = torch.bmm(v, attn.transpose(1, 2)).view(b, c, hh, ww)  # out[c, q] = sum_k attn[q,k] v[c,k]
    out = F.conv2d(out, sd[pre + "out.weight"], sd[pre + "out.bias"])
    return out + x


def resblock_film(sd: SD, pre: str, x: Tensor, t_emb: Tensor, norm_groups: int = 32,
                  use_affine_level: bool = False) -> Tensor:
    """Secondary SR3-style FiLM block: ResnetBlock/Block/FeatureWiseAffine, model/ucdir.py:32-45,75-100."""
    b = x.shape[0]
    h = F.group_norm(x, norm_groups, sd[pre + "block1.block.0.weight"], sd[pre + "block1.block.0.bias"], eps=1e-5)
    h = F.conv2d(swish(h), sd[pre + "block1.block.3.weight"], sd[pre + "block1.block.3.bias"], padding=1)
    e = F.linear(t_emb, sd[pre + "noise_func.noise_func.0.weight"], sd[pre + "noise_func.noise_func.0.bias"])
    if use_affine_level:
        gamma, beta = e.view(b, -1, 1, 1).chunk(2, dim=1)
        h = (1 + gamma) * h + beta
    else:
        h = h + e.view(b, -1, 1, 1)
    h = F.group_norm(h, norm_groups, sd[pre + "block2.block.0.weight"], sd[pre + "block2.block.0.bias"], eps=1e-5)
    h = F.conv2d(swish(h), sd[pre + "block2.block.3.weight"], sd[pre + "block2.block.3.bias"], padding=1)
    if (pre + "res_conv.weight") in sd:
        return h + F.conv2d(x, sd[pre + "res_conv.weight"], sd[pre + "res_conv.bias"])
    return h + x


class UNetLayout:
    """Layer list of DY3h as its constructor builds it (model/ucdir.py:205-268)."""

    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=1,
                 channel_mults=(1, 2, 4, 8, 8), attn_res=(8,), res_blocks=3, dropout=0,
                 with_noise_level_emb=True, image_size=128, resname="ResnetBlockDY3h"):
        self.inner = inner_channel
        self.in_channel, self.out_channel = in_channel, out_channel
        nm = len(channel_mults)
        pre = inner_channel
        feat = [pre]
        res = image_size
        self.downs: List[tuple] = [("conv", in_channel, inner_channel)]
        for ind in range(nm):
            last = ind == nm - 1
            use_attn = res in attn_res
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks):
                self.downs.append(("block", pre, cm, use_attn))
                feat.append(cm)
                pre = cm
            if not last:
                self.downs.append(("down", pre))
                feat.append(pre)
                res //= 2
        self.mid = [("block", pre, pre, True), ("block", pre, pre, False)]
        self.ups: List[tuple] = []
        for ind in reversed(range(nm)):
            last = ind < 1
            use_attn = res in attn_res
            cm = inner_channel * channel_mults[ind]
            for _ in range(res_blocks + 1):
                self.ups.append(("block", pre + feat.pop(), cm, use_attn))
                pre = cm
            if not last:
                self.ups.append(("up", pre))
                res *= 2
        self.final_in = pre


def unet_naiveforward(sd: SD, pre: str, layout: UNetLayout, x: Tensor, time: Tensor, guide: Tensor) -> Tensor:
    """DY3h.naiveforward, model/ucdir.py:270-293.  pre is e.g. 'denoise_fn.'"""
    t = noise_level_mlp(sd, pre, time, layout.inner)
    feats = []

    def run(kind, name, spec, x):
        if kind == "block":
            x = resblock_dy3h(sd, name + "res_block.", x, t, guide)
            if spec[3]:
                x = self_attention(sd, name + "attn.", x)
            return x
        if kind == "conv":
            return F.conv2d(x, sd[name + "weight"], sd[name + "bias"], padding=1)
        if kind == "down":                                     # ucdir.py:63-69
            return F.conv2d(x, sd[name + "conv.weight"], sd[name + "conv.bias"], stride=2, padding=1)
        if kind == "up":                                       # ucdir.py:53-60
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            return F.conv2d(x, sd[name + "conv.weight"], sd[name + "conv.bias"], padding=1)
        raise ValueError(kind)

    for i, spec in enumerate(layout.downs):
        x = run(spec[0], f"{pre}downs.{i}.", spec, x)
        feats.append(x)
    for i, spec in enumerate(layout.mid):
        x = run(spec[0], f"{pre}mid.{i}.", spec, x)
    for i, spec in enumerate(layout.ups):
        if spec[0] == "block":
            x = torch.cat((x, feats.pop()), dim=1)
        x = run(spec[0], f"{pre}ups.{i}.", spec, x)
    x = F.group_norm(x, 1, sd[pre + "final_conv.0.weight"], sd[pre + "final_conv.0.bias"], eps=1e-5)
    return F.conv2d(swish(x), sd[pre + "final_conv.3.weight"], sd[pre + "final_conv.3.bias"], padding=1)


def tile_windows(length: int, skip: int, padding: int) -> List[int]:
    """Window start list along one (already padded) axis: utils/util.py:122-134."""
    shift = skip - 2 * padding
    assert shift > 0, "stride skip-2*padding must be positive"
    return [min(i, length - skip) for i in range(0, length, shift)]


def tiler_pad(h: int, w: int, skip: int, padding: int) -> int:
    """utils/util.py:114-115."""
    pd = min(h, w)
    return skip - pd + padding if pd < skip else padding


def patch_forward_guide(noisy: Tensor, net, time: Tensor, guide: Tensor, skip: int, padding: int) -> Tensor:
    """utils/util.py:108-146 (without the is_cuda assert); tiles run sequentially, later tiles overwrite."""
    pd = tiler_pad(noisy.shape[-2], noisy.shape[-1], skip, padding)
    noisy = F.pad(noisy, (pd, pd, pd, pd), mode="reflect")
    guide_pad = F.pad(guide, (pd, pd, pd, pd), mode="reflect")
    den = torch.zeros_like(noisy)[:, :3]
    hh, ww = noisy.shape[-2:]
    for hs in tile_windows(hh, skip, padding):
        for ws in tile_windows(ww, skip, padding):
            out = net(noisy[..., hs:hs + skip, ws:ws + skip], time, guide_pad[..., hs:hs + skip, ws:ws + skip])
            den[..., hs + padding:hs + skip - padding, ws + padding:ws + skip - padding] = \
                out[..., padding:-padding, padding:-padding]
    return den[..., pd:-pd, pd:-pd]


def unet_forward(sd: SD, pre: str, layout: UNetLayout, x: Tensor, time: Tensor, guide: Tensor,
                 skip: int = 1024, padding: int = 64, force_tiler: bool = False) -> Tensor:
    """DY3h.forward, model/ucdir.py:295-307 (tile geometry exposed as parameters; defaults = reference)."""
    h, w = x.shape[-2:]
    net = lambda a, t, g: unet_naiveforward(sd, pre, layout, a, t, g)
    if h * w > 1024 * 1024 or force_tiler:
        return patch_forward_guide(x, net, time, guide, skip, padding)
    fac = 32
    padh, padw = (h // fac + 1) * fac - h, (w // fac + 1) * fac - w
    xp = F.pad(x, (0, padw, 0, padh), mode="reflect")
    gp = F.pad(guide, (0, padw, 0, padh), mode="reflect")
    return net(xp, time, gp)[..., :-padh, :-padw]


def predictor_forward(sd: SD, pre: str, x: Tensor) -> Tensor:
    """UNetSeeInDark.forward, model/ucdir.py:352-403.  pre is e.g. 'predictor.'"""
    h, w = x.shape[-2:]
    fac = 32
    padh, padw = (h // fac + 1) * fac - h, (w // fac + 1) * fac - w
    x = F.pad(x, (0, padw, 0, padh), mode="reflect")
    lrelu = lambda v: torch.max(0.2 * v, v)                     # ucdir.py:414-416
    c = lambda name, v, p=1: F.conv2d(v, sd[pre + name + ".weight"], sd[pre + name + ".bias"], padding=p)
    up = lambda name, v: F.conv_transpose2d(v, sd[pre + name + ".weight"], sd[pre + name + ".bias"], stride=2)
    c1 = lrelu(c("conv1_2", lrelu(c("conv1_1", x))))
    c2 = lrelu(c("conv2_2", lrelu(c("conv2_1", F.max_pool2d(c1, 2)))))
    c3 = lrelu(c("conv3_2", lrelu(c("conv3_1", F.max_pool2d(c2, 2)))))
    c4 = lrelu(c("conv4_2", lrelu(c("conv4_1", F.max_pool2d(c3, 2)))))
    c5 = lrelu(c("conv5_2", lrelu(c("conv5_1", F.max_pool2d(c4, 2)))))
    c6 = lrelu(c("conv6_2", lrelu(c("conv6_1", torch.cat([up("upv6", c5), c4], 1)))))
    c7 = lrelu(c("conv7_2", lrelu(c("conv7_1", torch.cat([up("upv7", c6), c3], 1)))))
    c8 = lrelu(c("conv8_2", lrelu(c("conv8_1", torch.cat([up("upv8", c7), c2], 1)))))
    c9 = lrelu(c("conv9_2", lrelu(c("conv9_1", torch.cat([up("upv9", c8), c1], 1)))))
    return c("conv10_1", c9, 0)[..., :-padh, :-padw]


# --------------------------------------------------------------------------------------
# sampler (model/diffusion.py)
# --------------------------------------------------------------------------------------
def p_sample(sched: Dict[str, np.ndarray], denoise, x: Tensor, t: int, cond: Tensor, guide: Tensor,
             noise: Optional[Tensor], return_eps: bool = False):
    """p_sample + p_mean_variance + predict_start_from_noise + q_posterior,
    model/diffusion.py:150-183.  `noise` is the injected z (ignored at t == 0)."""
    b = x.shape[0]
    level = torch.FloatTensor([sched["sqrt_alphas_cumprod_prev_f64"][t + 1]]).repeat(b, 1)
    eps = denoise(torch.cat([cond, x], dim=1), level, guide)
    a = torch.tensor(sched["sqrt_recip_alphas_cumprod"][t])
    bb = torch.tensor(sched["sqrt_recipm1_alphas_cumprod"][t])
    x0 = a * x - bb * eps
    x0 = x0.clamp(-1.0, 1.0)
    mean = torch.tensor(sched["posterior_mean_coef1"][t]) * x0 + torch.tensor(sched["posterior_mean_coef2"][t]) * x
    logvar = torch.tensor(sched["posterior_log_variance_clipped"][t])
    z = noise if t > 0 else torch.zeros_like(x)
    out = mean + z * (0.5 * logvar).exp()
    return (out, eps) if return_eps else out


def p_sample_loop(sched, denoise, x_in: Tensor, guide: Tensor, noises: Sequence[Tensor], continous: bool):
    """model/diffusion.py:185-211 (conditional branch).  noises[0] is the initial img,
    noises[1..] are consumed in call order by the steps with t > 0."""
    T = len(sched["betas"])
    inter = 1 | (T // 10)
    img = noises[0]
    ret = x_in
    k = 1
    for i in reversed(range(T)):
        z = None
        if i > 0:
            z = noises[k]
            k += 1
        img = p_sample(sched, denoise, img, i, x_in, guide, z)
        if i % inter == 0:
            ret = torch.cat([ret, img], dim=0)
    return ret if continous else ret[-1]


def ddim_sample(sched: Dict[str, np.ndarray], denoise, x_in: Tensor, guide: Tensor, noises: Sequence[Tensor],
                sampling_timesteps: int = 5, eta: float = 1.0) -> Tensor:
    """GaussianDiffusion.ddim_sample + model_predictions, model/diffusion.py:213-294 (objective 'pred_noise',
    clip_x_start=True, no re-derivation of the noise).  noises[0] is the initial image, noises[1..] the per-step z.
    Returns the stacked trajectory [B, 1 + steps, C, H, W] (continous=True layout)."""
    T = len(sched["betas"])
    times = list(reversed(torch.linspace(-1, T - 1, steps=sampling_timesteps + 1).int().tolist()))
    ac = torch.from_numpy(sched["alphas_cumprod"])
    img = noises[0]
    imgs = [img]
    k = 1
    b = x_in.shape[0]
    for time, time_next in zip(times[:-1], times[1:]):
        level = torch.FloatTensor([sched["sqrt_alphas_cumprod_prev_f64"][time + 1]]).repeat(b, 1)
        eps = denoise(torch.cat([x_in, img], dim=1), level, guide)
        x0 = torch.tensor(sched["sqrt_recip_alphas_cumprod"][time]) * img - torch.tensor(sched["sqrt_recipm1_alphas_cumprod"][time]) * eps
        x0 = x0.clamp(-1.0, 1.0)
        if time_next < 0:
            img = x0
            imgs.append(img)
            continue
        alpha, alpha_next = ac[time], ac[time_next]
        sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        img = x0 * alpha_next.sqrt() + c * eps + sigma * noises[k]
        k += 1
        imgs.append(img)
    return torch.stack(imgs, dim=1)


def super_resolution(sd: SD, layout: UNetLayout, sched, x_in: Tensor, noises: Sequence[Tensor],
                     continous: bool = False, skip: int = 1024, padding: int = 64,
                     force_tiler: bool = False, guide_from: str = "initx") -> Tuple[Tensor, Tensor]:
    """ResiGaussianGuideDY.super_resolution, model/diffusion.py:473-478 (guide_from="initx"), and
    ResiGaussianGuideDY_de.super_resolution, model/diffusion.py:518-523 (guide_from="input": the degraded input guides).
    Returns (result, initx)."""
    initx = predictor_forward(sd, "predictor.", x_in)
    den = lambda xc, lvl, g: unet_forward(sd, "denoise_fn.", layout, xc, lvl, g, skip, padding, force_tiler)
    guide = initx if guide_from == "initx" else x_in
    return p_sample_loop(sched, den, x_in, guide, noises, continous) + initx, initx


def tensor2img(tensor: Tensor, min_max=(-1, 1)) -> np.ndarray:
    """core/metrics.py:8-34 for the 3D / batch-1 4D uint8 case: clamp, rescale to [0,1], HWC, * 255.0, round (numpy: half to
    even), uint8."""
    t = tensor.squeeze().float().cpu().clamp(*min_max)
    t = (t - min_max[0]) / (min_max[1] - min_max[0])
    img = np.transpose(t.numpy(), (1, 2, 0)) if t.dim() == 3 else t.numpy()
    return (img * 255.0).round().astype(np.uint8)


def ddpm_test(sd: SD, layout: UNetLayout, sched, sr: Tensor, noises: Sequence[Tensor], continous: bool = False, pd: int = 64) -> Tensor:
    """DDPM.test, model/model.py:124-138: reflect-pad the degraded input by 64, super_resolution, crop."""
    x = F.pad(sr, (pd, pd, pd, pd), mode="reflect")
    out, _ = super_resolution(sd, layout, sched, x, noises, continous)
    return out[..., pd:-pd, pd:-pd]
