"""Host-side mirror of the reference's denoiser / predictor modules.

These classes keep the reference's constructor signatures, attribute names and
state_dict keys (model/ucdir.py:103-140, 155-198, 204-268, 310-350 in the
reference) so checkpoints load unchanged and `torch.manual_seed(s); define_G(opt)`
draws bit-identical initial weights (same parameter-creation order).  They hold
parameters only: all arithmetic runs in the sm_100a kernels of
libucdir_b200.so through `ucdir_b200.engine` -- there is no PyTorch/CPU compute
path here, and calling them without the CUDA library raises.
"""
from __future__ import annotations

import os

import torch
from torch import nn


class Swish(nn.Module):
    """Placeholder keeping Sequential indices aligned with the reference (ucdir.py:48-50)."""

    def forward(self, x):  # pragma: no cover - never executed, kernels fuse the activation
        raise RuntimeError("ucdir_b200: Swish is fused into the CUDA kernels; no eager path")


class SimpleGate(Swish):
    """Index placeholder for ucdir.py:149-152 (fused into the guidance kernel)."""


class PositionalEncoding(nn.Module):
    """Parameter-free; ucdir.py:19-29.  Computed inside the time-embedding kernel."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class FeatureWiseAffine(nn.Module):
    """Parameter container for ucdir.py:32-45 (FiLM: h + Linear(t) or (1 + gamma) * h + beta)."""

    def __init__(self, in_channels, out_channels, use_affine_level=False):
        super().__init__()
        self.use_affine_level = use_affine_level
        self.noise_func = nn.Sequential(nn.Linear(in_channels, out_channels * (1 + self.use_affine_level)))


class Block(nn.Module):
    """Parameter container for ucdir.py:75-83 (GroupNorm(groups) -> Swish -> [Dropout] -> conv3x3)."""

    def __init__(self, dim, dim_out, groups=32, dropout=0):
        super().__init__()
        self.block = nn.Sequential(nn.GroupNorm(groups, dim), Swish(),
                                   nn.Dropout(dropout) if dropout != 0 else nn.Identity(),
                                   nn.Conv2d(dim, dim_out, 3, padding=1))


class ResnetBlock(nn.Module):
    """SR3-style FiLM residual block, ucdir.py:86-100, as a standalone module-level op (it is not
    reachable through DY3h, whose forward always passes `guide`).  Same constructor, state_dict keys and
    forward(x, time_emb) signature; runs on the fp32 CUDA kernels."""

    def __init__(self, dim, dim_out, nl_emb_dim=None, dropout=0, use_affine_level=False, norm_groups=32):
        super().__init__()
        self.noise_func = FeatureWiseAffine(nl_emb_dim, dim_out, use_affine_level)
        self.block1 = Block(dim, dim_out, groups=norm_groups)
        self.block2 = Block(dim_out, dim_out, groups=norm_groups, dropout=dropout)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out, self.norm_groups = dim, dim_out, norm_groups

    @torch.no_grad()
    def forward(self, x, time_emb):
        from ..engine import run_film_block
        return run_film_block(self, x, time_emb)


class ResnetBlockDY3h(nn.Module):
    """Parameter container for ucdir.py:103-120 (ResBlock + spatially-adaptive integration)."""

    def __init__(self, dim, dim_out, nl_emb_dim=None, dropout=0, use_affine_level=False, norm_groups=1, nset=8):
        super().__init__()
        if norm_groups != 1 or nset != 8:
            raise NotImplementedError("ucdir_b200 kernels implement GroupNorm(1, C) and nset=8 (config/sid.yaml)")
        self.noise_func = nn.Sequential(nn.Linear(nl_emb_dim, nset), Swish(), nn.Linear(nset, nset))
        self.nset = nset
        self.dim = dim
        self.dim_out = dim_out
        self.norm1 = nn.GroupNorm(norm_groups, dim)
        self.conv1 = nn.Conv2d(dim, dim_out, 3, padding=1)
        self.norm2 = nn.GroupNorm(norm_groups, dim_out)
        self.conv2 = nn.Sequential(nn.Conv2d(3, nset * 2, 1), SimpleGate(),
                                   nn.Conv2d(nset, nset, kernel_size=3, padding=1))
        self.spdyconv = nn.Conv2d(dim_out, dim_out * nset, kernel_size=3, padding=1, groups=nset)
        self.swish = Swish()
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class SelfAttention(nn.Module):
    """Parameter container for ucdir.py:155-163 (single head, d = C)."""

    def __init__(self, in_channel, n_head=1, norm_groups=32):
        super().__init__()
        if n_head != 1 or norm_groups != 1:
            raise NotImplementedError("ucdir_b200 attention kernel: n_head=1, GroupNorm(1, C)")
        self.n_head = n_head
        self.norm = nn.GroupNorm(norm_groups, in_channel)
        self.qkv = nn.Conv2d(in_channel, in_channel * 3, 1, bias=False)
        self.out = nn.Conv2d(in_channel, in_channel, 1)


class ResnetBlocWithAttn(nn.Module):
    """ucdir.py:185-198.  Only resname='ResnetBlockDY3h' is reachable through DY3h."""

    def __init__(self, dim, dim_out, *, nl_emb_dim=None, norm_groups=1, dropout=0, with_attn=False,
                 resname="ResnetBlockDY3h"):
        super().__init__()
        if resname != "ResnetBlockDY3h":
            raise NotImplementedError("DY3h can only run ResnetBlockDY3h blocks (reference ucdir.py:276 passes guide)")
        self.with_attn = with_attn
        self.res_block = ResnetBlockDY3h(dim, dim_out, nl_emb_dim, norm_groups=norm_groups, dropout=dropout)
        if with_attn:
            self.attn = SelfAttention(dim_out, norm_groups=norm_groups)


class Upsample(nn.Module):
    """ucdir.py:53-60 (nearest x2 folded into the conv gather)."""

    def __init__(self, dim):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.conv = nn.Conv2d(dim, dim, 3, padding=1)


class Downsample(nn.Module):
    """ucdir.py:63-69."""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, 2, 1)


def dy3h_plan(in_channel, inner_channel, channel_mults, attn_res, res_blocks, image_size):
    """The layer table of the denoiser: what the reference's constructor (model/ucdir.py:205-268) builds, as data.
    Returns (downs, mid, ups, final_channels); entries are ("conv", cin, cout) | ("block", cin, cout, with_attn) | ("down", c) |
    ("up", c).  Attention is keyed on the NOMINAL resolution (image_size halved per level), not on the actual input size."""
    downs, ups, skip_channels = [("conv", in_channel, inner_channel)], [], [inner_channel]
    c, res, last = inner_channel, image_size, len(channel_mults) - 1
    for level, mult in enumerate(channel_mults):
        for _ in range(res_blocks):
            downs.append(("block", c, inner_channel * mult, res in attn_res))
            c = inner_channel * mult
            skip_channels.append(c)
        if level != last:
            downs.append(("down", c))
            skip_channels.append(c)
            res //= 2
    mid = [("block", c, c, True), ("block", c, c, False)]
    for level in range(last, -1, -1):
        for _ in range(res_blocks + 1):
            ups.append(("block", c + skip_channels.pop(), inner_channel * channel_mults[level], res in attn_res))
            c = inner_channel * channel_mults[level]
        if level:
            ups.append(("up", c))
            res *= 2
    return downs, mid, ups, c


class DY3h(nn.Module):
    """Conditional denoiser UNet (model/ucdir.py:204-307): a parameter container built from `dy3h_plan`, in the reference's
    parameter-creation order (noise_level_mlp, downs, mid, ups, final_conv) so that seeded construction draws the reference's
    weights and `state_dict()` carries its key names.  forward / naiveforward keep the reference signatures (ucdir.py:270-307) and
    run on the CUDA engine."""

    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=1, channel_mults=[1, 2, 4, 8, 8],
                 attn_res=[8], res_blocks=3, dropout=0, with_noise_level_emb=True, image_size=128,
                 resname="ResnetBlockDY3h"):
        super().__init__()
        if not with_noise_level_emb:
            raise NotImplementedError("ucdir_b200: with_noise_level_emb=False is not on the hot path")
        self.cfg = dict(in_channel=in_channel, out_channel=out_channel, inner_channel=inner_channel,
                        channel_mults=list(channel_mults), attn_res=list(attn_res), res_blocks=res_blocks,
                        image_size=image_size)
        self.noise_level_mlp = nn.Sequential(PositionalEncoding(inner_channel),
                                             nn.Linear(inner_channel, inner_channel * 4), Swish(),
                                             nn.Linear(inner_channel * 4, inner_channel))
        downs, mid, ups, self.prec = dy3h_plan(in_channel, inner_channel, channel_mults, attn_res, res_blocks, image_size)

        def make(spec):
            kind = spec[0]
            if kind == "conv":
                return nn.Conv2d(spec[1], spec[2], kernel_size=3, padding=1)
            if kind == "down":
                return Downsample(spec[1])
            if kind == "up":
                return Upsample(spec[1])
            return ResnetBlocWithAttn(spec[1], spec[2], nl_emb_dim=inner_channel, norm_groups=norm_groups, dropout=dropout,
                                      with_attn=spec[3], resname=resname)

        self.downs = nn.ModuleList([make(sp) for sp in downs])
        self.mid = nn.ModuleList([make(sp) for sp in mid])
        self.ups = nn.ModuleList([make(sp) for sp in ups])
        self.final_conv = nn.Sequential(nn.GroupNorm(1, self.prec), Swish(),
                                        nn.Dropout(dropout) if dropout != 0 else nn.Identity(),
                                        nn.Conv2d(self.prec, out_channel if out_channel is not None else in_channel, 3, padding=1))
        # Tile geometry of DY3h.forward (ucdir.py:298-300).  The reference hard-wires (1024, 64) and the
        # 1 Mpx trigger; they are semantic parameters, overridable but never changed silently.
        self.tile_skip = int(os.environ.get("UCDIR_TILE_SKIP", 1024))
        self.tile_padding = int(os.environ.get("UCDIR_TILE_PADDING", 64))
        self.tile_trigger = int(os.environ.get("UCDIR_TILE_TRIGGER", 1024 * 1024))
        self._engine = None

    # ---- engine plumbing -------------------------------------------------------------
    def engine(self):
        from ..engine import UNetEngine
        if self._engine is None:
            self._engine = UNetEngine(self)
        return self._engine

    def __deepcopy__(self, memo):  # model/model.py:50 deep-copies netG for EMA; engines are per-module
        eng, self._engine = self._engine, None
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        return new

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if self._engine is not None:
            self._engine.invalidate_weights()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:
            self._engine.invalidate_weights()
        return out

    # ---- reference API ---------------------------------------------------------------
    @torch.no_grad()
    def naiveforward(self, x, time, guide):
        """ucdir.py:270-293: one UNet evaluation at the given (already padded) size; H, W multiples of 16."""
        return self.engine().forward_plain(x, time, guide)

    @torch.no_grad()
    def forward(self, x, time, guide):
        """ucdir.py:295-307: tiler above the pixel trigger, else reflect-pad bottom/right to (h//32+1)*32."""
        return self.engine().forward(x, time, guide)


PREDICTOR_WIDTHS = (32, 64, 128, 256, 512)


def predictor_plan(in_channels=3, out_channels=3):
    """Layer table of the initial predictor (model/ucdir.py:310-350), in parameter-creation order: two 3x3 convs per encoder level
    with a 2x2 max-pool between levels, then per decoder level a 2x2 stride-2 transposed conv and two 3x3 convs on the concatenated
    features, and a final 1x1 conv.  Entries: (attribute name, kind, cin, cout)."""
    plan, c = [], in_channels
    for lvl, wd in enumerate(PREDICTOR_WIDTHS, start=1):
        plan += [("conv%d_1" % lvl, "conv3", c, wd), ("conv%d_2" % lvl, "conv3", wd, wd)]
        if lvl < len(PREDICTOR_WIDTHS):
            plan.append(("pool%d" % lvl, "pool", wd, wd))
        c = wd
    for lvl, wd in zip(range(6, 10), reversed(PREDICTOR_WIDTHS[:-1])):
        plan += [("upv%d" % lvl, "convT", c, wd), ("conv%d_1" % lvl, "conv3", 2 * wd, wd), ("conv%d_2" % lvl, "conv3", wd, wd)]
        c = wd
    plan.append(("conv10_1", "conv1", c, out_channels))
    return plan


class UNetSeeInDark(nn.Module):
    """Initial predictor; parameter container built from `predictor_plan` (same attribute names and creation order as the
    reference, model/ucdir.py:310-350)."""

    def __init__(self, in_channels=3, out_channels=3):
        super().__init__()
        for name, kind, cin, cout in predictor_plan(in_channels, out_channels):
            if kind == "conv3":
                layer = nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1)
            elif kind == "conv1":
                layer = nn.Conv2d(cin, cout, kernel_size=1, stride=1)
            elif kind == "convT":
                layer = nn.ConvTranspose2d(cin, cout, 2, stride=2)
            else:
                layer = nn.MaxPool2d(kernel_size=2)
            setattr(self, name, layer)
        self._engine = None

    def engine(self):
        from ..engine import PredictorEngine
        if self._engine is None:
            self._engine = PredictorEngine(self)
        return self._engine

    def __deepcopy__(self, memo):
        eng, self._engine = self._engine, None
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        return new

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if self._engine is not None:
            self._engine.invalidate_weights()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if self._engine is not None:
            self._engine.invalidate_weights()
        return out

    @torch.no_grad()
    def forward(self, x):
        """ucdir.py:352-358: reflect-pad to (h//32+1)*32, run, crop."""
        return self.engine().forward(x)
