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


class DY3h(nn.Module):
    """Conditional denoiser UNet; constructor mirrors ucdir.py:204-268 line by line in
    parameter-creation order.  forward/naiveforward keep the reference signatures
    (ucdir.py:270-307) and run on the CUDA engine."""

    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=1, channel_mults=[1, 2, 4, 8, 8],
                 attn_res=[8], res_blocks=3, dropout=0, with_noise_level_emb=True, image_size=128,
                 resname="ResnetBlockDY3h"):
        super().__init__()
        if not with_noise_level_emb:
            raise NotImplementedError("ucdir_b200: with_noise_level_emb=False is not on the hot path")
        self.cfg = dict(in_channel=in_channel, out_channel=out_channel, inner_channel=inner_channel,
                        channel_mults=list(channel_mults), attn_res=list(attn_res), res_blocks=res_blocks,
                        image_size=image_size)
        noise_level_channel = inner_channel
        self.noise_level_mlp = nn.Sequential(PositionalEncoding(inner_channel),
                                             nn.Linear(inner_channel, inner_channel * 4), Swish(),
                                             nn.Linear(inner_channel * 4, inner_channel))
        num_mults = len(channel_mults)
        pre_channel = inner_channel
        feat_channels = [pre_channel]
        now_res = image_size
        downs = [nn.Conv2d(in_channel, inner_channel, kernel_size=3, padding=1)]
        for ind in range(num_mults):
            is_last = ind == num_mults - 1
            use_attn = now_res in attn_res
            channel_mult = inner_channel * channel_mults[ind]
            for _ in range(res_blocks):
                downs.append(ResnetBlocWithAttn(pre_channel, channel_mult, nl_emb_dim=noise_level_channel,
                                                norm_groups=norm_groups, dropout=dropout, with_attn=use_attn,
                                                resname=resname))
                feat_channels.append(channel_mult)
                pre_channel = channel_mult
            if not is_last:
                downs.append(Downsample(pre_channel))
                feat_channels.append(pre_channel)
                now_res //= 2
        self.downs = nn.ModuleList(downs)
        self.mid = nn.ModuleList([
            ResnetBlocWithAttn(pre_channel, pre_channel, nl_emb_dim=noise_level_channel, norm_groups=norm_groups,
                               dropout=dropout, with_attn=True, resname=resname),
            ResnetBlocWithAttn(pre_channel, pre_channel, nl_emb_dim=noise_level_channel, norm_groups=norm_groups,
                               dropout=dropout, with_attn=False, resname=resname)])
        ups = []
        for ind in reversed(range(num_mults)):
            is_last = ind < 1
            use_attn = now_res in attn_res
            channel_mult = inner_channel * channel_mults[ind]
            for _ in range(res_blocks + 1):
                ups.append(ResnetBlocWithAttn(pre_channel + feat_channels.pop(), channel_mult,
                                              nl_emb_dim=noise_level_channel, norm_groups=norm_groups,
                                              dropout=dropout, with_attn=use_attn, resname=resname))
                pre_channel = channel_mult
            if not is_last:
                ups.append(Upsample(pre_channel))
                now_res *= 2
        self.ups = nn.ModuleList(ups)
        self.prec = pre_channel
        dim_out = out_channel if out_channel is not None else in_channel
        self.final_conv = nn.Sequential(nn.GroupNorm(1, pre_channel), Swish(),
                                        nn.Dropout(dropout) if dropout != 0 else nn.Identity(),
                                        nn.Conv2d(pre_channel, dim_out, 3, padding=1))
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


class UNetSeeInDark(nn.Module):
    """Initial predictor; parameter container mirroring ucdir.py:310-350."""

    def __init__(self, in_channels=3, out_channels=3):
        super().__init__()
        self.conv1_1 = nn.Conv2d(in_channels, 32, kernel_size=3, stride=1, padding=1)
        self.conv1_2 = nn.Conv2d(32, 32, kernel_size=3, stride=1, padding=1)
        self.pool1 = nn.MaxPool2d(kernel_size=2)
        self.conv2_1 = nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1)
        self.conv2_2 = nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1)
        self.pool2 = nn.MaxPool2d(kernel_size=2)
        self.conv3_1 = nn.Conv2d(64, 128, kernel_size=3, stride=1, padding=1)
        self.conv3_2 = nn.Conv2d(128, 128, kernel_size=3, stride=1, padding=1)
        self.pool3 = nn.MaxPool2d(kernel_size=2)
        self.conv4_1 = nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1)
        self.conv4_2 = nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1)
        self.pool4 = nn.MaxPool2d(kernel_size=2)
        self.conv5_1 = nn.Conv2d(256, 512, kernel_size=3, stride=1, padding=1)
        self.conv5_2 = nn.Conv2d(512, 512, kernel_size=3, stride=1, padding=1)
        self.upv6 = nn.ConvTranspose2d(512, 256, 2, stride=2)
        self.conv6_1 = nn.Conv2d(512, 256, kernel_size=3, stride=1, padding=1)
        self.conv6_2 = nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1)
        self.upv7 = nn.ConvTranspose2d(256, 128, 2, stride=2)
        self.conv7_1 = nn.Conv2d(256, 128, kernel_size=3, stride=1, padding=1)
        self.conv7_2 = nn.Conv2d(128, 128, kernel_size=3, stride=1, padding=1)
        self.upv8 = nn.ConvTranspose2d(128, 64, 2, stride=2)
        self.conv8_1 = nn.Conv2d(128, 64, kernel_size=3, stride=1, padding=1)
        self.conv8_2 = nn.Conv2d(64, 64, kernel_size=3, stride=1, padding=1)
        self.upv9 = nn.ConvTranspose2d(64, 32, 2, stride=2)
        self.conv9_1 = nn.Conv2d(64, 32, kernel_size=3, stride=1, padding=1)
        self.conv9_2 = nn.Conv2d(32, 32, kernel_size=3, stride=1, padding=1)
        self.conv10_1 = nn.Conv2d(32, out_channels, kernel_size=1, stride=1)
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
