"""The inference half of the reference's `DDPM` wrapper (model/model.py) around `define_G`'s network: `feed_data`, `test`
(reflect-pad by 64, `super_resolution`, crop: model/model.py:124-138) and `get_current_visuals` (:167-179), so a caller that
only validates / samples (`sr.py -p val`) can use ucdir_b200 without the reference tree.  Training (`optimize_parameters`,
EMA, optimizers, checkpoints) stays in the reference; `DDPM` there works unchanged with `ucdir_b200.define_G`."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import networks


class DDPMInference:
    PAD = 64                                           # model/model.py:127

    def __init__(self, opt, device="cuda"):
        self.opt = opt
        self.device = torch.device(device)
        self.netG = networks.define_G(opt).to(self.device)
        self.schedule_phase = None
        self.data, self.SR = None, None

    def set_new_noise_schedule(self, schedule_opt, schedule_phase="train", force=False):
        """model/model.py:155-163."""
        if self.schedule_phase is None or self.schedule_phase != schedule_phase or force:
            self.schedule_phase = schedule_phase
            self.netG.set_new_noise_schedule(schedule_opt, self.device)

    def feed_data(self, data):
        """model/base_model.py:29-40 (set_device on a dict of tensors)."""
        self.data = {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in data.items()}

    def test(self, continous=False):
        """model/model.py:124-138, including the eval()/train() toggling the reference does around the call."""
        self.netG.eval()
        pd = self.PAD
        self.data["SR"] = F.pad(self.data["SR"], (pd, pd, pd, pd), mode="reflect")
        with torch.no_grad():
            self.SR = self.netG.super_resolution(self.data["SR"], continous)
        self.netG.train()
        self.SR = self.SR[..., pd:-pd, pd:-pd]
        self.data["SR"] = self.data["SR"][..., pd:-pd, pd:-pd]

    def get_current_visuals(self, need_LR=True, sample=False):
        """model/model.py:167-179."""
        out = OrderedDict()
        if sample:
            out["SAM"] = self.SR.detach().float().cpu()
        else:
            out["SR"] = self.SR.detach().float().cpu()
            out["INF"] = self.data["SR"].detach().float().cpu()
            if "HR" in self.data:
                out["HR"] = self.data["HR"].detach().float().cpu()
            out["LR"] = self.data["LR"].detach().float().cpu() if need_LR and "LR" in self.data else out["INF"]
        return out

    def current_image(self, min_max=(-1, 1)):
        """`tensor2img(visuals['SR'])` of sr.py without the fp32 round trip through host memory: the last row of `self.SR`
        (the final sample when continous) converted on the device (ucdir_b200.utils.image.tensor2img)."""
        from ..utils.image import tensor2img
        return tensor2img(self.SR[-1], min_max=min_max)
