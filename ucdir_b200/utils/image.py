"""Caller-side tail of the path on the device: `tensor2img` (core/metrics.py:8-34) fused with the `[..., pd:-pd, pd:-pd]`
crop of `DDPM.test` (model/model.py:137), one kernel (UCDIR_OP_TO_IMAGE_U8) and one uint8 device->host copy.

Same signature and result as the reference's `tensor2img` for what `sr.py` passes it (a (1,C,H,W) or (C,H,W) RGB tensor,
`out_type=np.uint8`), bit for bit: clamp, `(x - min) / (max - min)`, `* 255.0`, round half to even, HWC uint8.  The input
must be a CUDA tensor -- like the rest of ucdir_b200 there is no CPU route (the reference calls `.cpu()` first; here the
conversion happens before the copy so one quarter of the bytes crosses PCIe)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, engine


def tensor2img(tensor: torch.Tensor, out_type=np.uint8, min_max=(-1, 1), crop: int = 0) -> np.ndarray:
    """core/metrics.py:8-34.  `crop` = pixels removed from every side first (the 64 of model/model.py:127,137)."""
    if out_type != np.uint8:
        raise NotImplementedError("ucdir_b200.tensor2img: only out_type=np.uint8 (what sr.py saves) is built")
    t = tensor.squeeze()
    if t.dim() == 2:
        t = t[None]
    if t.dim() != 3:
        raise TypeError("Only 3D (C,H,W), 2D (H,W) or batch-1 4D tensors are built; received dimension %d" % t.dim())
    engine._require_cuda(t.device)
    t = t.detach().float().contiguous()
    C, H, W = t.shape
    if C > 4:
        raise TypeError("tensor2img: at most 4 channels, got %d" % C)
    out = torch.empty(H - 2 * crop, W - 2 * crop, C, dtype=torch.uint8, device=t.device)
    ol = engine.OpList()
    ol.add("UCDIR_OP_TO_IMAGE_U8", {"UCDIR_IMG_P_SRC": t.data_ptr(), "UCDIR_IMG_P_DST": out.data_ptr()},
           {"UCDIR_IMG_I_B": 1, "UCDIR_IMG_I_C": C, "UCDIR_IMG_I_H": H, "UCDIR_IMG_I_W": W, "UCDIR_IMG_I_PD": crop},
           {"UCDIR_IMG_F_MIN": float(min_max[0]), "UCDIR_IMG_F_MAX": float(min_max[1])})
    engine._run_ops(ol.array(), len(ol), engine._stream(t.device))
    img = out.cpu().numpy()
    return img[..., 0] if tensor.squeeze().dim() == 2 else img
