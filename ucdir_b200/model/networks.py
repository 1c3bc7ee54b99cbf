"""Factory mirror: `define_G(opt)` with the reference's contract (model/networks.py:88-95).

Reads opt['model'][{which_model_G, unet_name, unet, diffusion_name, diffusion}] and returns an
nn.Module exposing .super_resolution / .set_new_noise_schedule / .set_loss with the reference's
state_dict key names, so `model/model.py` (DDPM) and `sr.py -p val` run unchanged.  See
INTEGRATION.md for the one-line binding a maintainer adds to the reference.
"""
import logging

logger = logging.getLogger("base")


def define_G(opt):
    model_opt = opt["model"]
    if model_opt["which_model_G"] != "ucdir":
        raise NotImplementedError("ucdir_b200.define_G: which_model_G=%r" % (model_opt["which_model_G"],))
    from . import diffusion, ucdir
    unet_cls = getattr(ucdir, model_opt["unet_name"])
    diff_cls = getattr(diffusion, model_opt["diffusion_name"])
    model = unet_cls(**model_opt["unet"])
    netG = diff_cls(model, **model_opt["diffusion"])
    logger.info("**model net G %s %s" % (model.__class__.__name__, netG.__class__.__name__))
    return netG
