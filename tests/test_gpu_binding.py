"""Binding regressions on the real CUDA caching allocator: a per-image validation loop
(sr.py:518-525 -> DDPM.test, model/model.py:124-138) over DIFFERENT same-shape images, each image's tensors freed before the
next one is created so that the allocator recycles their blocks.  Every image must equal a fresh-session run and the
oracle; round 1 returned the first image's guidance for all of them."""
import numpy as np
import pytest
import torch

import ucdir_b200
from oracle import ucdir_oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4


def _net(name, sid_weights):
    from ucdir_b200.model.networks import define_G
    _, sd = sid_weights
    torch.manual_seed(1234)
    net = define_G({"model": dict(ucdir_b200.SID_MODEL_OPT, diffusion_name=name)})
    assert all(torch.equal(v, sd[k]) for k, v in net.state_dict().items())
    net = net.to("cuda")
    net.denoise_fn.engine().set_precision("fp32_tc")          # the library's default mode (tests/conftest.py pins the SIMT one otherwise)
    return net


def _images(n, h=40, w=48, seed=31):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(1, 3, h, w, generator=g) * 2 - 1 for _ in range(n)], torch.randn(1, 3, h, w, generator=g)


@pytest.mark.parametrize("variant", ["ResiGaussianGuideDY", "ResiGaussianGuideDY_de", "ResiGaussianGuideDY_initxloss"])
def test_validation_loop_over_same_shape_images(sid_weights, variant):
    _, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    so = dict(schedule="linear", n_timestep=2, linear_start=1e-6, linear_end=0.4)
    net = _net(variant, sid_weights)
    net.set_new_noise_schedule(so, torch.device("cuda"))
    imgs, z = _images(3)
    noises = [z, z * 0.5]
    outs, ptrs = [], []
    for img in imgs:
        it = iter(noises)
        net._noise_source = lambda shape: next(it)
        x = img.cuda()
        ptrs.append(x.data_ptr())
        out = net.super_resolution(x, False)
        outs.append(out.cpu())
        del x, out                                     # blocks go back to the caching allocator before the next image
        if hasattr(net, "pre_initx"):
            del net.pre_initx
    assert len(set(ptrs)) < len(ptrs), "allocator did not recycle the input block; the test would not exercise the bug"
    for k, img in enumerate(imgs):
        with torch.no_grad():
            want, _ = O.super_resolution(sd, lay, O.schedule_buffers(so), img, noises, continous=False,
                                         guide_from="input" if variant.endswith("_de") else "initx")
        err = (outs[k] - want).abs()
        assert torch.allclose(outs[k], want, rtol=RTOL, atol=ATOL), "image %d: max abs err %.3e" % (k, err.max().item())
        # and equal to a run on a brand-new module (no session state at all)
        fresh = _net(variant, sid_weights)
        fresh.set_new_noise_schedule(so, torch.device("cuda"))
        it = iter(noises)
        fresh._noise_source = lambda shape: next(it)
        assert torch.equal(fresh.super_resolution(img.cuda(), False).cpu(), outs[k])
    assert not torch.allclose(outs[0], outs[1]) and not torch.allclose(outs[1], outs[2])


def test_ddpm_inference_loop_over_images(sid_weights, monkeypatch):
    """The same loop through the caller-side mirror (DDPMInference.feed_data / test / get_current_visuals): F.pad creates
    a fresh padded tensor per image, at the address of the previous one."""
    from ucdir_b200.model import model as M
    _, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    net = _net("ResiGaussianGuideDY", sid_weights)
    monkeypatch.setattr(M.networks, "define_G", lambda opt: net)
    dd = M.DDPMInference({"model": ucdir_b200.SID_MODEL_OPT}, device="cuda")
    so = dict(schedule="linear", n_timestep=1, linear_start=1e-6, linear_end=0.4)
    dd.set_new_noise_schedule(so, schedule_phase="val")
    imgs, _ = _images(3, 72, 80, seed=4)
    noise = torch.randn(1, 3, 72 + 128, 80 + 128, generator=torch.Generator().manual_seed(9))
    net._noise_source = lambda shape: noise
    res = []
    for img in imgs:
        dd.feed_data({"SR": img.clone(), "Index": 0})
        dd.test(continous=False)
        res.append(dd.get_current_visuals()["SR"])
        dd.data, dd.SR = None, None
    for k, img in enumerate(imgs):
        with torch.no_grad():
            want = O.ddpm_test(sd, lay, O.schedule_buffers(so), img, [noise], continous=False)
        assert torch.allclose(res[k], want, rtol=RTOL, atol=ATOL), "image %d: %.3e" % (k, (res[k] - want).abs().max().item())


def test_reference_sampler_loop_with_fresh_cat_tensors(sid_weights):
    """model/diffusion.py:166 / sr.py:203-205: `denoise_fn(torch.cat([cond, x_t], 1), level, guide=initx)` with a new cat
    tensor per step.  Three steps with different x_t: each eps equals the oracle's; the guidance maps are built once."""
    _, sd = sid_weights
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    net = _net("ResiGaussianGuideDY", sid_weights)
    g = torch.Generator().manual_seed(5)
    cond = torch.rand(1, 3, 40, 48, generator=g) * 2 - 1
    guide = torch.rand(1, 3, 40, 48, generator=g) * 2 - 1
    lvl = torch.full((1, 1), 0.7)
    cond_d, guide_d, lvl_d = cond.cuda(), guide.cuda(), lvl.cuda()
    ptrs = []
    for k in range(3):
        xt = torch.randn(1, 3, 40, 48, generator=g)
        x6 = torch.cat([cond_d, xt.cuda()], 1)
        ptrs.append(x6.data_ptr())
        eps = net.denoise_fn(x6, lvl_d, guide_d).cpu()
        del x6
        with torch.no_grad():
            want = O.unet_forward(sd, "denoise_fn.", lay, torch.cat([cond, xt], 1), lvl, guide)
        assert torch.allclose(eps, want, rtol=RTOL, atol=ATOL), "step %d: %.3e" % (k, (eps - want).abs().max().item())
    sess = next(iter(net.denoise_fn.engine()._sessions.values()))
    assert sess.n_binds == 1
