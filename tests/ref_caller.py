"""Runner (subprocess of tests/test_gpu_reference_caller.py): drives the REFERENCE's own caller code against ucdir_b200.

    python tests/ref_caller.py ddpm  <workdir>     # model/model.py DDPM(opt): DDP wrap, EMA deepcopy, load_network, test()
    python tests/ref_caller.py sr_py <workdir>     # the unmodified sr.py source, `-p val`, executed as __main__

The only binding added is the one INTEGRATION.md section 1 documents: `model.networks.define_G = ucdir_b200...define_G`
(+ import shims for modules this image lacks, tests/shims).  For comparability with the CPU oracle the sampler noise is drawn
from a seeded CPU generator (the `_noise_source` test hook) -- CUDA and CPU generators cannot produce the same stream.
TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import shims  # noqa: E402

shims.install()          # at import time: DataLoader workers are spawned and re-import this file as __mp_main__
NOISE_SEED = 77


def bind():
    ref = shims.install()
    if ref is None:
        raise SystemExit("reference tree not found (need /root/reference or baseline/_ref/ucdir_reference)")
    import model.networks as refnet
    import ucdir_b200.model.networks as ours
    made = []

    def define_G(opt):
        net = ours.define_G(opt)
        gen = torch.Generator().manual_seed(NOISE_SEED)
        net._noise_source = lambda shape: torch.randn(shape, generator=gen)
        made.append(net)
        return net

    refnet.define_G = define_G                      # the one-line binding
    return ref, made


def run_ddpm(work):
    ref, made = bind()
    import yaml
    import core.logger as Logger
    import model as Model
    torch.cuda.set_device(0)
    torch.distributed.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % (29600 + os.getpid() % 300), rank=0, world_size=1)
    opt = yaml.safe_load(open(os.path.join(ref, "config", "sid.yaml")))
    opt["phase"] = "val"
    opt["path"]["resume_state"] = os.path.join(work, "ckpt", "I_E")
    opt["model"]["beta_schedule"]["val"] = json.load(open(os.path.join(work, "sched.json")))
    opt["rank"], opt["world_size"], opt["distributed"] = 0, 1, True
    opt = Logger.dict_to_nonedict(opt)
    diffusion = Model.create_model(opt)             # DDPM(opt): .to(cuda), DDP wrap, deepcopy -> EMA, set_loss, schedule, load_network
    assert isinstance(diffusion.netG, torch.nn.parallel.DistributedDataParallel)
    diffusion.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], schedule_phase="val")
    data = np.load(os.path.join(work, "inputs.npz"))
    out = {}
    for k in range(data["sr"].shape[0]):            # a validation loop over same-shape images (sr.py:518-525)
        sr = torch.from_numpy(data["sr"][k:k + 1])
        diffusion.feed_data({"SR": sr.clone(), "HR": sr.clone(), "Index": torch.tensor([k])})
        diffusion.test(continous=True)
        vis = diffusion.get_current_visuals()
        out["SR%d" % k] = vis["SR"].numpy()
        out["INF%d" % k] = vis["INF"].numpy()
        out["initx%d" % k] = diffusion.netG.module.pre_initx.detach().float().cpu().numpy()
    # dpm_solver's contract (sr.py:203-205): the wrapper calls denoise_fn(cat[cond, x], t_input, guide=...) directly
    net = diffusion.netG.module
    g = torch.Generator().manual_seed(3)
    cond = torch.rand(2, 3, 40, 48, generator=g) * 2 - 1
    x = torch.randn(2, 3, 40, 48, generator=g)
    t_in = torch.tensor([[0.3], [0.8]])
    guide = torch.rand(2, 3, 40, 48, generator=g) * 2 - 1
    eps = net.denoise_fn(torch.cat([cond, x], dim=1).cuda(), t_in.cuda(), guide=guide.cuda())
    out.update(dpm_cond=cond.numpy(), dpm_x=x.numpy(), dpm_t=t_in.numpy(), dpm_guide=guide.numpy(), dpm_eps=eps.cpu().numpy())
    out["betas"] = net.betas.cpu().numpy()
    np.savez(os.path.join(work, "ddpm_out.npz"), **out)
    torch.distributed.destroy_process_group()
    print("ddpm ok")


def run_sr_py(work):
    ref, made = bind()
    import core.metrics as Metrics
    saved = {}
    orig = Metrics.save_jpg

    def save_jpg(img, path, *a, **k):               # keep the uint8 arrays the script writes as JPEG
        saved[os.path.basename(path)] = np.asarray(img).copy()
        return orig(img, path, *a, **k)

    Metrics.save_jpg = save_jpg
    os.chdir(work)
    sys.argv = ["sr.py", "-p", "val", "-c", os.path.join(ref, "config", "sid.yaml"), "-launcher", "pytorch", "-d",
                "--checkpoint", os.path.join(work, "ckpt", "I_E")]
    # The script's source is executed unchanged in THIS process's __main__ namespace rather than through runpy.run_path: runpy
    # swaps sys.modules['__main__'] for a temporary module whose __file__ is sr.py, and the DataLoader workers sr.py spawns
    # (utils/dist_utils.py:11 sets the 'spawn' start method) would then re-import sr.py -- without the import shims.
    path = os.path.join(ref, "sr.py")
    try:
        exec(compile(open(path).read(), path, "exec"), sys.modules["__main__"].__dict__)
    finally:
        np.savez(os.path.join(work, "sr_py_out.npz"), **{k.replace(".", "_"): v for k, v in saved.items()})
    print("sr.py ok", sorted(saved))


if __name__ == "__main__":
    {"ddpm": run_ddpm, "sr_py": run_sr_py}[sys.argv[1]](sys.argv[2])
