"""The reference's OWN caller code on top of ucdir_b200: `model/model.py` `DDPM(opt)`
(.to(cuda) through set_device, DistributedDataParallel wrap, deepcopy for EMA, set_loss, set_new_noise_schedule, load_network of
a real `_gen_ema.pth`), then `feed_data -> test -> get_current_visuals` in a loop over images, `dpm_solver`'s direct
`denoise_fn(cat, t_input, guide=...)` call, and finally the unmodified `sr.py -p val` script end to end on a synthetic dataset.
Each runs in a subprocess (tests/ref_caller.py) with the one-line binding of INTEGRATION.md and import shims for modules this
image lacks; results are compared with the CPU oracle.  The reference tree comes from baseline/_ref/ucdir_reference (vendored by
__graft_entry__.build()) -- /root/reference does not exist on the GPU box."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import ucdir_b200
from oracle import ucdir_oracle as O
from tests import shims

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-3, 1e-4
NOISE_SEED = 77


def _need_reference():
    if shims.reference_root() is None:
        pytest.skip("reference tree not vendored (run __graft_entry__.build() where /root/reference exists)")


def _checkpoint(work, sd):
    os.makedirs(os.path.join(work, "ckpt"), exist_ok=True)
    cpu = {k: v.cpu() for k, v in sd.items()}
    torch.save(cpu, os.path.join(work, "ckpt", "I_E_gen_ema.pth"))
    torch.save(cpu, os.path.join(work, "ckpt", "I_E_gen.pth"))


def _run(mode, work, precision="fp32_tc"):
    env = dict(os.environ, UCDIR_PRECISION=precision, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(29700 + os.getpid() % 200))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_caller.py"), mode, str(work)], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-4000:]
    return p.stdout


def test_reference_ddpm_wraps_define_g(tmp_path, sid_weights):
    _need_reference()
    _, sd = sid_weights
    # a checkpoint whose weights differ from what the seeded constructor draws: load_network must take effect
    sd2 = {k: (v * 0.9 if k.endswith("final_conv.3.weight") else v.clone()) for k, v in sd.items()}
    _checkpoint(tmp_path, sd2)
    so = dict(schedule="linear", n_timestep=3, linear_start=1e-6, linear_end=0.4)
    json.dump(so, open(tmp_path / "sched.json", "w"))
    g = torch.Generator().manual_seed(12)
    imgs = torch.rand(2, 3, 72, 80, generator=g) * 2 - 1
    np.savez(tmp_path / "inputs.npz", sr=imgs.numpy())
    _run("ddpm", tmp_path)
    out = np.load(tmp_path / "ddpm_out.npz")
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    sched = O.schedule_buffers(so)
    assert np.array_equal(out["betas"], sched["betas"])
    gen = torch.Generator().manual_seed(NOISE_SEED)
    shape = (1, 3, 72 + 128, 80 + 128)
    for k in range(2):
        noises = [torch.randn(shape, generator=gen) for _ in range(3)]
        with torch.no_grad():
            want = O.ddpm_test(sd2, lay, sched, imgs[k:k + 1], noises, continous=True)
        got = torch.from_numpy(out["SR%d" % k])
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), "image %d: max abs err %.3e" % (k, (got - want).abs().max().item())
        assert np.array_equal(out["INF%d" % k], imgs[k:k + 1].numpy())
    T = torch.from_numpy
    with torch.no_grad():
        want_eps = O.unet_forward(sd2, "denoise_fn.", lay, torch.cat([T(out["dpm_cond"]), T(out["dpm_x"])], 1), T(out["dpm_t"]), T(out["dpm_guide"]))
    assert torch.allclose(T(out["dpm_eps"]), want_eps, rtol=RTOL, atol=ATOL)


def test_sr_py_val_runs_unchanged(tmp_path, sid_weights):
    """`python sr.py -p val -c config/sid.yaml -launcher pytorch -d --checkpoint ...` (README.md:55-58), its source executed unchanged as __main__:
    core/logger.py parses the yaml (val schedule: 'sid' -> T = 50, debug -> T = 10), data/ builds the PairDataset loader over
    two synthetic low-light PNG pairs, DDPM wraps define_G's network, every image goes through feed_data / test(continous=True) /
    get_current_visuals / tensor2img / save_jpg and the PSNR / SSIM summary is logged.  The final SR image of every input equals
    the oracle's within 1 grey level (fp32 tolerance before the uint8 rounding)."""
    _need_reference()
    from PIL import Image
    _, sd = sid_weights
    _checkpoint(tmp_path, sd)
    lq = tmp_path / "dataset" / "Sony" / "test_LH0.04s" / "input"
    gt = tmp_path / "dataset" / "Sony" / "test_LH0.04s" / "target"
    os.makedirs(lq); os.makedirs(gt)
    rng = np.random.RandomState(5)
    names = ["a0001", "a0002"]
    for n in names:
        scene = rng.randint(0, 256, size=(72, 80, 3)).astype(np.uint8)
        Image.fromarray(scene).save(gt / (n + ".png"))
        Image.fromarray((scene * 0.1).astype(np.uint8)).save(lq / (n + ".png"))
    log = _run("sr_py", tmp_path)
    assert "# Validation # PSNR" in log
    out = np.load(tmp_path / "sr_py_out.npz")
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    so = dict(schedule="linear", n_timestep=10, linear_start=1e-6, linear_end=0.4)      # core/logger.py:58-61 (sid: 0.4), debug: T = 10
    sched = O.schedule_buffers(so)
    gen = torch.Generator().manual_seed(NOISE_SEED)
    for n in names:
        key = [k for k in out.files if k.startswith(n) and k.endswith("_sr_png")]
        assert len(key) == 1, out.files
        lq_img = torch.from_numpy(np.asarray(Image.open(lq / (n + ".png")).convert("RGB")).copy()).permute(2, 0, 1).float() / 255.0
        x = (lq_img * 2 - 1).unsqueeze(0)               # data/util.py transform_augment(split='val', min_max=(-1, 1))
        noises = [torch.randn((1, 3, 72 + 128, 80 + 128), generator=gen) for _ in range(10)]
        with torch.no_grad():
            want = O.ddpm_test(sd, lay, sched, x, noises, continous=True)
        want_img = O.tensor2img(want[-1])
        got = out[key[0]]
        assert got.shape == want_img.shape
        diff = np.abs(got.astype(int) - want_img.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01, (diff.max(), (diff > 0).mean())
    assert len([k for k in out.files if k.endswith("_hr_png")]) == 2 and len([k for k in out.files if k.endswith("_inf_png")]) == 2
