"""N > 1 path on CPU: two `gloo` ranks shard the tiles of one image, all-gather the eps tile buffer
once per step and must both reproduce the single-process result -- for one UNet evaluation (tiler golden
vector) and for a 2-step sampler run (identical noise on every rank from the broadcast seed).
Kernels are replaced by the CPU op interpreter (tests/op_emulator.py); the sharding / collective / RNG host logic
is the code under test."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["UCDIR_ROOT"])
import torch.distributed as dist
import ucdir_b200
from ucdir_b200 import engine
from ucdir_b200.model.networks import define_G
from tests import op_emulator
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
if world > 1:
    dist.init_process_group("gloo", rank=rank, world_size=world)
engine._RUNNER = op_emulator.run_ops
engine._TEST_CPU_PLAN = True
torch.manual_seed(1234)
net = define_G({"model": ucdir_b200.SID_MODEL_OPT})
unet = net.denoise_fn
unet.tile_skip, unet.tile_padding, unet.tile_trigger = 64, 16, 0
assert unet.engine().shard_mode == "none"          # sharding is opt-in: the reference launcher runs one image per rank
unet.engine().set_shard_mode("tiles")
g = np.load(os.path.join(os.environ["UCDIR_ROOT"], "tests", "golden", "tiler.npz"))
T = lambda a: torch.from_numpy(np.asarray(a))
out = unet(T(g["x"]), T(g["level"]), T(g["guide"]))
sess = next(iter(unet.engine()._sessions.values()))
assert sess.world == world and (sess.my_tiles[1] - sess.my_tiles[0]) <= -(-9 // world), (sess.world, sess.my_tiles)
err = (out - T(g["out"])).abs().max().item()
# 2-step sampler on a 72x80 image with the tiler forced: rank-identical noise via the broadcast seed
net.set_new_noise_schedule(dict(schedule="linear", n_timestep=2, linear_start=1e-6, linear_end=0.4), torch.device("cpu"))
gen = torch.Generator().manual_seed(7)
x_in = torch.rand(1, 3, 72, 80, generator=gen) * 2 - 1
guide = torch.rand(1, 3, 72, 80, generator=gen) * 2 - 1
ng = torch.Generator().manual_seed(99)
net._noise_source = lambda shape: torch.randn(shape, generator=ng)     # injected noise: comparable across world sizes
res_a = net.p_sample_loop(x_in, True, kwargs={"guide": guide})
net._noise_source = None
torch.manual_seed(100 + rank)            # ranks deliberately start from different default-generator states
res_b = net.p_sample_loop(x_in, True, kwargs={"guide": guide})          # noise from the broadcast seed
np.save(os.environ["UCDIR_OUT"] + ".rank%d.npy" % rank, torch.stack([res_a, res_b]).numpy())
print("RANK", rank, "tiles", sess.my_tiles, "tiler_err", err)
assert err < 2e-4, err
if world > 1:
    # ranks that hold DIFFERENT images must not be stitched together: the bind-time checksum exchange raises on every rank
    bad = x_in + 0.01 * rank
    try:
        net.p_sample_loop(bad, True, kwargs={"guide": guide})
        raise SystemExit("tile sharding accepted rank-divergent inputs")
    except RuntimeError as e:
        assert "different conditioning images" in str(e), e
# batch sharding (SURVEY 8e(2)): B = 3 samples over `world` ranks, whole trajectory per rank, ONE gather at the end
unet.tile_skip, unet.tile_padding, unet.tile_trigger = 1024, 64, 1 << 30
unet.engine().set_shard_mode("batch")
gen = torch.Generator().manual_seed(11)
xb = torch.rand(3, 3, 40, 48, generator=gen) * 2 - 1
gb = torch.rand(3, 3, 40, 48, generator=gen) * 2 - 1
ng = torch.Generator().manual_seed(5)
net._noise_source = lambda shape: torch.randn(shape, generator=ng)
import torch.distributed as _d
calls = {"n": 0}
_orig = _d.all_gather_into_tensor
def _count(*a, **k):
    calls["n"] += 1
    return _orig(*a, **k)
_d.all_gather_into_tensor = _count
res_c = net.p_sample_loop(xb, True, kwargs={"guide": gb})
_d.all_gather_into_tensor = _orig
assert calls["n"] == (1 if world > 1 else 0), calls       # collective-free trajectory
np.save(os.environ["UCDIR_OUT"] + ".batch.rank%d.npy" % rank, res_c.numpy())
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


def _run(world, tmp_path, tag):
    procs = []
    port = 29500 + (os.getpid() % 1000)
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   UCDIR_ROOT=ROOT, UCDIR_OUT=str(tmp_path / tag), OMP_NUM_THREADS="4")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return [(np.load(str(tmp_path / tag) + ".rank%d.npy" % r), np.load(str(tmp_path / tag) + ".batch.rank%d.npy" % r))
            for r in range(world)]


def test_two_ranks_match_single_process(tmp_path):
    single, single_b = _run(1, tmp_path, "w1")[0]
    (r0, b0), (r1, b1) = _run(2, tmp_path, "w2")
    # batch sharding: rank 0 ran samples 0-1, rank 1 sample 2; after the single end-of-trajectory gather both hold all rows
    assert np.array_equal(b0, b1) and b0.shape == single_b.shape == (3 * 3, 3, 40, 48)
    np.testing.assert_allclose(b0, single_b, rtol=0, atol=1e-6)
    # both ranks hold the full stitched result after the all-gather and drew the same noise (broadcast seed)
    assert np.array_equal(r0, r1)
    assert np.isfinite(r0).all() and r0.shape == single.shape
    # with injected noise the 2-rank run equals the 1-rank run: per-tile arithmetic does not depend on the rank
    np.testing.assert_allclose(r0[0], single[0], rtol=0, atol=1e-6)
    # the shared-seed stream really produced noise (snapshots differ from the injected-noise run)
    assert not np.array_equal(r0[0][1:], r0[1][1:])


def test_four_ranks_with_an_idle_rank_match_single_process(tmp_path):
    """9 tiles / 3 samples over FOUR ranks: the last rank owns no tile and no sample (empty step body, gather-only participant)."""
    single, single_b = _run(1, tmp_path, "w1")[0]
    res = _run(4, tmp_path, "w4")
    for r, b in res:
        assert np.array_equal(r, res[0][0]) and np.array_equal(b, res[0][1])
    np.testing.assert_allclose(res[0][0][0], single[0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(res[0][1], single_b, rtol=0, atol=1e-6)
