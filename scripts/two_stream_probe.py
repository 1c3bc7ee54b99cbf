"""Probe: do two independent half-batches on two CUDA streams hide the per-launch fixed cost (pipeline fill / drain of ~135
dependent launches per step)?  16 tiles in one session vs 2 x 8 tiles in two sessions, sequential and concurrent (default), or any
image size: `python scripts/two_stream_probe.py 1024 1024` = the 121-tile step against two half images on two streams."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("UCDIR_PRECISION", "bf16")
import ucdir_b200
from ucdir_b200.engine import Session, geometry_tiled
from ucdir_b200.model.networks import define_G

dev = torch.device("cuda", 0)
torch.manual_seed(1234)
net = define_G({"model": ucdir_b200.SID_MODEL_OPT}).to(dev).eval()
net.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, dev)
unet = net.denoise_fn
unet.tile_skip, unet.tile_padding, unet.tile_trigger = 128, 16, 0
eng = unet.engine(); eng.ensure_weights()
table = net._params_table(dev)
pdl = os.environ.get("UCDIR_PDL", "1")

def make(h, w):
    x = torch.rand(1, 3, h, w, device=dev) * 2 - 1
    g = torch.rand(1, 3, h, w, device=dev) * 2 - 1
    s = Session(eng, geometry_tiled(1, h, w, 128, 16), 3)
    s.bind(x, g)
    s.load_state(torch.randn(1, 3, h, w, device=dev))
    return s

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

FH, FW = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (384, 384)     # e.g. 1024 1024: the 121-tile step
full = make(FH, FW)
print("tiles", full.geo.n_tiles)
t_full = timeit(lambda: full.step_resident(table[10]))
a, b = make(FH, FW // 2), make(FH, FW // 2)
print("half tiles", a.geo.n_tiles)
t_seq = timeit(lambda: (a.step_resident(table[10]), b.step_resident(table[10])))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def conc():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): a.step_resident(table[10])
    with torch.cuda.stream(s2): b.step_resident(table[10])
    cur.wait_stream(s1); cur.wait_stream(s2)
t_conc = timeit(conc)
print("PDL=%s: %d tiles one graph %.3f ms (%.4f ms/tile) | 2 x %d tiles sequential %.3f ms | on two streams %.3f ms (%.4f ms/tile)" % (
    pdl, full.geo.n_tiles, t_full, t_full / full.geo.n_tiles, a.geo.n_tiles, t_seq, t_conc, t_conc / (2 * a.geo.n_tiles)))
