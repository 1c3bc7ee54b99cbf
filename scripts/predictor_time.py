"""Time the initial predictor (once per image) on both plans at the size DDPM.test gives a 1024x1024 image (1152 -> S = 1184)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ucdir_b200
from ucdir_b200.model.networks import define_G
torch.manual_seed(1234)
net = define_G({"model": ucdir_b200.SID_MODEL_OPT}).cuda().eval()
for side in (1152, 256):
    x = torch.rand(1, 3, side, side, device="cuda") * 2 - 1
    for mode in ("fp32", "tc"):
        net.predictor.engine().set_mode(mode)
        for _ in range(2): y = net.predictor(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y = net.predictor(x)
        e1.record(); torch.cuda.synchronize()
        print("predictor %dx%d %s: %.3f ms" % (side, side, mode, e0.elapsed_time(e1) / 5))
