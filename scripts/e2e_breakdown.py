"""Where the end-to-end leg of bench.py spends its host time: wall-clock per phase of GaussianDiffusion.p_sample with host
buffers (each phase followed by a device synchronize), for the C3 and C2 workloads.  Diagnostic only."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ucdir_b200  # noqa: E402
from ucdir_b200.model.networks import define_G  # noqa: E402


def run(workload):
    dev = torch.device("cuda", 0)
    batch, side, skip, padding, force = bench.WORKLOADS[workload]
    torch.manual_seed(bench.WEIGHT_SEED)
    net = define_G({"model": ucdir_b200.SID_MODEL_OPT}).to(dev).eval()
    net.denoise_fn.engine().set_precision("bf16")
    net.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, dev)
    unet = net.denoise_fn
    unet.tile_skip, unet.tile_padding = skip, padding
    if force:
        unet.tile_trigger = 0
    x_in = bench.synth_input(batch, side).pin_memory().to(dev)
    initx = net.predictor(x_in)
    x_pin = torch.randn(x_in.shape).pin_memory()
    out_pin = torch.empty_like(x_pin).pin_memory()
    sync = torch.cuda.synchronize
    names = ["h2d", "session+bind", "table", "load_state", "noise", "step", "clone", "d2h", "whole p_sample call (no inner syncs)"]
    acc = [0.0] * len(names)
    n = 12
    for k in range(n + 3):
        t = 49 - k
        tt = [time.perf_counter()]
        xt = x_pin.to(dev, non_blocking=True); sync(); tt.append(time.perf_counter())
        sess = unet.engine().session(x_in, initx); net._sync_noise_stream(sess, dev); sync(); tt.append(time.perf_counter())
        table = net._params_table(dev, True); sync(); tt.append(time.perf_counter())
        sess.load_state(xt); sync(); tt.append(time.perf_counter())
        net._fill_noise(sess.noise); sync(); tt.append(time.perf_counter())
        sess.step_resident(table[t]); sync(); tt.append(time.perf_counter())
        out = sess.state().clone(); sync(); tt.append(time.perf_counter())
        out_pin.copy_(out, non_blocking=True); sync(); tt.append(time.perf_counter())
        xt = x_pin.to(dev, non_blocking=True)
        out = net.p_sample(xt, t, condition_x=x_in, kwargs={"guide": initx})
        out_pin.copy_(out, non_blocking=True); torch.cuda.current_stream().synchronize(); tt.append(time.perf_counter())
        if k >= 3:
            for j in range(len(names)):
                acc[j] += tt[j + 1] - tt[j]
    # the bench's situation: a long run of resident steps first (GPU at its power-capped clock), then stateless calls
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    table = net._params_table(dev, True)
    sess = unet.engine().session(x_in, initx)
    sess.load_state(torch.randn(x_in.shape, device=dev, generator=gen))
    for k in range(60):
        sess.noise.normal_(generator=gen); sess.step_resident(table[49 - k % 49])
    sync()
    series = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu_ms = []
    for k in range(20):
        t0 = time.perf_counter()
        xt = x_pin.to(dev, non_blocking=True)
        e0.record()
        out = net.p_sample(xt, 49 - k, condition_x=x_in, kwargs={"guide": initx})
        e1.record()
        out_pin.copy_(out, non_blocking=True); torch.cuda.current_stream().synchronize()
        series.append(1e3 * (time.perf_counter() - t0)); gpu_ms.append(e0.elapsed_time(e1))
    print("  after 60 resident steps: wall per call", " ".join("%.1f" % v for v in series))
    print("  device time of the p_sample part (events):", " ".join("%.1f" % v for v in gpu_ms))
    print(workload, "cpus", os.cpu_count(), "loadavg", os.getloadavg(), "threads", torch.get_num_threads())
    for nme, a in zip(names, acc):
        print("  %-45s %8.3f ms" % (nme, 1e3 * a / n))


if __name__ == "__main__":
    os.environ.setdefault("UCDIR_PRECISION", "bf16")
    for w in sys.argv[1:] or ["c3_1024_tile128", "c2_256_b8"]:
        run(w)
