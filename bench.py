#!/usr/bin/env python
"""bench.py -- denoising steps/s of UCDIR's iterative-denoising hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)

A "step" is one full-image p_sample (model/diffusion.py:160-183 of the reference): all tiles' UNet forwards,
the tile stitch and the posterior update, plus (N > 1) the per-step tile all-gather.  Workload = BASELINE
config C3/C4: 1x3x1024x1024 synthetic low-light image, inter-step patch-splitting with 128-px tiles
(skip, padding) = (128, 16) -> 121 tiles, sid val schedule (T = 50, linear 1e-6 -> 0.4), seeded random weights of
config/sid.yaml's architecture.  See DESIGN.md "Measurement" for every definition used in the JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEIGHT_SEED, INPUT_SEED, NOISE_SEED = 1234, 0, 42
WORKLOADS = {
    # name: (batch, side, tile skip, tile padding, force tiler)
    "c3_1024_tile128": (1, 1024, 128, 16, True),
    "c2_256_b8": (8, 256, 1024, 64, False),
    "c1_128": (1, 128, 1024, 64, False),
    # what `sr.py -p val` runs for a 1024x1024 image: DDPM.test pads to 1152x1152 -> reference-default tiler (1024, 64)
    "c3_1152_ref_tiling": (1, 1152, 1024, 64, False),
    # 16 tiles of 128x128 = the per-rank share of the headline workload at 8 GPUs (for per-launch overhead studies)
    "rank_share_16_tiles": (1, 384, 128, 16, True),
}


def synth_input(batch, side, seed=INPUT_SEED):
    """low-passed uniform 'scene' at 10 % exposure plus sensor noise, mapped to [-1, 1]."""
    g = torch.Generator().manual_seed(seed)
    low = torch.nn.functional.interpolate(torch.rand(batch, 3, max(side // 16, 4), max(side // 16, 4), generator=g),
                                          size=(side, side), mode="bilinear", align_corners=False)
    x = (low * 0.1 + 0.05 * torch.randn(batch, 3, side, side, generator=g)).clamp(0, 1) * 2 - 1
    return x.contiguous()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# algorithmic work of an op array (DESIGN.md "Roofline"): 2*MAC for conv / GEMM ops
# ----------------------------------------------------------------------------------------------
def unet_algorithmic_flops(unet, H, W):
    """2*MAC of one DY3h.naiveforward on an HxW sample, from the reference's layer shapes (model/ucdir.py:205-293):
    returns (conv_flops, attention_core_flops).  Independent of how a precision mode executes the layers
    (zero-padded K chunks, phase-decomposed upsampling and padded output columns do not count)."""
    from ucdir_b200.model import ucdir as U
    conv = attn = 0.0
    feats = []

    def block(layer, cin, h, w):
        nonlocal conv, attn
        rb = layer.res_block
        co = rb.dim_out
        conv += 2.0 * h * w * 9 * cin * co                       # conv1
        conv += 2.0 * h * w * 9 * (co // rb.nset) * co * rb.nset  # spdyconv (grouped, C -> 8C)
        if cin != co:
            conv += 2.0 * h * w * cin * co                       # res_conv 1x1
        if layer.with_attn:
            n = h * w
            conv += 2.0 * n * co * 3 * co + 2.0 * n * co * co    # qkv, out 1x1
            attn += 4.0 * n * n * co                             # QK^T and PV
        return co

    c, h, w = None, H, W
    for layer in unet.downs:
        if isinstance(layer, torch.nn.Conv2d):
            conv += 2.0 * h * w * 9 * layer.in_channels * layer.out_channels
            c = layer.out_channels
        elif isinstance(layer, U.Downsample):
            h, w = h // 2, w // 2
            conv += 2.0 * h * w * 9 * c * c
        else:
            c = block(layer, c, h, w)
        feats.append(c)
    for layer in unet.mid:
        c = block(layer, c, h, w)
    for layer in unet.ups:
        if isinstance(layer, U.Upsample):
            h, w = h * 2, w * 2
            conv += 2.0 * h * w * 9 * c * c
        else:
            c = block(layer, c + feats.pop(), h, w)
    conv += 2.0 * h * w * 9 * c * unet.cfg["out_channel"]
    return conv, attn


def dump_op_profile(path, step_ops, prof, steps, K):
    """Per-op device time (CUDA events, averaged over the timed steps) with the shape fields of each record."""
    per = {}
    n_ops = len(step_ops)
    for ms, idx, kind in prof:
        if idx < n_ops and int(step_ops[idx].kind) == kind:
            per[idx] = per.get(idx, 0.0) + ms
    names = {v: k for k, v in K.items() if k.startswith("UCDIR_OP_") and not k.endswith(("NPTR", "NINT", "NFLT"))}
    rows = []
    for idx, o in enumerate(step_ops):
        kind = int(o.kind)
        r = {"op": idx, "kind": names.get(kind, str(kind)), "ms": round(per.get(idx, 0.0) / steps, 4)}
        if kind == K["UCDIR_OP_TC_CONV"]:
            g = lambda n: int(o.i[K["UCDIR_TC_I_" + n]])
            r.update(B=g("B"), H=g("H"), W=g("W"), C0=g("C0"), C1=g("C1"), N=g("NTOT"), taps=g("NTY") * g("NTX"), stride=g("STRIDE"),
                     groups=g("GROUPS"), KC=g("KC"), NT=g("NT"), mode=g("MODE"), gn=g("GN"))
            cin = (g("C0") + g("C1")) // g("GROUPS")
            r["gflop"] = round(2e-9 * g("B") * g("H") * g("W") * g("NTY") * g("NTX") * cin * g("NTOT"), 2)
        elif kind == K["UCDIR_OP_CONV_F32"]:
            g = lambda n: int(o.i[K["UCDIR_CONV_I_" + n]])
            r.update(B=g("B"), H=g("H"), W=g("W"), C0=g("C0"), C1=g("C1"), N=g("COUT"), taps=g("KSIZE") ** 2, stride=g("STRIDE"),
                     groups=g("GROUPS"), mode=g("MODE"))
            r["gflop"] = round(2e-9 * g("B") * g("H") * g("W") * g("KSIZE") ** 2 * ((g("C0") + g("C1")) // g("GROUPS")) * g("COUT"), 2)
        elif kind == K["UCDIR_OP_SGEMM_F32"]:
            g = lambda n: int(o.i[K["UCDIR_SGEMM_I_" + n]])
            r.update(B=g("BATCH"), M=g("M"), N=g("N"), K=g("K"))
            r["gflop"] = round(2e-9 * g("BATCH") * g("M") * g("N") * g("K"), 2)
        if r.get("gflop") and r["ms"] > 0:
            r["tflops"] = round(r["gflop"] / r["ms"], 1)
        rows.append(r)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    json.dump(rows, open(path, "w"), indent=0)


def run_ours(args):
    import ucdir_b200
    from ucdir_b200 import _lib
    from ucdir_b200.model.networks import define_G
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one process per GPU); see the module docstring" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    _lib.load()
    K = _lib.C
    batch, side, skip, padding, force = WORKLOADS[args.workload]
    torch.manual_seed(WEIGHT_SEED)
    net = define_G({"model": ucdir_b200.SID_MODEL_OPT}).to(dev).eval()
    net.set_new_noise_schedule(ucdir_b200.SID_VAL_SCHEDULE, dev)
    unet = net.denoise_fn
    unet.tile_skip, unet.tile_padding = skip, padding
    if force:
        unet.tile_trigger = 0
    Tn = net.num_timesteps
    x_host = synth_input(batch, side).pin_memory()
    x_in = x_host.to(dev, non_blocking=True)
    torch.manual_seed(NOISE_SEED)
    initx = net.predictor(x_in)                                # once per image (model/diffusion.py:475), not a step
    sess = unet.engine().session(x_in, initx)
    gen = torch.Generator(device=dev); gen.manual_seed(NOISE_SEED)
    table = net._params_table(dev)
    sess.load_state(torch.randn(x_in.shape, device=dev, generator=gen))        # resident state + CUDA graphs

    def one_step(k):
        """One p_sample exactly as GaussianDiffusion.p_sample_loop issues it: draw z_t, replay the step graph."""
        t = (Tn - 1 - k) % Tn
        if t > 0:
            sess.noise.normal_(generator=gen)
        sess.step_resident(table[t])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        one_step(k)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = _lib.launch_count()
    sess.time_collective = world > 1
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        one_step(args.warmup + k)
    ev1.record()
    barrier()
    sess.time_collective = False
    coll_ms = sum(a.elapsed_time(b) for a, b in sess.collective_events) / max(args.steps, 1)
    launches = _lib.launch_count() - launches0
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms_total, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    clk = clocks.stop() if rank == 0 else None

    # ---- per-kernel split of the same K steps: identical ops launched one by one (no graph) with a CUDA event
    # after every launch on the launching stream; used for the roofline line only, never for `value` ----
    img = sess.state().clone(); nxt = torch.empty_like(img)
    _lib.profile_begin()
    for k in range(args.steps):
        t = (Tn - 1 - (args.warmup + k)) % Tn
        noise = torch.randn(x_in.shape, device=dev, generator=gen) if t > 0 else None
        sess.step(img, nxt, net.noise_level(t), net._step_scalars(t), noise, True)
        img, nxt = nxt, img
    barrier()
    prof = _lib.profile_end()

    # ---- end-to-end leg: same steps through the public module API with HOST buffers every step ----
    x_pin = img.detach().cpu().pin_memory()
    out_pin = torch.empty_like(x_pin).pin_memory()
    cond_dev, guide = x_in, initx
    e_steps = max(3, args.steps)

    def e2e_step(k):
        t = (Tn - 1 - k) % Tn
        xt = x_pin.to(dev, non_blocking=True)                  # H2D of the step's state from pinned memory
        out = net.p_sample(xt, t, condition_x=cond_dev, kwargs={"guide": guide})
        out_pin.copy_(out, non_blocking=True)                  # D2H of the step's result
        torch.cuda.current_stream().synchronize()

    for k in range(3):
        e2e_step(k)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(e_steps):
        e2e_step(k)
    t1.record()
    barrier()
    e_ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(e_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_value = e_steps / (float(e_ms.item()) / 1e3)

    # ---- roofline of the dominant kernel class (convolution implicit GEMMs), from the live per-op events ----
    step_ops = sess.step_ops.ops
    conv_kinds = {K["UCDIR_OP_CONV_F32"], K["UCDIR_OP_TC_CONV"]}
    by_kind_ms = {}
    for ms, idx, kind in prof:
        by_kind_ms[kind] = by_kind_ms.get(kind, 0.0) + ms
    my_tiles = sess.my_tiles[1] - sess.my_tiles[0]
    conv_tile, attn_tile = unet_algorithmic_flops(unet, sess.geo.TH, sess.geo.TW)
    conv_flops, all_flops = conv_tile * my_tiles, (conv_tile + attn_tile) * my_tiles
    conv_ms = sum(v for k, v in by_kind_ms.items() if k in conv_kinds) / args.steps
    n_conv = sum(1 for o in step_ops if int(o.kind) in conv_kinds)
    if args.dump_ops and rank == 0:
        dump_op_profile(args.dump_ops, step_ops, prof, args.steps, K)
    peaks = load_peaks()
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    names = {v: k for k, v in K.items() if k.startswith("UCDIR_OP_") and k not in ("UCDIR_OP_NPTR", "UCDIR_OP_NINT", "UCDIR_OP_NFLT")}
    share = {names.get(k, str(k)): round(v / args.steps, 4) for k, v in sorted(by_kind_ms.items(), key=lambda kv: -kv[1])}
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_top_kernel.json")
    if args.precision == "bf16" and os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = int(tj["dram_bytes_per_launch"])
        traffic_note = "dram read+write of the most expensive launch (%s; %s), ncu --set full; algorithmic bytes of that launch %d" % (
            tj["kernel"].split("(")[0], tj["what"].split(",")[0], tj["algorithmic_bytes"])
    roofline = {"bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["bf16_sustained"], 4), "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "conv implicit-GEMM family (%s), %d launches/step on this rank" % (
                    "fp32 SIMT conv_f32_kernel" if args.precision == "fp32" else "tcgen05 bf16", n_conv),
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "algorithmic_tflop_per_step_this_rank": round(all_flops / 1e12, 4),
                "conv_ms_per_step": round(conv_ms, 3), "ms_per_step_by_op_kind": share}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    cpu = cpu_baseline(args, n_tiles=sess.geo.n_tiles, sample_tiles=args.cpu_tiles) if world == 1 and not args.no_cpu else None
    ms_per_step = ms_total / args.steps
    geo = sess.geo
    line = {
        "metric": "denoising steps/sec (1024x1024, 50-step sampler)" if args.workload.startswith("c3") else "denoising steps/sec",
        "value": round(args.steps / (ms_total / 1e3), 4), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": args.workload, "image": [batch, 3, side, side], "tiling": "inter-step patch-split"
                   if geo.kind == "tiled" else geo.kind, "tile": [geo.TH, geo.TW], "tile_padding": padding if geo.kind == "tiled" else 0,
                   "tiles_per_step": geo.n_tiles, "tiles_this_rank": sess.my_tiles[1] - sess.my_tiles[0],
                   "schedule": "linear T=50 1e-6->0.4 (sid val)", "sampler": "ancestral p_sample", "weights": "seeded random (sid.yaml arch)",
                   "parallelism": "tiles sharded over %d rank(s), 1 NCCL all-gather/step" % world if world > 1 else "1 GPU",
                   "l2": "no flush: per-step activation working set (%.1f GB) >> 126 MB L2" % (sess.pool.total_bytes() / 1e9)},
        "unet_ms_per_step": round(sum(v for k, v in by_kind_ms.items() if k != K["UCDIR_OP_SCATTER"]) / args.steps, 3),
        "e2e": {"value": round(e2e_value, 4), "unit": "steps/s", "h2d_bytes_per_step": x_pin.numel() * 4,
                "d2h_bytes_per_step": out_pin.numel() * 4, "steps": e_steps,
                "api": "GaussianDiffusion.p_sample(x_host->dev, t, condition_x, guide) -> host"},
        "gpu_launches": int(launches),
        "collective": None if world == 1 else {"kind": "NCCL all_gather_into_tensor of tile interiors, 1 per step",
                                               "bytes_total_per_step": int(sess.eps.numel() * 4), "ms_per_step": round(coll_ms, 3)},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clk,
    }
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# CPU legs: the reference algorithm (oracle port, torch CPU fp32) on the host cores
# ----------------------------------------------------------------------------------------------
def _cpu_setup(workload):
    import ucdir_b200
    from oracle import ucdir_oracle as O
    from ucdir_b200.model.networks import define_G
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(WEIGHT_SEED)
    net = define_G({"model": ucdir_b200.SID_MODEL_OPT})
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    lay = O.UNetLayout(**ucdir_b200.SID_MODEL_OPT["unet"])
    sched = O.schedule_buffers(ucdir_b200.SID_VAL_SCHEDULE)
    return O, sd, lay, sched


def _cpu_tile_step(O, sd, lay, sched, tiles_x6, guide_tiles, t):
    """Sequential B=1 tile forwards, as utils/util.py:124-145 runs them, at the level of step t."""
    lvl = torch.full((1, 1), float(np.float32(sched["sqrt_alphas_cumprod_prev_f64"][t + 1])))
    with torch.no_grad():
        for k in range(tiles_x6.shape[0]):
            O.unet_naiveforward(sd, "denoise_fn.", lay, tiles_x6[k:k + 1], lvl, guide_tiles[k:k + 1])


def cpu_baseline(args, n_tiles, sample_tiles):
    batch, side, skip, padding, force = WORKLOADS[args.workload]
    O, sd, lay, sched = _cpu_setup(args.workload)
    g = torch.Generator().manual_seed(INPUT_SEED)
    ts = skip if force else (side // 32 + 1) * 32
    x6 = torch.rand(sample_tiles, 6, ts, ts, generator=g) * 2 - 1
    gd = torch.rand(sample_tiles, 3, ts, ts, generator=g) * 2 - 1
    _cpu_tile_step(O, sd, lay, sched, x6[:1], gd[:1], 10)          # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    _cpu_tile_step(O, sd, lay, sched, x6, gd, 10)
    dt = time.perf_counter() - t0
    per_tile = dt / sample_tiles
    return {"value": round(1.0 / (per_tile * n_tiles), 6), "unit": "steps/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": "%d of %d tile forwards (1x6x%dx%d, oracle/ucdir_oracle.py on torch CPU fp32) timed in "
            "%.1f s; steps/s extrapolated = 1 / (s_per_tile * %d); posterior update excluded (<0.1%%)" % (
                sample_tiles, n_tiles, ts, ts, dt, n_tiles)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, side, skip, padding, force = WORKLOADS[args.workload]
    O, sd, lay, sched = _cpu_setup(args.workload)
    from ucdir_b200.engine import geometry_tiled, geometry_direct
    geo = geometry_tiled(batch, side, side, skip, padding) if force else geometry_direct(batch, side, side)
    n_tiles = geo.n_tiles
    sample = args.cpu_tiles
    g = torch.Generator().manual_seed(INPUT_SEED)
    x6 = torch.rand(sample, 6, geo.TH, geo.TW, generator=g) * 2 - 1
    gd = torch.rand(sample, 3, geo.TH, geo.TW, generator=g) * 2 - 1
    for k in range(args.warmup):
        _cpu_tile_step(O, sd, lay, sched, x6, gd, 49 - k % 50)
    t0 = time.perf_counter()
    for k in range(args.steps):
        _cpu_tile_step(O, sd, lay, sched, x6, gd, 49 - (args.warmup + k) % 50)
    dt = time.perf_counter() - t0
    per_tile = dt / (args.steps * sample)
    value = 1.0 / (per_tile * n_tiles)
    cores = torch.get_num_threads()
    desc = "each step = %d of %d sequential B=1 tile forwards (%dx%d) of the reference algorithm (oracle port, torch CPU " \
           "fp32, %d threads); steps/s = 1 / (s_per_tile * %d)" % (sample, n_tiles, geo.TH, geo.TW, cores, n_tiles)
    line = {"metric": "denoising steps/sec (1024x1024, 50-step sampler)" if args.workload.startswith("c3") else "denoising steps/sec",
            "value": round(value, 6), "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 / value, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload, "image": [batch, 3, side, side], "tile": [geo.TH, geo.TW],
                       "tile_padding": padding if force else 0, "tiles_per_step": n_tiles,
                       "schedule": "linear T=50 1e-6->0.4 (sid val)", "weights": "seeded random (sid.yaml arch)"},
            "cpu_baseline": {"value": round(value, 6), "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": round(value, 6), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_1024_tile128", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("UCDIR_PRECISION", "bf16"), choices=["fp32", "bf16"],
                    help="bf16 = tcgen05 tensor-core path (default, stated tolerance); fp32 = SIMT parity path (rtol 1e-3 / atol 1e-4)")
    ap.add_argument("--cpu-tiles", type=int, default=4, help="tile forwards per CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--dump-ops", default="", help="write the per-op device-time profile of the timed steps to this JSON file")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    os.environ["UCDIR_PRECISION"] = args.precision
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
