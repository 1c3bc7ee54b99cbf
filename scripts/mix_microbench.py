"""Micro-benchmark of the integration-module conv (UCDIR_OP_TC_CONV, MODE = 1) at the C3 layer shapes, for one or more builds of the
library: python scripts/mix_microbench.py [--split] [--out FILE] LIB [LIB ...]

Every LIB (a libucdir_b200.so variant, e.g. built with UCDIR_NVCC_EXTRA=-D...) is dlopen'ed next to the others and runs the same
op record on the same device buffers: CUDA-event time over 20 launches (inputs larger than L2: 121 tiles) and the maximum
difference to torch.nn.functional in the reference's form (model/ucdir.py:112,135-140) computed on the same GPU.  Development
tool: timings of experimental epilogues / schedules before they go into the product build."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ucdir_b200 import _lib, engine as E  # noqa: E402

BF = torch.bfloat16


def build_case(C, B, H, W, split, dev):
    g = torch.Generator().manual_seed(C + H)
    rnd = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    h1 = torch.nn.functional.silu(rnd(B, H, W, C))
    w = rnd(8 * C, C // 8, 3, 3, scale=1.0 / np.sqrt(C // 8 * 9))
    bias = rnd(8 * C, scale=0.1)
    gamma, beta = 1 + 0.3 * rnd(C), 0.2 * rnd(C)
    kc, kb, nt, nsplit = E.tc_mix_tiling(C)
    wp, tb, tg = E.pack_tc_grouped(w, bias, 8, kb, gamma, beta, split=split)
    att, attw, res = rnd(B, H, W, 8), rnd(B, 8), rnd(B, H, W, C)

    def planes(t):
        if not split:
            return t.to(BF)
        hi, lo = E.split_hi_lo(t)
        return torch.cat([hi, lo], dim=-1)

    def value(t):
        return t[..., :C].float() + t[..., C:].float() if split else t.float()
    t = {"h1": planes(h1), "res": planes(res), "w": wp, "tb": tb, "tg": tg, "att": att, "attw": attw}
    h1v = value(t["h1"])
    v = h1v.double().reshape(B, -1)
    t["s0"] = torch.stack([v.sum(1), (v * v).sum(1)], dim=1)
    t["dst"] = torch.zeros(B, H, W, (2 if split else 1) * C, dtype=BF)
    t["dstats"] = torch.zeros(B, 2, dtype=torch.float64)
    t = {k: x.contiguous().to(dev) for k, x in t.items()}
    # torch reference on the GPU
    F = torch.nn.functional
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        hh = F.group_norm(h1v.to(dev).permute(0, 3, 1, 2), 1, gamma.to(dev), beta.to(dev), eps=1e-5)
        hset = F.conv2d(hh, w.to(dev), bias.to(dev), padding=1, groups=8).view(B, C, 8, H, W)
        a = att.to(dev).permute(0, 3, 1, 2) * attw.to(dev).view(B, 8, 1, 1)
        hs = torch.sum(hset * a.unsqueeze(1), dim=2)
        want = (hs * torch.sigmoid(hs)).permute(0, 2, 3, 1) + value(t["res"])
    ol = E.OpList()
    A = lambda x, st=None: E.Act(x, C, H, W, st.data_ptr() if st is not None else 0, True, split)
    E._tc_op(ol, split=1 if split else 0, src0=A(t["h1"], t["s0"]), w=t["w"].data_ptr(), tb=t["tb"].data_ptr(), tg=t["tg"].data_ptr(), gn=1,
             ncls=9, groups=8, kc=kc, kb=kb, nsplit=nsplit, nt=nt, mode=1, att=t["att"].data_ptr(), attw=t["attw"].data_ptr(), attw_stride=8,
             res=A(t["res"]), dst=A(t["dst"], t["dstats"]), ntot=8 * C, B=B, halo=1)
    return t, ol, want, value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--split", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--tiles", type=int, default=121)
    ap.add_argument("--shapes", default="", help="C:S pairs instead of the default halo shapes, e.g. 512:16,512:8 (the streamed C = 512 levels)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    shapes = [(64, 128), (128, 64)] + ([] if a.split else [(256, 32)])
    if a.shapes:
        shapes = [tuple(int(v) for v in t.split(":")) for t in a.shapes.split(",")]
    libs = []
    for path in a.libs:
        lib = ctypes.CDLL(os.path.abspath(path))
        lib.ucdir_run_ops.argtypes = [ctypes.POINTER(_lib.Op), ctypes.c_int, ctypes.c_void_p]
        lib.ucdir_run_ops.restype = ctypes.c_int
        lib.ucdir_last_error.restype = ctypes.c_char_p
        libs.append((os.path.basename(path), lib))
    rows = []
    for C, S in shapes:
        t, ol, want, value = build_case(C, a.tiles, S, S, a.split, dev)
        arr, n = ol.array(), len(ol)
        st = torch.cuda.current_stream().cuda_stream
        gflop = 2.0 * a.tiles * S * S * 8 * C * (C // 8) * 9 / 1e9
        for name, lib in libs:
            t["dst"].zero_(); t["dstats"].zero_()
            rc = lib.ucdir_run_ops(arr, n, ctypes.c_void_p(st))
            if rc:
                print(name, C, "FAILED", lib.ucdir_last_error()); continue
            torch.cuda.synchronize()
            err = (value(t["dst"]) - want).abs().max().item()
            for _ in range(3):
                lib.ucdir_run_ops(arr, n, ctypes.c_void_p(st))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.ucdir_run_ops(arr, n, ctypes.c_void_p(st))
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1000
            rows.append({"lib": name, "C": C, "S": S, "us": round(us, 1), "tflops": round(gflop / us * 1e3, 1), "max_err": err})
            print(rows[-1], flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
