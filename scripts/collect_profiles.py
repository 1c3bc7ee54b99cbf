#!/usr/bin/env python
"""Copy the round-end measurements from gpurun_out/final (scratch, produced by scripts/gpu_final.sh) into profiles/ (tracked):
bench lines of every arm / workload, per-op profiles, parity-error reports, the ncu launch list (+ per-kernel summary), ncu
--set full summaries of the top kernels, and the top-kernel DRAM-traffic record bench.py reads for `roofline.traffic`.

    python scripts/collect_profiles.py r02
"""
import csv
import gzip
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out", "final"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def cp(src, dst):
    s = os.path.join(G, src)
    if os.path.exists(s) and os.path.getsize(s) > 0:
        shutil.copyfile(s, os.path.join(P, dst))
        print("profiles/" + dst)


for src, dst in [("bench.json", "bench_c3_bf16.json"), ("bench_reference.json", "bench_reference_arm.json"),
                 ("bench_torch_gpu.json", "bench_torch_gpu_arm.json"), ("bench_fp32tc.json", "bench_c3_fp32tc.json"),
                 ("bench_c2_256_b8.json", "bench_c2_bf16.json"), ("bench_c5_sid_512_b32.json", "bench_c5_bf16.json"),
                 ("bench_c3_1152_ref_tiling.json", "bench_c3_1152_ref_tiling_bf16.json"),
                 ("bench_rank_share_16_tiles.json", "bench_rank_share_16_tiles_bf16.json"),
                 ("bench_halo0.json", "bench_c3_bf16_streamed_only.json"), ("bench_ref_tiling_flash0.json", "bench_c3_1152_ref_tiling_no_flash.json"),
                 ("ops_profile.json", "ops_profile_c3_bf16.json"), ("ops_profile_fp32tc.json", "ops_profile_c3_fp32tc.json"),
                 ("ops_profile_c3_1152_ref_tiling.json", "ops_profile_c3_1152_ref_tiling.json"),
                 ("ops_profile_rank_share_16_tiles.json", "ops_profile_rank_share_16_tiles.json"),
                 ("bf16_errors.json", "bf16_errors.json"), ("fp32_errors.json", "fp32_errors.json"), ("smoke.log", "smoke.log"),
                 ("pytest_gpu.log", "pytest_gpu.log")]:
    cp(src, TAG + "_" + dst)

lc = os.path.join(G, "launches.csv")
if os.path.exists(lc):
    with open(lc, "rb") as f, gzip.open(os.path.join(P, TAG + "_launches_bf16.csv.gz"), "wb") as g:
        g.write(f.read())
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), lc], capture_output=True, text=True).stdout
    with open(os.path.join(P, TAG + "_launches_bf16.md"), "w") as f:
        f.write("# bf16 tcgen05 path: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu --no-eager --no-parity-mode` "
                "(c3_1024_tile128, 1 B200), our kernels only\n\nCold-cache, serialised timings: compare shares.  Covers the predictor "
                "(fp32 SIMT kernels, once per image), the eager warm-up runs before graph capture, the graph replays and the eager "
                "per-op profiling pass.\n\n" + out)
    print("profiles/%s_launches_bf16.{csv.gz,md}" % TAG)

summ = []
top = None
WHAT = {"mix_halo_kernel<8, 0>": ("downs.1 spdyconv + integration mix, C=64, 121 tiles x 128x128 (the single most expensive launch of the step; 5 such launches per step)",
                               824705024, "per pixel: h1 in (64 bf16) + guidance map (8 fp32) + residual (64 bf16) + out (64 bf16); weights 0.15 MB")}
for rep in ("prof_halo.ncu-rep", "prof_tc.ncu-rep", "prof_attn.ncu-rep"):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_stalls.py"), path, "22"], capture_output=True, text=True).stdout
    summ.append("## %s\n\n```\n%s```\n" % (rep, txt))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f = lambda k: float(d[k].replace(",", ""))
        for key, (what, alg, note) in WHAT.items():
            if key in d["Kernel Name"] and top is None:
                top = {"source": "ncu --set full --clock-control none --import-source on (scripts/gpu_final.sh), build of round " + TAG,
                       "kernel": d["Kernel Name"], "what": what, "gpu_time_ms": f("gpu__time_duration.sum") / 1e3,
                       "dram_bytes_read": f("dram__bytes_read.sum") * 1e6, "dram_bytes_write": f("dram__bytes_write.sum") * 1e6,
                       "algorithmic_bytes": alg, "algorithmic_bytes_note": note,
                       "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                       "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                       "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                       "ipc": f("sm__inst_executed.avg.per_cycle_elapsed"), "registers_per_thread": d["launch__registers_per_thread"],
                       "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size")}
                top["dram_bytes_per_launch"] = top["dram_bytes_read"] + top["dram_bytes_write"]
if summ:
    with open(os.path.join(P, TAG + "_ncu_kernels.md"), "w") as f:
        f.write("# ncu --set full summaries (headline metrics, stall reasons per issue, SASS lines with the most warp-stall samples)\n\n"
                "`mix_halo_kernel` / `dense_halo_kernel` (prof_halo), the streamed `tc_conv_kernel` (prof_tc), `flash_attn_kernel` at 16 384 tokens "
                "(prof_attn).  Produced by `python scripts/ncu_stalls.py <report>` from the captures of `scripts/gpu_final.sh`.\n\n" + "\n".join(summ))
    print("profiles/%s_ncu_kernels.md" % TAG)
if top:
    json.dump(top, open(os.path.join(P, TAG + "_ncu_top_kernel.json"), "w"), indent=1)
    print("profiles/%s_ncu_top_kernel.json" % TAG, "traffic", top["dram_bytes_per_launch"])
san = os.path.join(ROOT, "gpurun_out", "sanitizer")
if os.path.isdir(san):
    lines = []
    for name in sorted(os.listdir(san)):
        if name.endswith(".log") and name != "build.log":
            txt = open(os.path.join(san, name), errors="replace").read()
            keep = [l for l in txt.splitlines() if any(k in l for k in ("ERROR SUMMARY", "RACECHECK SUMMARY", "passed", "failed", "Error:", "Hazard", "Invalid"))]
            lines.append("## %s\n\n```\n%s\n```\n" % (name, "\n".join(keep[:60])))
    if lines:
        with open(os.path.join(P, TAG + "_sanitizer.md"), "w") as f:
            f.write("# compute-sanitizer over tests/test_gpu_tc_ops.py (scripts/run_sanitizer.sh): summary lines of every tool's log\n\n" + "\n".join(lines))
        print("profiles/%s_sanitizer.md" % TAG)
