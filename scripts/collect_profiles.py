#!/usr/bin/env python
"""Copy the round-end measurements from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, per-op profile, ncu launch
list (+ summary), ncu --set full summaries of the halo kernels, and the top-kernel traffic record bench.py reads."""
import csv
import gzip
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def cp(src, dst):
    s = os.path.join(G, src)
    if os.path.exists(s) and os.path.getsize(s) > 0:
        shutil.copyfile(s, os.path.join(P, dst))
        print("profiles/" + dst)


cp("bench.json", TAG + "_bench_bf16_halo.json")
cp("bench_reference.json", TAG + "_bench_reference_arm.json")
cp("bench_c2.json", TAG + "_bench_c2_bf16.json")
cp("bench_halo0.json", TAG + "_bench_bf16_streamed_only.json")
cp("ops_profile.json", TAG + "_ops_profile_bf16_halo.json")
cp("bf16_errors.json", TAG + "_bf16_errors.json")
lc = os.path.join(G, "launches.csv")
if os.path.exists(lc):
    with open(lc, "rb") as f, gzip.open(os.path.join(P, TAG + "_launches_bf16_halo.csv.gz"), "wb") as g:
        g.write(f.read())
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), lc], capture_output=True, text=True).stdout
    with open(os.path.join(P, TAG + "_launches_bf16_halo.md"), "w") as f:
        f.write("# bf16 tcgen05 path with the halo kernels: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu` "
                "(c3_1024_tile128, 1 B200), our kernels only\n\nCold-cache, serialised timings: compare shares.  Covers the predictor "
                "(fp32 SIMT kernels, once per image), the eager warm-up runs before graph capture, the graph replays and the eager "
                "per-op profiling pass.\n\n" + out)
    print("profiles/%s_launches_bf16_halo.{csv.gz,md}" % TAG)

summ = []
top = None
for rep in ("prof_halo.ncu-rep", "prof_final.ncu-rep"):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_stalls.py"), path, "22"], capture_output=True, text=True).stdout
    summ.append("## %s\n\n```\n%s```\n" % (rep, txt))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "mix_halo_kernel<8>" in d["Kernel Name"]:
            f = lambda k: float(d[k].replace(",", ""))
            top = {"source": "ncu --set full --clock-control none --import-source on -k regex:(mix_halo|dense_halo)_kernel -s 1 -c 2 python bench.py "
                             "--steps 1 --warmup 3 --no-cpu (scripts/gpu_final.sh)",
                   "kernel": d["Kernel Name"],
                   "what": "downs.1 spdyconv + integration mix, C=64, 121 tiles x 128x128 (the single most expensive launch of the step; 5 such launches per step)",
                   "gpu_time_ms": f("gpu__time_duration.sum") / 1e3,
                   "dram_bytes_read": f("dram__bytes_read.sum") * 1e6, "dram_bytes_write": f("dram__bytes_write.sum") * 1e6,
                   "algorithmic_bytes": 824705024,
                   "algorithmic_bytes_note": "per pixel: h1 in (64 bf16) + guidance map (8 fp32) + residual (64 bf16) + out (64 bf16); weights 0.15 MB",
                   "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                   "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                   "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                   "ipc": f("sm__inst_executed.avg.per_cycle_elapsed"), "registers_per_thread": d["launch__registers_per_thread"],
                   "grid": d.get("launch__grid_size"), "block": d.get("launch__block_size")}
            top["dram_bytes_per_launch"] = top["dram_bytes_read"] + top["dram_bytes_write"]
if summ:
    with open(os.path.join(P, TAG + "_ncu_halo_kernels.md"), "w") as f:
        f.write("# ncu --set full summaries of the halo kernels (headline metrics, stall reasons per issue, SASS lines with the most warp-stall samples)\n\n"
                "Produced by `python scripts/ncu_stalls.py <report>` from the captures of `scripts/gpu_final.sh`.\n\n" + "\n".join(summ))
    print("profiles/%s_ncu_halo_kernels.md" % TAG)
if top:
    json.dump(top, open(os.path.join(P, TAG + "_ncu_top_kernel.json"), "w"), indent=1)
    print("profiles/%s_ncu_top_kernel.json" % TAG, "traffic", top["dram_bytes_per_launch"])
