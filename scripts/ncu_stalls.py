"""Summarise an ncu --set full --import-source on report: headline metrics per kernel and the SASS instructions that
collect the most warp-stall samples.  Usage: python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "launch__registers_per_thread", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("====", d["Kernel Name"][:100])
    for k in KEYS:
        if k in d:
            print("  %-75s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = {k: float(d[k].replace(",", "")) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")}
    print("  stalls/issue:", ", ".join("%s %.2f" % (k.split("stalled_")[1].split("_per_")[0], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kernels, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for k in kernels:
    h = k["hdr"]
    iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iS]) for r in k["rows"])
    print("=====", k["name"][:80], "| samples", tot, "| SASS rows", len(k["rows"]))
    for idx, r in sorted(sorted(enumerate(k["rows"]), key=lambda ir: -int(ir[1][iS]))[:top_n]):
        st = sorted(((h[i], int(r[i])) for i in stall_cols if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print("%5d %-72s %6s %9s %s" % (idx, r[iSrc].strip()[:72], r[iS], r[iEx], st))
