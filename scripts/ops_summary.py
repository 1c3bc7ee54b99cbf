#!/usr/bin/env python
"""Aggregate bench.py --dump-ops output by layer shape."""
import collections
import json
import sys

rows = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ops_profile.json"))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("total ms %.3f" % sum(r["ms"] for r in rows))
bykind = collections.defaultdict(float)
for r in rows:
    bykind[r["kind"]] += r["ms"]
print({k: round(v, 3) for k, v in bykind.items()})
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    if r["kind"] not in ("UCDIR_OP_TC_CONV", "UCDIR_OP_CONV_F32"):
        continue
    key = (r["H"], r["C0"] + r["C1"], r["N"], r["taps"], r["stride"], r["groups"], r.get("KC"), r.get("NT"), r["mode"])
    a = agg[key]; a[0] += 1; a[1] += r["ms"]; a[2] += r["gflop"]
print("%-55s %3s %8s %8s %8s" % ("(H,Cin,N,taps,stride,groups,KC,NT,mode)", "n", "ms", "gflop", "TF/s"))
for k, (n, ms, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-55s %3d %8.3f %8.1f %8.1f" % (str(k), n, ms, gf, gf / ms if ms else 0))
