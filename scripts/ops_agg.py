"""Aggregate a --dump-ops per-op profile by layer shape: python scripts/ops_agg.py FILE [N]"""
import json, sys
from collections import defaultdict
d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(len(d), "ops", round(sum(x.get("ms", 0) for x in d), 3), "ms")
agg = defaultdict(lambda: [0, 0.0, 0.0])
for x in d:
    k = (x["kind"],) if x["kind"] != "UCDIR_OP_TC_CONV" else (x["H"], x["C0"], x["C1"], x["N"], x["taps"], x["groups"], x["stride"], x["NT"])
    agg[k][0] += 1; agg[k][1] += x["ms"]; agg[k][2] += x.get("gflop", 0)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(k, v[0], round(v[1], 4), "ms; per launch", round(v[1] / v[0] * 1000, 1), "us; TF", round(v[2] / v[1]) if v[1] else 0)
