#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total us, share)."""
import collections
import csv
import re
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| `%s` | %d | %.1f | %.4f |" % (k[:90], n, t, t / tot))
    print("\ntotal device time in list: %.1f us over %d launches" % (tot, sum(n for n, _ in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
