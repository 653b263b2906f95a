#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table of one train step.

    python tools/summarize_launches.py gpurun_out/launches.csv <launches_per_step> > profiles/<name>.md

The per-launch times are cold-cache and serialised (ncu replays each kernel alone): compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys


def main():
    path, per_step = sys.argv[1], int(sys.argv[2])
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    vals = [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("mfp::", "").replace("void ", ""), float(r["Metric Value"].replace(",", "")) / 1e3,
             r["Grid Size"], r["Block Size"]) for r in rows]
    last = vals[-per_step:]
    agg = collections.OrderedDict()
    for n, v, g, b in last:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches/step | us/step | share |")
    print("|---|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (n, a[0], a[1], 100 * a[1] / total))
    print("| **total** | %d | %.1f | 100%% |" % (len(last), total))
    print()
    print("<details><summary>every launch of the step, in order</summary>\n")
    print("| # | kernel | us | grid | block |")
    print("|---:|---|---:|---|---|")
    for i, (n, v, g, b) in enumerate(last):
        print("| %d | `%s` | %.1f | %s | %s |" % (i, n, v, g, b))
    print("\n</details>")


if __name__ == "__main__":
    main()
