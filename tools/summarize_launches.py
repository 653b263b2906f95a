#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list into a
per-kernel table of one train step.

    python tools/summarize_launches.py gpurun_out/launches.csv <launches_per_step> [traffic.json] > profiles/<name>.md

The per-launch times are cold-cache and serialised (ncu replays each kernel alone): compare SHARES, not absolutes.
With the DRAM metrics present the table also carries the measured DRAM bytes per kernel class, and `traffic.json` (if given)
receives {"<kernel class>": {"launches": n, "dram_bytes_per_launch": x}} -- what bench.py reports as roofline.traffic.
"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}


def klass(name):
    name = re.sub(r"\(.*", "", name).replace("mfp::", "").replace("void ", "")
    return name


def main():
    path, per_step = sys.argv[1], int(sys.argv[2])
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    lines = [l for l in open(path) if not l.startswith("==")]
    launches = collections.OrderedDict()  # ID -> dict
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], {"name": klass(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"], "us": 0.0, "rd": None, "wr": None})
        val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = val
        elif r["Metric Name"] == "dram__bytes_read.sum":
            d["rd"] = val
        elif r["Metric Name"] == "dram__bytes_write.sum":
            d["wr"] = val
    last = list(launches.values())[-per_step:]
    have_dram = all(d["rd"] is not None for d in last)
    agg = collections.OrderedDict()
    for d in last:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["us"]
        if have_dram:
            a[2] += d["rd"]
            a[3] += d["wr"]
    total = sum(a[1] for a in agg.values())
    if have_dram:
        print("| kernel | launches/step | us/step | share | DRAM read MB/step | DRAM write MB/step | DRAM GB/s |")
        print("|---|---:|---:|---:|---:|---:|---:|")
    else:
        print("| kernel | launches/step | us/step | share |")
        print("|---|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if have_dram:
            print("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f | %.0f |" % (n, a[0], a[1], 100 * a[1] / total, a[2] / 1e6, a[3] / 1e6, (a[2] + a[3]) / a[1] / 1e3))
        else:
            print("| `%s` | %d | %.1f | %.1f%% |" % (n, a[0], a[1], 100 * a[1] / total))
    if have_dram:
        rd, wr = sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
        print("| **total** | %d | %.1f | 100%% | %.1f | %.1f | %.0f |" % (len(last), total, rd / 1e6, wr / 1e6, (rd + wr) / total / 1e3))
    else:
        print("| **total** | %d | %.1f | 100%% |" % (len(last), total))
    print()
    print("<details><summary>every launch of the step, in order</summary>\n")
    print("| # | kernel | us | grid | block |" + (" DRAM MB |" if have_dram else ""))
    print("|---:|---|---:|---|---|" + ("---:|" if have_dram else ""))
    for i, d in enumerate(last):
        print("| %d | `%s` | %.1f | %s | %s |" % (i, d["name"], d["us"], d["grid"], d["block"]) + (" %.1f |" % ((d["rd"] + d["wr"]) / 1e6) if have_dram else ""))
    print("\n</details>")
    if out_json and have_dram:
        classes = {}
        for n, a in agg.items():
            key = "gemm" if n.startswith("gemm_tf32_tcgen05") else ("attention" if n.startswith("attention_") else n)
            c = classes.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
            c["launches"] += a[0]
            c["dram_bytes"] += a[2] + a[3]
            c["us"] += a[1]
        for c in classes.values():
            c["dram_bytes_per_launch"] = c["dram_bytes"] / c["launches"]
        json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one train step (cfg2)", "classes": classes},
                  open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
