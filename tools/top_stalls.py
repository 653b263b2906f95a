#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu --page source --csv` output.  usage: top_stalls.py file.csv [N]"""
import csv
import sys

lines = open(sys.argv[1]).read().split("\n")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
data = [r for r in rows[1:] if len(r) == len(hdr) and r[2].isdigit()]
seen, uniq = set(), []
for r in data:
    if r[0] in seen:
        continue
    seen.add(r[0])
    uniq.append(r)
tot = sum(int(r[2]) for r in uniq)
print("kernel:", lines[0][:100])
print("total samples", tot, "instructions", len(uniq))
first = next(i for i, h in enumerate(hdr) if h.startswith("stall_"))
for r in sorted(uniq, key=lambda r: -int(r[2]))[:n]:
    reasons = sorted([(int(r[i]), hdr[i].replace(" (Not Issued)", "*")) for i in range(first, len(hdr)) if r[i].isdigit() and int(r[i]) > 0], reverse=True)[:2]
    print("%5s %4d  %-60s %s" % (r[2], uniq.index(r), r[1].strip()[:60], reasons))
