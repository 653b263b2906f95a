"""``ncu --set full`` raw-page CSV exports (profiles/raw_*/full_*.raw.csv) -> one markdown table per file with the metrics the profiling
recipe names: duration, DRAM bytes and throughput, tensor-pipe activity, warps active, registers / shared memory, L2 hit rate.
Usage: python tools/summarize_ncu_raw.py profiles/raw_r1_final/full_gemm.raw.csv [...] > profiles/<name>.md"""
import csv
import re
import sys

WANT = [  # (column title, regex over the metric name, format)
    ("µs", r"^gpu__time_duration\.sum$", lambda v, u: "%.1f" % (v / 1e3 if u.startswith("n") else v if u.startswith("u") else v * 1e3 if u.startswith("m") else v)),
    ("DRAM read MB", r"^dram__bytes_read\.sum$", None),
    ("DRAM write MB", r"^dram__bytes_write\.sum$", None),
    ("DRAM % of peak", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", lambda v, u: "%.1f" % v),
    ("L2 % of peak", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", lambda v, u: "%.1f" % v),
    ("L2 hit %", r"^lts__t_sector_hit_rate\.pct$", lambda v, u: "%.1f" % v),
    ("tensor pipe active % (elapsed)", r"sm__pipe_tensor_cycles_active(_realtime)?\.avg\.pct_of_peak_sustained_elapsed$", lambda v, u: "%.1f" % v),
    ("tensor pipe active % (SM active)", r"sm__pipe_tensor_cycles_active(_realtime)?\.avg\.pct_of_peak_sustained_active$", lambda v, u: "%.1f" % v),
    ("SM busy %", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", lambda v, u: "%.1f" % v),
    ("issue active %", r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$", lambda v, u: "%.1f" % v),
    ("warps active %", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", lambda v, u: "%.1f" % v),
    ("regs/thread", r"^launch__registers_per_thread$", lambda v, u: "%d" % v),
    ("smem/block KB", r"^launch__shared_mem_per_block_dynamic$", None),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def number(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return None


def short(name):
    m = re.match(r"(?:void )?([A-Za-z_0-9]+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        header, units, data = rows[0], rows[1], rows[2:]
        strip = [h.split(".TriageCompute.")[-1] if ".Triage" in h else h for h in header]
        cols = []
        for title, pattern, fmt in WANT:
            idx = [i for i, h in enumerate(strip) if re.search(pattern, h) and ".Triage" not in header[i]] or [i for i, h in enumerate(strip) if re.search(pattern, h)]
            cols.append((title, idx[0] if idx else None, fmt))
        print("### %s\n" % path)
        print("| # | kernel | grid × block | " + " | ".join(t for t, _, _ in cols) + " |")
        print("|---:|---|---|" + "---:|" * len(cols))
        for r in data:
            cells = []
            for title, i, fmt in cols:
                v = number(r[i]) if i is not None and i < len(r) else None
                if v is None:
                    cells.append("n/a")
                elif fmt is None:  # byte quantities
                    scale = SCALE.get(units[i].split("/")[0], 1.0)
                    cells.append("%.1f" % (v * scale / (1e3 if "KB" in title else 1e6)))
                else:
                    cells.append(fmt(v, units[i]))
            print("| %s | `%s` | %s × %s | %s |" % (r[0], short(r[4]), r[8], r[7], " | ".join(cells)))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
