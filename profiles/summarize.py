#!/usr/bin/env python
"""Summarises an ncu launch list (csv) and an `ncu --page raw --csv` dump into markdown.
usage: summarize.py <launches.csv> <raw.csv> <title>"""
import collections
import csv
import sys


def read_csv(path):
    lines = open(path).read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.reader(lines[i:]))


def main():
    launches, raw, title = sys.argv[1:4]
    print("# %s\n" % title)
    print("Captured under gpurun after the same command exited 0 without ncu. Per-launch times under ncu are cold-cache")
    print("and serialised: compare SHARES with bench.py's live CUDA-event shares, not absolutes.\n")
    rows = read_csv(launches)
    h = {n: i for i, n in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[h["Kernel Name"]].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[h["Metric Value"]])
    tot = sum(v[1] for v in agg.values())
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
    print("| kernel | launches | avg us | share of kernel time |\n|---|---|---|---|")
    for k, v in agg.items():
        print("| %s | %d | %.1f | %.3f |" % (k, v[0], v[1] / v[0] / 1e3, v[1] / tot))
    rows = read_csv(raw)
    hdr, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU(POPC) %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
            ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp insts")]
    print("\n## `ncu --set full --clock-control none` — one row per profiled launch\n")
    print("| kernel | grid | " + " | ".join(c[1] for c in cols) + " |\n|---|---|" + "---|" * len(cols))
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        vals = []
        for c, _ in cols:
            v, u = r[idx[c]], units[idx[c]]
            try:
                v = "%.2f" % float(v)
            except ValueError:
                pass
            vals.append(v + (" " + u if u in ("Mbyte", "Kbyte", "Gbyte", "byte") else ""))
        print("| %s | %s | " % (name, r[idx["Grid Size"]]) + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
