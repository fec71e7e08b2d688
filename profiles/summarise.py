#!/usr/bin/env python
"""Turns ncu exports into the small text summaries kept under profiles/.
   summarise.py launches <launches.csv>            -> per-kernel share of the step
   summarise.py raw <raw.csv>                      -> key metrics per profiled launch
"""
import collections, csv, sys

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':58s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:58]:58s} {a[0]:8d} {a[1]:12.1f} {a[1]/a[0]:10.1f} {a[1]/tot*100:6.1f}%")

def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")][:100])
        for w in want:
            if w in hdr:
                print(f"    {w:72s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
