#!/usr/bin/env python
"""Turns the captures of tools/capture_profiles.sh (gpurun_out/<tag>_*.csv) into what profiles/ keeps:
  <tag>_launches.txt            per-kernel share of a bench step (ncu gpu__time_duration.sum list)
  <tag>_t1_full_summary.txt     key metrics of the --set full capture of the Tier-1 / MCT kernels
  <tag>_dwt_full_summary.txt    the same for the ten wavelet launches of one configs[1] and one configs[2] image
  t1_issue.json                 warp instructions per MQ decision of the three Tier-1 kernels   } tied to the kernel sources by
  dwt_traffic.json              DRAM bytes of the forward / inverse transform of one image      } a sha1: bench.py refuses stale ones
    python profiles/update_constants.py <tag> [gpurun_out]
"""
import contextlib
import csv
import io
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import bench
import summarise

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")


def capture(fn, path):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(path)
    return buf.getvalue()


def raw_rows(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    return hdr, rows[2:]


p = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(p):
    open(os.path.join(HERE, f"{tag}_launches.txt"), "w").write(capture(summarise.launches, p))
    print("wrote", f"{tag}_launches.txt")

p = os.path.join(src, f"{tag}_t1_raw.csv")
if os.path.exists(p):
    text = capture(summarise.raw, p)
    hdr, rows = raw_rows(p)
    ki, ii, ti = hdr.index("Kernel Name"), hdr.index("smsp__inst_executed.sum"), hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")
    log = open(os.path.join(src, f"{tag}_t1_plain.log")).read()
    decisions = int(re.search(r"decisions (\d+)", log).group(1))
    inst = {}
    for r in rows:
        name = r[ki]
        key = "decode" if ("t1_decode_kernel" in name or "t1_decode_uniform_kernel" in name) else "mq" if "t1_mq_kernel" in name else "model" if "t1_model_kernel" in name else None
        if key and key not in inst:
            inst[key] = float(r[ii].replace(",", "")) / decisions
            text += f"---- {key}: {inst[key]:.2f} warp instructions per MQ decision ({decisions} decisions), {float(r[ti]):.2f} active threads per instruction\n"
    open(os.path.join(HERE, f"{tag}_t1_full_summary.txt"), "w").write(text)
    if len(inst) == 3:
        inst = {k: round(v, 2) for k, v in inst.items()}
        inst["source"] = f"smsp__inst_executed.sum / MQ decisions of configs[1], ncu --set full capture (profiles/{tag}_t1_full_summary.txt)"
        inst["source_sha1"] = bench.source_sha1(bench.T1_SOURCES)
        json.dump(inst, open(os.path.join(HERE, "t1_issue.json"), "w"))
        print("wrote t1_issue.json", inst)

traffic, text = {}, ""
for w in ("c2", "c3"):
    p = os.path.join(src, f"{tag}_dwt_{w}_raw.csv")
    if not os.path.exists(p):
        continue
    text += f"==== {w}\n" + capture(summarise.raw, p)
    hdr, rows = raw_rows(p)
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = list(csv.reader(open(p)))[1]

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    for key, sub in ((w, "dwt_fwd"), (w + "_inverse", "dwt_inv")):
        sel = [r for r in rows if sub in r[ki]]
        if len(sel) == 5:
            traffic[key] = int(sum(to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi]) for r in sel))
if text:
    open(os.path.join(HERE, f"{tag}_dwt_full_summary.txt"), "w").write(text)
if traffic:
    traffic["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum summed over the five level launches of one image per direction, ncu --set full "
                       f"--clock-control none (profiles/{tag}_dwt_full_summary.txt); c2 = configs[1] planes (9/7), c3 = configs[2] planes (5/3)")
    traffic["source_sha1"] = bench.source_sha1(bench.DWT_SOURCES)
    json.dump(traffic, open(os.path.join(HERE, "dwt_traffic.json"), "w"))
    print("wrote dwt_traffic.json", traffic)
