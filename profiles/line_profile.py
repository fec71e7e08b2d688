#!/usr/bin/env python
"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv --kernel-name regex:K) with the
line table of the cubin (nvdisasm -g --print-line-info-inline) to get executed warp instructions and stall
samples per CUDA source line.   usage: line_profile.py src.csv file.cubin kernel_substring [topN]"""
import csv, re, subprocess, sys, collections
src, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src)))
for _i in range(2, len(rows)):
    if rows[_i] and rows[_i][0] == "Kernel Name":
        rows = rows[:_i]
        break
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(r[ia], int(r[ii] or 0), int(r[isamp] or 0), r[hdr.index("Source")]) for r in rows[2:] if len(r) > ii]
dis = subprocess.check_output(["nvdisasm", "-c", "-g", "--print-line-info-inline", cubin], text=True).splitlines()
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith("\t.text.") or l.startswith(".text."):
        infn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
n = min(len(lines), len(sass))
agg = collections.defaultdict(lambda: [0, 0])
for k in range(n):
    a = agg[lines[k]]
    a[0] += sass[k][1]
    a[1] += sass[k][2]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"sass rows {len(sass)} disasm rows {len(lines)}; total warp-inst {tot_i} samples {tot_s}")
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{str(key):40s} inst {a[0]:>13d} {a[0]/tot_i*100:5.1f}%   samples {a[1]:>8d} {a[1]/tot_s*100:5.1f}%")
