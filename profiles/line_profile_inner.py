#!/usr/bin/env python
"""Like line_profile.py, but attributes every SASS instruction to the INNERMOST source line of its inline chain
(and, with a depth argument, to the chain's d-th frame from the inside), so that a kernel that is one big inlined
loop still shows where its warp instructions go.
usage: line_profile_inner.py src.csv file.cubin kernel_substring [topN] [units] [depth]
units: divide instruction counts by this number (e.g. MQ decisions of the run) for a per-unit column"""
import csv, re, subprocess, sys, collections
src, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
units = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
depth = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rows = list(csv.reader(open(src)))
for _i in range(2, len(rows)):
    if rows[_i] and rows[_i][0] == "Kernel Name":
        rows = rows[:_i]
        break
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = [(r[ia], int(r[ii] or 0), int(r[isamp] or 0), r[hdr.index("Source")]) for r in rows[2:] if len(r) > ii]
dis = subprocess.check_output(["nvdisasm", "-c", "-g", "--print-line-info-inline", cubin], text=True).splitlines()
lines, chain, infn, fresh = [], [], False, True
for l in dis:
    if l.startswith("\t.text.") or l.startswith(".text.") or l.startswith("\t.section\t.text."):
        infn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append(int(m.group(2)))
        continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        fresh = True
        lines.append(chain[min(depth, len(chain) - 1)] if chain else 0)
n = min(len(lines), len(sass))
agg = collections.defaultdict(lambda: [0, 0])
for k in range(n):
    a = agg[lines[k]]
    a[0] += sass[k][1]
    a[1] += sass[k][2]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"sass rows {len(sass)} disasm rows {len(lines)}; total warp-inst {tot_i} samples {tot_s}")
text = open("/root/repo/grokimagecompression_b200/csrc/" + (sys.argv[7] if len(sys.argv) > 7 else "t1_dec.cu")).read().splitlines()
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    per = f"{a[0] / units:6.2f}/unit" if units else ""
    print(f"{key:5d} inst {a[0]/tot_i*100:5.1f}% {per}  samples {a[1]/tot_s*100:5.1f}%  | {text[key-1].strip()[:110] if 0 < key <= len(text) else ''}")
