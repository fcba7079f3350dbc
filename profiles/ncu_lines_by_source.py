#!/usr/bin/env python
"""Executed warp-instructions and stall samples of one captured kernel, aggregated per CUDA source line (no GPU needed):
the SASS page of the .ncu-rep gives the counts per instruction, `nvdisasm -g` of the shipped cubin gives each instruction's
source line; the two listings are matched by position.
    python profiles/ncu_lines_by_source.py <rep> <cubin> <mangled-kernel-substring> [n_top]"""
import csv, io, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
while rows and (not rows[0] or rows[0][0] != "Address"):
    rows.pop(0)
h = rows[0]
ci, cs, cx = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
inst = [(r[ci].strip(), float(r[cs] or 0), float(r[cx] or 0)) for r in rows[1:] if len(r) > cx]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
lines, cur = [], ("?", 0)
stack = ""
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(1)))
print("instructions: ncu %d, nvdisasm %d" % (len(inst), len(lines)))
n = min(len(inst), len(lines))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for k in range(n):
    a = agg[lines[k][0]]
    a[0] += inst[k][2]; a[1] += inst[k][1]; a[2] += 1
tx = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-instructions %.0f, stall samples %.0f" % (tx, ts))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("%-22s:%-5d  %5.1f%% of instructions  %5.1f%% of samples  (%d SASS)" % (key[0], key[1], 100 * a[0] / tx, 100 * a[1] / max(ts, 1), a[2]))
