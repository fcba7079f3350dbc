#!/usr/bin/env python
"""Aggregate an .ncu-rep's stall samples and executed warp-instructions per source line (needs -lineinfo, --import-source on).
    python profiles/ncu_lines.py rep.ncu-rep [n_top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or cur is None or not r[0].isdigit(): continue
    cs = hdr.index("# Samples"); cx = hdr.index("Instructions Executed")
    def num(x):
        try: return float(x)
        except ValueError: return 0.0
    k = (cur, int(r[0])); old = agg.get(k, (0.0, 0.0, r[1]))
    agg[k] = (old[0] + num(r[cs]), old[1] + num(r[cx]), r[1])
ts = sum(v[0] for v in agg.values()); tx = sum(v[1] for v in agg.values())
print("total samples %d, warp-instructions %d" % (ts, tx))
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), v in agg.items(): byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in byfile.items(): print("  %-28s samples %5.1f%%  instr %5.1f%%" % (f, 100 * v[0] / ts, 100 * v[1] / tx))
print("top lines by samples:")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("  %5.2f%% smp %5.2f%% ins  %s:%d  %s" % (100 * v[0] / ts, 100 * v[1] / tx, f, l, v[2].strip()[:90]))
if len(sys.argv) > 3:   # ranges "file:lo-hi,..." -> sums
    print("ranges:")
    for spec in sys.argv[3].split(","):
        f, r = spec.split(":"); lo, hi = map(int, r.split("-"))
        s_ = sum(v[0] for (ff, l), v in agg.items() if ff == f and lo <= l <= hi); x_ = sum(v[1] for (ff, l), v in agg.items() if ff == f and lo <= l <= hi)
        print("  %-40s samples %5.1f%%  instr %5.1f%% (%.0f per thread-tile)" % (spec, 100 * s_ / ts, 100 * x_ / tx, x_ / 4096 / 4))
