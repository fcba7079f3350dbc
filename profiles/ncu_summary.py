#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): headline metrics from the raw page and the hottest SASS lines from the source page.
    python profiles/ncu_summary.py gpurun_out/prof_x.ncu-rep [n_top]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, n in enumerate(hdr):
    if n in want or ("issue_stalled" in n and n.endswith("per_issue_active.ratio")):
        try:
            v = float(vals[i].replace(",", ""))
        except ValueError:
            continue
        if "issue_stalled" in n and v < 0.25:
            continue
        print("%-86s %14.3f %s" % (n, v, units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
while rows and (not rows[0] or rows[0][0] != "Address"):
    rows.pop(0)
if len(rows) > 2:
    h = rows[0]
    def col(name):
        for i, n in enumerate(h):
            if n == name:
                return i
        return None
    ci, cs, cx = col("Source"), col("# Samples"), col("Instructions Executed")
    if cs is None:
        cs = col("Warp Stall Sampling (All Samples)")
    if ci is not None and cs is not None:
        tot_s = sum(float(r[cs] or 0) for r in rows[1:] if len(r) > cs and r[cs].replace('.', '', 1).isdigit())
        tot_x = sum(float(r[cx] or 0) for r in rows[1:] if cx is not None and len(r) > cx and r[cx].replace('.', '', 1).isdigit())
        byop = collections.defaultdict(lambda: [0.0, 0.0])
        lines = []
        for r in rows[1:]:
            if len(r) <= max(ci, cs):
                continue
            try:
                sv = float(r[cs] or 0); xv = float(r[cx] or 0) if cx is not None else 0
            except ValueError:
                continue
            toks = r[ci].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            op = op.split(".")[0]
            byop[op][0] += xv; byop[op][1] += sv
            lines.append((sv, xv, r[ci]))
        print("warp-instructions %.0f, stall samples %.0f" % (tot_x, tot_s))
        for op, (x, sv) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:14]:
            print("   %-10s %5.1f%% of instructions  %5.1f%% of stall samples" % (op, 100 * x / max(tot_x, 1), 100 * sv / max(tot_s, 1)))
        for sv, xv, txt in sorted(lines, key=lambda t: -t[0])[:ntop]:
            print("   top stall: %5.2f%% samples  %s" % (100 * sv / max(tot_s, 1), txt[:110]))
