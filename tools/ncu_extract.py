"""Dev script: condenses an ncu report (.ncu-rep, captured with --set full --import-source on) into the tracked evidence
under profiles/: per-kernel key metrics (CSV), the SASS opcode mix and the hottest source lines.
  python tools/ncu_extract.py <report.ncu-rep> <out prefix> [kernel-name substring]"""
import csv, collections, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sass__inst_executed_global_loads", "sass__inst_executed_global_stores", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "memory_l2_theoretical_sectors_global", "memory_l2_theoretical_sectors_local", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]

def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout

rep, out = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
hdr, units = rows[0], rows[1]
with open(out + "_metrics.csv", "w", newline="") as f:
    wr = csv.writer(f); wr.writerow(["kernel", "metric", "value", "unit"])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if flt and flt not in d.get("Kernel Name", ""): continue
        for k in KEYS:
            if k in d: wr.writerow([d["Kernel Name"], k, d[k], units[hdr.index(k)]])
src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["-k", "regex:" + flt] if flt else [])))))
h = next(i for i, r in enumerate(src) if r and r[0] == "Line No")
H = src[h]; iI, iS = H.index("Instructions Executed"), H.index("# Samples")
lines, ops, cur = {}, collections.Counter(), None
for r in src[h + 1:]:
    if len(r) <= iI: continue
    if r[0] != "":
        try: lines[r[0] + " | " + r[1][:110]] = (float(r[iI]), float(r[iS]))
        except ValueError: pass
        continue
    t = r[3].split()
    if not t or t[0] == "...": continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    try: ops[op] += float(r[iI])
    except ValueError: pass
tot = sum(ops.values()) or 1.0
with open(out + "_hot.md", "w") as f:
    f.write(f"source: {rep}\n\n## SASS opcode mix (share of executed warp instructions, total {tot:.4g})\n\n")
    for k, v in ops.most_common(24): f.write(f"    {k:10s} {100 * v / tot:5.1f} %\n")
    ts = sum(v[1] for v in lines.values()) or 1.0; ti = sum(v[0] for v in lines.values()) or 1.0
    f.write("\n## hottest source lines (share of instructions attributed incl. inlining / share of stall samples)\n\n")
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:40]: f.write(f"    {100 * v[0] / ti:5.2f} % inst {100 * v[1] / ts:5.2f} % samples  {k}\n")
print("wrote", out + "_metrics.csv", out + "_hot.md")
