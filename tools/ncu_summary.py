"""Text summary of an .ncu-rep (one block per kernel launch): python tools/ncu_summary.py <report> [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = [f"# {rep} -- ncu --set full --clock-control none (times are cold-cache, serialised)"]
for r in rows[2:]:
    lines.append("")
    lines.append(f"== {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
    for w in want:
        if w in idx and r[idx[w]] != "":
            lines.append(f"   {w:95s} {r[idx[w]]:>18s} {units[idx[w]]}")
    try:
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")); wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
        lines.append(f"   dram traffic per launch (read + write)                                                          {tot / 1e9:18.4f} GB")
    except Exception as e:  # units differ per row in some versions
        lines.append(f"   (traffic: {e})")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
