"""Summary of an ncu full capture of k_tile_pass: headline metrics, stall mix and instruction mix per launch."""
import csv, sys, collections, subprocess
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
for k in keys:
    for i, h in enumerate(hdr):
        if h == k or (k != "Kernel Name" and h.startswith(k) and h == k):
            print("%-70s %-10s %s" % (h, units[i], [r[i][:44] for r in rows[2:]]))
print("stall ratios (warps stalled per issue-active cycle):")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        vals = [float(r[i]) for r in rows[2:]]
        if max(vals) > 0.15:
            print("   %-22s %s" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], ["%.2f" % v for v in vals]))
