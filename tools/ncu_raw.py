"""Print selected metrics of an ncu report. usage: python tools/ncu_raw.py X.ncu-rep [substr ...]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
                        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
                        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit",
                        "smsp__average_warp", "smsp__warp_issue_stalled", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct",
                        "lts__throughput.avg.pct"]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:80])
    for i, k in enumerate(h):
        if any(s in k for s in keys):
            print("  %-90s %s %s" % (k, r[i], rows[1][i]))
