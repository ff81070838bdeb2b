"""Pick the judged metrics out of an `ncu --page raw --csv` export: one column per captured launch."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sm__cycles_elapsed.max", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio"]
def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    data = rows[2:]
    name_i = hdr.index("Kernel Name")
    cols = {k: hdr.index(k) for k in KEYS if k in hdr}
    # every warp-state (stall reason per issue) and pipe-utilisation metric of the capture as well
    for i, k in enumerate(hdr):
        if k not in cols and (("issue_stalled" in k and k.endswith("per_issue_active.ratio")) or k.startswith("sm__pipe_") or k.startswith("sm__inst_executed_pipe_")):
            cols[k] = i
    names = [r[name_i].split("(")[0].replace("kzg::", "").replace("void ", "") for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for k, i in cols.items():
        u = units[i]
        print(f"| {k} [{u}] | " + " | ".join(r[i] for r in data) + " |")
if __name__ == "__main__":
    main(sys.argv[1])
