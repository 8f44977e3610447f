"""Summarise `ncu -i rep --page raw --csv` output (file argument) as `metric [unit] = value` lines: the subset of
--set full that the profiles/ summaries quote (durations, DRAM bytes, pipe utilisations, stall samples)."""
import csv
import sys

WANT = ("Kernel Name", "dram__bytes", "gpu__dram_throughput", "gpu__time_duration", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "l1tex__data_pipe_lsu_wavefronts", "launch__registers", "launch__cluster_size", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit", "lts__t_sector_hit_rate", "sm__cycles_elapsed.max", "sm__pipe_tensor", "sm__pipe_fp64",
        "sm__inst_executed_pipe_fp64", "sm__pipe_fma", "sm__pipe_alu", "sm__throughput", "smsp__inst_executed.sum",
        "smsp__issue_active", "sm__warps_active", "l1tex__throughput", "lts__throughput", "smsp__pcsamp_warps_issue_stalled",
        "sm__inst_executed_pipe_lsu", "dram__throughput", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__sass_average_data_bytes_per_sector")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for i, k in enumerate(hdr):
        if any(k.startswith(w) for w in WANT) and r[i] not in ("", "0"):
            print(f"{k} [{units[i]}] = {r[i]}")
    print()
