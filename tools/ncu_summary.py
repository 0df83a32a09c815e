"""Key figures of `ncu --set full` captures exported with `ncu -i x.ncu-rep --page raw --csv` (one row per profiled launch).
    python tools/ncu_summary.py file.raw.csv [...]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle")]

for fn in sys.argv[1:]:
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## {fn}: {d.get('Kernel Name', '?')}")
        for k, label in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                print(f"  {label:28s} {d[k]} {u.get(k, '')}")
        print()
