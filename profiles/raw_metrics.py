#!/usr/bin/env python
"""Print the headline metrics of an `ncu --page raw --csv` dump. usage: raw_metrics.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
'launch__shared_mem_per_block_dynamic','launch__block_size','launch__grid_size','smsp__thread_inst_executed_per_inst_executed.ratio',
'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sectors_op_write.sum','lts__t_bytes.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
'smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_dispatch_stall','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_drain','smsp__pcsamp_warps_issue_stalled_membar','smsp__pcsamp_warps_issue_stalled_imc_miss','smsp__pcsamp_warps_issue_stalled_sleeping','smsp__pcsamp_warps_issue_stalled_tex_throttle']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:72s} {units[i]:16s}", [r[i] for r in data])
