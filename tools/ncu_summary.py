#!/usr/bin/env python
"""Print the roofline-relevant metrics of every kernel in an `ncu --page raw --csv` export."""
import csv
import sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-' * 100)
    for w in WANT:
        if w in idx:
            print(f'{w:72s} {r[idx[w]]:>22s} {units[idx[w]]}')
