import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','launch__shared_mem_per_block_dynamic','launch__grid_size','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct','smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f"{w:75s} {r[i]} {units[i]}")
