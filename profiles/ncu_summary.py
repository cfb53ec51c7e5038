import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.max','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.sum','smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_not_selected','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_dispatch_stall','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_tex_throttle','sm__sass_thread_inst_executed_op_fp64_pred_on.sum','sm__inst_executed_pipe_fp64.sum']
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr:
            i=hdr.index(w); print('%-70s %s %s'%(w, r[i], units[i]))
