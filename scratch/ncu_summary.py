import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','launch__grid_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
 'smsp__pcsamp_warps_issue_stalled_long_scoreboard','smsp__pcsamp_warps_issue_stalled_barrier','smsp__pcsamp_warps_issue_stalled_short_scoreboard','smsp__pcsamp_warps_issue_stalled_wait','smsp__pcsamp_warps_issue_stalled_selected','smsp__pcsamp_warps_issue_stalled_not_selected',
 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle','smsp__pcsamp_warps_issue_stalled_lg_throttle','smsp__pcsamp_warps_issue_stalled_mio_throttle','smsp__pcsamp_warps_issue_stalled_tex_throttle','smsp__pcsamp_warps_issue_stalled_dispatch_stall','smsp__pcsamp_warps_issue_stalled_branch_resolving','smsp__pcsamp_warps_issue_stalled_no_instructions','smsp__pcsamp_warps_issue_stalled_membar','smsp__pcsamp_warps_issue_stalled_drain','smsp__pcsamp_warps_issue_stalled_sleeping','smsp__pcsamp_warps_issue_stalled_misc','smsp__pcsamp_warps_issue_stalled_imc_miss']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:90])
    for k in want:
        if k in hdr:
            i = hdr.index(k); print('  %-85s %s %s' % (k, r[i], units[i]))
