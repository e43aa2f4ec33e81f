#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (raw page): python scripts/ncu_keys.py file.ncu-rep [substring...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_config_size', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic'] + extra
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:110])
    for w in want:
        for i, h in enumerate(hdr):
            if h == w or (w in extra and w in h):
                print(f'   {h:85s} {r[i]:>18s} {units[i]}')
