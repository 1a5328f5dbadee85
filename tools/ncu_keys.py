"""Key metrics per kernel of an ncu report: python tools/ncu_keys.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [('gpu__time_duration.sum','time'), ('launch__registers_per_thread','regs'), ('launch__grid_size','grid'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active','occ%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','fp64pipe%'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio','thr/inst'),
        ('dram__bytes_read.sum','dram_rd'), ('dram__bytes_write.sum','dram_wr'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed','dram%'),
        ('smsp__inst_executed.sum','inst'),
        ('smsp__inst_executed_pipe_fp64.sum','fp64inst'),
        ('l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','local_ld_sect'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','st_longsb'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','st_wait'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','st_shortsb'),
        ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','st_noinst'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','st_mathpipe'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','st_mio'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','st_lg'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','st_barrier'),
        ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','st_branch'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','st_notsel'),
        ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','st_dispatch'),
        ('smsp__average_warps_issue_stalled_selected_per_issue_active.ratio','st_selected'),
        ]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('----', r[ki][:60])
    line = []
    for w, n in want:
        if w in hdr:
            i = hdr.index(w); line.append(f"{n}={r[i]}{units[i] if n in ('time','dram_rd','dram_wr') else ''}")
    print('  ' + '  '.join(line))
