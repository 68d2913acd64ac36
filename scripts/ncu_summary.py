"""Print a compact per-kernel summary of an `ncu --set full` raw CSV export (ncu -i X.ncu-rep --page raw --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index('Kernel Name')
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max',
        'sm__cycles_active.min', 'sm__cycles_active.avg', 'sm__cycles_active.max',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active']
seen = {}
for i, r in enumerate(data):
    nm = r[ki].split('(')[0]
    seen.setdefault(nm, []).append(i)
for nm, idx in seen.items():
    pick = idx[len(idx) // 2]
    print(f"== {nm}  ({len(idx)} captured, showing #{pick})")
    for w in want:
        if w in hdr:
            j = hdr.index(w)
            print(f"   {w:85s} {data[pick][j]:>16s} {units[j]}")
