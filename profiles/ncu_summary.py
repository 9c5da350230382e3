"""ncu report -> the per-kernel counter table committed under profiles/ (run here, no GPU needed):
    python profiles/ncu_summary.py gpurun_out/<report>.ncu-rep > profiles/<name>.csv"""
import csv, subprocess, sys
WANT = ['launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'sm__cycles_elapsed.avg.per_second',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
names = [r[hdr.index('Kernel Name')] for r in data]
print('metric,unit,' + ','.join('"' + n + '"' for n in names))
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(w + ',' + units[i] + ',' + ','.join(r[i] for r in data))
