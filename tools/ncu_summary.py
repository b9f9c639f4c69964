"""Print the metrics we care about from an `ncu --page raw --csv` dump (one block per kernel)."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'sm__pipe_fp64_cycles_active', 'sm__inst_executed_pipe_fp64',
        'smsp__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct', 'lts__t_bytes.sum ', 'l1tex__t_bytes.sum ',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled', 'launch__grid_size',
        'launch__block_size', 'sm__inst_executed.sum', 'smsp__inst_executed_op', 'sm__sass_thread_inst_executed_op_d',
        'launch__waves_per_multiprocessor', 'dram__throughput', 'sm__inst_executed_pipe_uniform', 'smsp__inst_issued.sum',
        'sm__inst_executed_pipe_lsu', 'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_fma']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if any(w.strip() in h for w in WANT)]
for r in rows[2:]:
    print('-----')
    for i in idx:
        v = r[i]
        if v not in ('0', ''):
            print(f"  {hdr[i]} [{units[i]}] = {v}")
