"""Print the headline metrics of an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'launch__grid_size', 'launch__block_size',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__warps_eligible.avg.per_cycle_active']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print('==', path, v[h.index('Kernel Name')][:70])
        for i, name in enumerate(h):
            if name in WANT or 'warp_issue_stalled' in name and name.endswith('per_warp_active.pct'):
                try:
                    if 'stalled' in name and float(v[i].replace(',', '')) < 2.0:
                        continue
                except ValueError:
                    pass
                print('  %-86s %s %s' % (name, v[i], u[i]))
