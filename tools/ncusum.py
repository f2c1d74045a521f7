"""Summarise an `ncu --page raw --csv` dump: one block per profiled launch with the metrics the roofline needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__cycles_active.avg']
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if h in want or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if 'issue_stalled' in h and v < 0.1:
                continue
            print('  %-90s %-12s %s' % (h, units[i], r[i]))
