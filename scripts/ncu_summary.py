"""Summarise a raw-page CSV export of an ncu report (ncu -i x.ncu-rep --page raw --csv > x.csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("== kernel:", d.get("Kernel Name"), "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
            'smsp__sass_average_branch_targets_threads_uniform.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
            'sm__sass_thread_inst_executed_op_dadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_dmul_pred_on.sum',
            'sm__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__sass_thread_inst_executed_op_fadd_pred_on.sum',
            'sm__sass_thread_inst_executed_op_fmul_pred_on.sum', 'sm__sass_thread_inst_executed_op_ffma_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
            'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'smsp__thread_inst_executed.sum']
    for k in keys:
        if k in d:
            print("  %-70s %-12s %s" % (k, u[k], d[k]))
    print("  -- pipe utilisation (pct of peak)")
    for k in hdr:
        if ('pipe' in k and 'pct_of_peak_sustained_active' in k):
            print("  %-70s %s" % (k, d[k]))
    print("  -- warp stall reasons (warps per issue-active cycle)")
    st = [(float(d[k].replace(',', '')), k) for k in hdr if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio') and 'not_issued' not in k]
    for v, k in sorted(st, reverse=True)[:10]:
        print("  %-40s %.3f" % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
