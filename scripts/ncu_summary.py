#!/usr/bin/env python
"""scripts/ncu_summary.py <report.ncu-rep> — the few ncu numbers DESIGN.md / profiles/ quote (run where ncu is installed)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel", d.get("Kernel Name"), "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
            'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__inst_executed_pipe_fp64.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
            'smsp__thread_inst_executed_per_inst_executed.ratio']
    for k in keys:
        if k in d:
            print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = []
    for k, v in d.items():
        if 'issue_stalled' in k and k.endswith('_per_warp_active.pct') is False and k.endswith('.ratio') and 'not_issued' not in k:
            try:
                st.append((float(v), k))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:10]:
        print("  stall %6.3f %s" % (v, k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")))
