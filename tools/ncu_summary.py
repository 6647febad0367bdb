"""Key counters of an ncu report (raw page) as JSON lines: python tools/ncu_summary.py rep [rep..]"""
import csv, io, json, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_blocks', 'launch__occupancy_limit_warps',
        'sm__maximum_warps_per_active_cycle_pct', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for d in rows[2:]:
        rec = {"report": rep, "kernel": d[hdr.index("Kernel Name")][:60]}
        for k in KEYS:
            if k in hdr:
                rec[k] = d[hdr.index(k)] + (" " + units[hdr.index(k)] if units[hdr.index(k)] else "")
        st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(d[hdr.index(h)]))
              for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
        rec["stalls_per_issue"] = {k: round(v, 2) for k, v in sorted(st, key=lambda x: -x[1])[:7]}
        print(json.dumps(rec, indent=1))
