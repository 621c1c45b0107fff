"""Key metrics of every kernel in an ncu report: python tools/ncu_summary.py X.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in want:
        if k in d:
            print(f"{k} = {d[k][:110]}")
    st = {k: float(v) for k, v in d.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and v not in ('', 'n/a')}
    print("stalls per issue:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
    print()
