"""Key metrics per kernel from an .ncu-rep: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_bytes.sum']
want += [x for x in h if x.startswith('smsp__average_warps_issue_stalled') and x.endswith('per_issue_active.ratio')]
idx = {w: h.index(w) for w in want if w in h}
units = rows[1]
for r in rows[2:]:
    print('----')
    for w, i in idx.items():
        v = r[i]
        if w.startswith('smsp__average_warps_issue_stalled'):
            try:
                if float(v) < 0.15: continue
            except ValueError: pass
            w = w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')
        print(f"  {w[:70]:70s} {v[:60]} {units[i]}")
