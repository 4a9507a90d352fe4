#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + per-region SASS statistics.  usage: ncu_summary.py <rep> [--sass lo hi]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_shared_atom.sum', 'launch__occupancy_limit_registers',
        'sm__cycles_elapsed.max']
print("#", rows[2][h.index("Kernel Name")] if "Kernel Name" in h else "")
for i, k in enumerate(h):
    if k in want or ('issue_stalled' in k and 'per_issue_active' in k):
        print(f'{k:95s} {u[i]:16s} {v[i]}')
if "--nosass" in sys.argv:
    sys.exit(0)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ix = {k: i for i, k in enumerate(h)}
ti = sum(int(r[ix['Instructions Executed']]) for r in data); ts = sum(int(r[ix['# Samples']]) for r in data)
print(f"# SASS: {len(data)} instructions, {ti} warp-inst executed, {ts} samples")
if "--sass" in sys.argv:
    a = sys.argv.index("--sass"); lo, hi2 = int(sys.argv[a + 1]), int(sys.argv[a + 2])
    for i in range(lo, min(hi2, len(data))):
        r = data[i]
        print(f"{i:5d} {int(r[ix['Instructions Executed']])/1e6:7.2f}M thr {r[ix['Avg. Threads Executed']]:>5s} smp {r[ix['# Samples']]:>6s} lsb {r[ix['stall_long_sb']]:>5s} ssb {r[ix['stall_short_sb']]:>5s} noi {r[ix['stall_no_inst']]:>4s} wait {r[ix['stall_wait']]:>4s} br {r[ix['stall_branch_resolving']]:>4s} mio {r[ix['stall_mio']]:>4s} | {r[ix['Source']][:100]}")
else:
    B = 64
    for b in range(0, len(data), B):
        seg = data[b:b + B]
        ie = sum(int(r[ix['Instructions Executed']]) for r in seg); te = sum(int(r[ix['Thread Instructions Executed']]) for r in seg)
        s = sum(int(r[ix['# Samples']]) for r in seg)
        if ie: print(f"{b:5d} inst {100*ie/ti:5.1f}%  samples {100*s/ts:5.1f}%  avgthr {te/max(ie,1):5.1f}")
