#!/usr/bin/env python3
"""Distil an .ncu-rep of the render kernel into the small JSON bench.py's `roofline` object reads
(profiles/r02_queue_ncu_metrics.json).  usage: ncu_metrics_json.py <rep> <samples_per_launch> <source label> > out.json"""
import csv, io, json, subprocess, sys
rep, samples, label = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]


def get(name):
    i = h.index(name)
    x = float(v[i].replace(",", ""))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1.0, "us": 1e-3}.get(u[i], 1.0)
    return x * scale


inst = get("smsp__inst_executed.sum")
out = {
    "source": label,
    "kernel": v[h.index("Kernel Name")],
    "kernel_ms_under_ncu": get("gpu__time_duration.sum"),
    "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
    "lane_efficiency": get("smsp__thread_inst_executed_per_inst_executed.ratio") / 32.0,
    "threads_per_warp_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warp_inst_per_sample": inst / samples,
    "issue_slot_utilisation": get("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
    "ipc_per_sm": get("sm__inst_executed.avg.per_cycle_elapsed"),
    "l1_hit_rate": get("l1tex__t_sector_hit_rate.pct") / 100.0,
    "l2_hit_rate": get("lts__t_sector_hit_rate.pct") / 100.0,
    "l1_sectors_per_launch": get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    "registers_per_thread": get("launch__registers_per_thread"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
}
print(json.dumps(out, indent=1))
