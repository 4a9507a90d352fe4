#!/usr/bin/env python3
"""Stall-reason totals per SASS index range.  usage: ncu_regions.py <rep> lo:hi[:name] ..."""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
ts = sum(int(r[ix['# Samples']]) for r in data); ti = sum(int(r[ix['Instructions Executed']]) for r in data)
regs = []
for a in sys.argv[2:]:
    p = a.split(":"); regs.append((int(p[0]), int(p[1]), p[2] if len(p) > 2 else a))
if not regs: regs = [(0, len(data), "all")]
print(f"{'region':28s} {'inst%':>6s} {'smp%':>6s} {'thr':>5s} | " + " ".join(f"{k[6:][:8]:>8s}" for k in stalls))
for lo, hi2, name in regs:
    seg = data[lo:hi2]
    ie = sum(int(r[ix['Instructions Executed']]) for r in seg); te = sum(int(r[ix['Thread Instructions Executed']]) for r in seg)
    s = sum(int(r[ix['# Samples']]) for r in seg)
    st = [sum(int(r[ix[k]]) for r in seg) for k in stalls]
    print(f"{name:28s} {100*ie/ti:6.1f} {100*s/ts:6.1f} {te/max(ie,1):5.1f} | " + " ".join(f"{100*x/ts:8.1f}" for x in st))
