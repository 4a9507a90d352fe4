#!/usr/bin/env python3
"""Per-SASS-range totals of an ncu report (source page): instructions, samples, lanes and the main stall reasons per block of
N instructions, with the address of the block, so that the stages of an inlined kernel can be told apart.
usage: ncu_sass_ranges.py <rep> [block=128]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ix = {}
for i, k in enumerate(h):
    ix.setdefault(k, i)
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
ti = sum(int(r[ix['Instructions Executed']]) for r in data); ts = sum(int(r[ix['# Samples']]) for r in data)
print(f"# {len(data)} SASS instructions, {ti/1e6:.0f} M warp-inst, {ts} samples")
for b in range(0, len(data), B):
    seg = data[b:b + B]
    ie = sum(int(r[ix['Instructions Executed']]) for r in seg); te = sum(int(r[ix['Thread Instructions Executed']]) for r in seg)
    s = sum(int(r[ix['# Samples']]) for r in seg)
    if not ie: continue
    st = {k: sum(int(r[ix[k]]) for r in seg) for k in stalls}
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"{b:5d} @{seg[0][0][-5:]} inst {100*ie/ti:5.1f}%  smp {100*s/ts:5.1f}%  thr {te/max(ie,1):5.1f}  " + " ".join(f"{k[6:][:8]}={100*v/max(s,1):.0f}%" for k, v in tops))
