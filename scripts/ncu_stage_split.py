#!/usr/bin/env python3
"""Attribute the SASS of k_render_queue to its stages: joins `nvdisasm -gi` line info (which ccu_queue.cuh line an instruction
belongs to, directly or through its innermost inlined-at record) with the per-instruction counters of an ncu report.
usage: ncu_stage_split.py <rep> <cubin> <function-substring> [block]"""
import csv, io, re, subprocess, sys
rep, cubin, fn = sys.argv[1:4]
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and fn in l)
qline = []   # per instruction: ccu_queue.cuh line (or None)
cur = None
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if m.group(1).endswith("ccu_queue.cuh"): cur = int(m.group(2))
        elif m.group(3) and m.group(3).endswith("ccu_queue.cuh"): cur = int(m.group(4))
        # otherwise keep the last known queue line (code of a helper inlined deeper)
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,5}\*/", l):
        qline.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ix = {}
for i, k in enumerate(h): ix.setdefault(k, i)
n = min(len(data), len(qline))
print(f"# {len(data)} profiled instructions, {len(qline)} disassembled")
ti = sum(int(r[ix['Instructions Executed']]) for r in data); ts = sum(int(r[ix['# Samples']]) for r in data)
for b in range(0, n, B):
    seg = data[b:b + B]
    ie = sum(int(r[ix['Instructions Executed']]) for r in seg)
    if not ie: continue
    s = sum(int(r[ix['# Samples']]) for r in seg); te = sum(int(r[ix['Thread Instructions Executed']]) for r in seg)
    ni = sum(int(r[ix['stall_no_inst']]) for r in seg) if 'stall_no_inst' in ix else 0
    ql = [q for q in qline[b:b + B] if q]
    print(f"{b:5d} inst {100*ie/ti:5.2f}% smp {100*s/ts:5.2f}% thr {te/max(ie,1):5.1f} no_inst {100*ni/max(s,1):3.0f}%  queue.cuh lines {min(ql) if ql else 0}-{max(ql) if ql else 0}")
