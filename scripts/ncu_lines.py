#!/usr/bin/env python3
"""Per-source-line totals of an ncu report captured with --import-source on (kernels built with -lineinfo).

usage: ncu_lines.py <rep> [--top N] [--kernel substr] [--by inst|samples]
Prints, per source line of the (first matching) kernel: share of warp instructions executed, share of stall samples,
average active threads, and the dominant stall reasons; then totals per file and per function-sized line range.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = 60
    by = "inst"
    kernel = None
    a = sys.argv[2:]
    while a:
        k = a.pop(0)
        if k == "--top": top = int(a.pop(0))
        elif k == "--by": by = a.pop(0)
        elif k == "--kernel": kernel = a.pop(0)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, cur_fn, hdr = None, None, None
    lines = []   # (file, line, text, inst, thr_inst, samples, {stall: n})
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]; continue
        if r[0] == "Function Name":
            cur_fn = r[1]; continue
        if r[0] == "Line No":
            hdr = r; ix = {}
            for i, k in enumerate(hdr):
                ix.setdefault(k, i)
            stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
            continue
        if hdr is None or r[0] == "":
            continue
        if kernel and kernel not in (cur_fn or ""):
            continue
        try:
            inst = int(r[ix["Instructions Executed"]]); thr = int(r[ix["Thread Instructions Executed"]]); smp = int(r[ix["# Samples"]])
        except ValueError:
            continue
        st = {k: int(r[ix[k]]) for k in stalls}
        lines.append((cur_file.split("/")[-1], int(r[0]), r[1].strip(), inst, thr, smp, st))
    ti = sum(l[3] for l in lines) or 1
    ts = sum(l[5] for l in lines) or 1
    print(f"total warp instructions {ti}, samples {ts}, avg threads {sum(l[4] for l in lines) / ti:.2f}")
    key = (lambda l: -l[3]) if by == "inst" else (lambda l: -l[5])
    print(f"{'file:line':28s} {'inst%':>6s} {'smp%':>6s} {'thr':>5s}  top stalls | source")
    for f, ln, text, inst, thr, smp, st in sorted(lines, key=key)[:top]:
        tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        s = " ".join(f"{k[6:][:9]}={100 * v / max(smp, 1):.0f}%" for k, v in tops if v)
        print(f"{f + ':' + str(ln):28s} {100 * inst / ti:6.2f} {100 * smp / ts:6.2f} {thr / max(inst, 1):5.1f}  {s:40s} | {text[:90]}")
    print("\nper file:")
    files = {}
    for f, ln, text, inst, thr, smp, st in lines:
        d = files.setdefault(f, [0, 0, 0]); d[0] += inst; d[1] += smp; d[2] += thr
    for f, d in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:24s} inst {100 * d[0] / ti:6.2f}%  samples {100 * d[1] / ts:6.2f}%  thr {d[2] / max(d[0], 1):5.1f}")
    allst = {}
    for l in lines:
        for k, v in l[6].items():
            allst[k] = allst.get(k, 0) + v
    print("\nstall totals: " + " ".join(f"{k[6:]}={100 * v / ts:.1f}%" for k, v in sorted(allst.items(), key=lambda kv: -kv[1]) if v))


if __name__ == "__main__":
    main()
