#!/usr/bin/env python
"""Stall samples / executed instructions of solve_kernel per source region, from an ncu capture made with --import-source on.
usage: python scripts/ncu_source_regions.py <file.ncu-rep> [kernel-id]   (kernel id like ::solve_kernel:1)"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "::solve_kernel:1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, data, sass = None, None, [], []
stall_cols = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if hdr is None or len(r) <= iI:
        continue
    if r[0].isdigit():
        num = lambda v: float(v) if v not in ("", "-") else 0.0
        data.append((cur_file, int(r[0]), r[1], num(r[iS]), num(r[iI]), {k: num(r[i]) for k, i in stall_cols.items()}))
tot_s, tot_i = sum(d[3] for d in data), sum(d[4] for d in data)
# region markers from the full source text stored in the report (lines without instructions are not in the metric rows)
full = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--kernel-id", kid], capture_output=True, text=True).stdout
marks, in_soft = [], False
for r in csv.reader(full.splitlines()):
    if r and r[0] == "File Name":
        in_soft = r[1].endswith("soft.cuh")
        continue
    if not in_soft or len(r) < 2 or not r[0].isdigit():
        continue
    ln, text = int(r[0]), r[1]
    m = re.search(r"auto (\w+) = \[&\]", text)
    if m:
        marks.append((ln, "lambda " + m.group(1)))
    m = re.match(r"\s*// -{20,} (.*)", text)
    if m:
        marks.append((ln, m.group(1)[:50]))
    m = re.match(r"__device__ __forceinline__ \S+ (\w+)\(", text)
    if m:
        marks.append((ln, "fn " + m.group(1)))
    if re.match(r"__global__ void", text):
        marks.append((ln, "kernel prologue"))
marks.sort()


def region(f, l):
    if f != "soft.cuh":
        return f
    n = "prologue"
    for ln, nm in marks:
        if ln <= l:
            n = nm
        else:
            break
    return n


agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for f, ln, t, s, i, st in data:
    a = agg[region(f, ln)]
    a[0] += s
    a[1] += i
    a[2].update(st)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
print(f"{'region':52s} {'samples%':>8s} {'inst%':>7s}  top stalls")
for k, (s, i, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    top = " ".join(f"{n.replace('stall_', '')}:{100 * v / max(s, 1):.0f}%" for n, v in st.most_common(3))
    print(f"{k:52s} {100 * s / tot_s:8.1f} {100 * i / tot_i:7.1f}  {top}")
print("top lines by samples:")
for f, ln, t, s, i, st in sorted(data, key=lambda d: -d[3])[:30]:
    top = " ".join(f"{n.replace('stall_', '')}:{100 * v / max(s, 1):.0f}%" for n, v in collections.Counter(st).most_common(2))
    print(f"{100 * s / tot_s:5.1f}% inst {100 * i / tot_i:4.1f}%  {f}:{ln}: {t.strip()[:90]}  [{top}]")
