#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/summarize_ncu.py <round tag> <launch csv> <full .ncu-rep> [kernel regex]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, data = None, []
    for r in rows:
        if r[0] == "ID":
            hdr = r
        elif hdr and r[0].isdigit():
            data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
        k = d["Kernel Name"].split("(")[0][-70:]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out.write(f"## Launch list ({os.path.basename(path)}): gpu__time_duration.sum per kernel, {len(data)} launches\n\n")
    out.write("cold-cache, serialised (ncu replay): compare SHARES, not absolutes\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |\n")
    out.write("\n")


def full(rep, out, pattern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.write(f"## Full capture ({os.path.basename(rep)}), `ncu --set full --clock-control none --import-source on`\n\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pattern not in name:
            continue
        out.write(f"### launch id {r[0]}: `{name.split('(')[0]}` grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in hdr:
                out.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
        stalls = [(h, float(r[i].replace(",", "") or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        out.write("\nwarp stall reasons (average warps stalled per issue-active cycle, top 8):\n\n| reason | ratio |\n|---|---:|\n")
        for h, v in sorted(stalls, key=lambda t: -t[1])[:8]:
            out.write(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |\n")
        out.write("\n")


if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    pattern = sys.argv[4] if len(sys.argv) > 4 else "solve_kernel"
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md")
    with open(path, "w") as out:
        out.write(f"# ncu summary {tag}\n\nCommands (scripts/profile_round.sh): `ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 260 -c 400 python bench.py --steps 40 --warmup 10 --preroll 60 --no-cpu` (launch list) and `ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 150 -c 2` on the same command (full capture: one step launch and one reset-preparation launch), 4096 soft-torso envs, 1 B200.\n\n")
        launches(lcsv, out)
        full(rep, out, pattern)
    print(open(path).read())
