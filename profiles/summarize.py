#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full) and/or a launch-list csv into the small text summaries that are
committed under profiles/.   usage: summarize.py rep <file.ncu-rep> <out.txt> | launches <file.csv> <out.txt>"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__pipe_alu_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__throughput.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct", "dram__cycles_active.avg.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "sm__cycles_elapsed.max", "smsp__average_warp", "smsp__warp_issue_stalled", "launch__shared_mem_per_block"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {path}\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"\n== kernel: {name}  (ID {r[0]})\n")
            for h, u, v in zip(hdr, units, r):
                if any(k in h for k in KEYS) and "min." not in h and "max.pct" not in h and ".sum.pct" not in h:
                    f.write(f"{h} [{u}] = {v}\n")


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    d = defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) > mv:
            try:
                d[r[kn]].append(float(r[mv].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    with open(out, "w") as f:
        f.write(f"# launch list ({path}): ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised -> compare SHARES\n")
        f.write(f"{'kernel':90s} {'n':>5s} {'total_ms':>12s} {'avg_ms':>10s} {'share':>7s}\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:90]:90s} {len(v):5d} {sum(v) / 1e6:12.3f} {sum(v) / len(v) / 1e6:10.4f} {sum(v) / tot:7.4f}\n")


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
