#!/usr/bin/env python3
"""profiles/ summaries from ncu output (the evidence bench.py's roofline object points at):
    python tools/ncu_summarize.py launches gpurun_out/x_launches.csv  > profiles/rN_ncu_launch_list.csv
        x_launches.csv = ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py 1024 2
        (two steps: the SECOND, warm one is summarised)
    python tools/ncu_summarize.py full gpurun_out/x_full.ncu-rep      > profiles/rN_ncu_full_summary.csv
        x_full.ncu-rep = ncu --set full --clock-control none --import-source on [-k regex:...] -o ... python tools/profile_step.py 1024 2"""
import csv
import subprocess
import sys
from collections import OrderedDict

FULL = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = [(short(r[ik]), float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[iu]]) for r in rows[1:]]
    first = sys.argv[3] if len(sys.argv) > 3 else "k_smooth3"  # a step starts with the smooth: its last launch opens the warm step
    seq = seq[max(i for i, (k, _) in enumerate(seq) if k.startswith(first)):]
    agg = OrderedDict()
    for k, ms in seq:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    w = csv.writer(sys.stdout, lineterminator="\n")
    print(f'"# ncu --metrics gpu__time_duration.sum --clock-control none, python tools/profile_step.py 1024 2: second (warm) step of '
          f'G1024 -p1 -l1 -b1 Lewiner; {len(seq)} launches, {tot:.3f} ms under ncu (cold-cache, serialised)"')
    w.writerow(["kernel", "launches_per_step", "ms_per_step_under_ncu", "share_of_step"])
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{ms:.4f}", f"{ms / tot:.4f}"])


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics", ",".join(FULL)], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    ik = h.index("Kernel Name")
    cols = [i for i, n in enumerate(h) if n in FULL]
    w = csv.writer(sys.stdout, lineterminator="\n")
    w.writerow(["Kernel Name"] + [h[i] for i in cols])
    w.writerow([""] + [u[i] for i in cols])
    seen = set()
    for r in rows[2:][::-1]:  # the last capture of every kernel signature (+ grid) = the warm step
        key = (r[ik], r[h.index("Grid Size")] if "Grid Size" in h else "")
        if key in seen:
            continue
        seen.add(key)
        w.writerow([r[ik]] + [r[i] for i in cols])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
