#!/usr/bin/env python3
"""scheduling jitter of this host: the largest gaps between two consecutive clock reads of a tight loop (a vCPU that is
descheduled for milliseconds stalls every kernel launch of a launch-bound pipeline; recorded next to the benchmarks)"""
import json
import os
import sys
import time

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
t_end = time.perf_counter_ns() + int(secs * 1e9)
last = time.perf_counter_ns()
gaps = []
n = 0
while last < t_end:
    now = time.perf_counter_ns()
    d = now - last
    if d > 100_000:
        gaps.append(d)
    last = now
    n += 1
steal = None
try:
    steal = int(open("/proc/stat").readline().split()[8])
except Exception:  # noqa: BLE001
    pass
print(json.dumps({"loop_iterations": n, "gaps_over_100us": len(gaps), "gaps_over_1ms": sum(g > 1_000_000 for g in gaps),
                  "max_gap_ms": max(gaps) / 1e6 if gaps else 0.0, "total_gap_ms": sum(gaps) / 1e6, "seconds": secs,
                  "cores": len(os.sched_getaffinity(0)), "loadavg": os.getloadavg(), "proc_stat_steal_jiffies": steal}))
