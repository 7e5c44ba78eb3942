"""Single-frame latency breakdown: per-kernel device time (CUDA-event pairs) + wall clock of one pcop_process call.
python tools/latency_trace.py [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
p = synth.params(2)
cloud = synth.frame(2, 0)
with ObstacleProcessor(p, len(cloud), max_batch=1) as op:
    for _ in range(10):
        op.process(cloud)
    ts = []
    for _ in range(reps):
        a = time.perf_counter()
        op.process(cloud)
        ts.append((time.perf_counter() - a) * 1e6)
    ts.sort()
    print("wall p50 %.0f us  min %.0f us (pageable numpy input)" % (ts[len(ts) // 2], ts[0]))
    op.enable_kernel_timing(True)
    for _ in range(reps):
        op.process(cloud)
    kt = op.kernel_times()
    tot = 0.0
    for k, (us, n) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        print("%-22s %6.1f us x %.1f launches per frame" % (k, us / n, n / reps))
        tot += us / reps
    print("sum of kernel time per frame %.0f us, launches per frame %.0f, device elapsed %.0f us" % (
        tot, sum(n for _, n in kt.values()) / reps, op.last_elapsed_us))
    print(op.stage_times_us())
