"""Per-kernel device times of one BASELINE config: python tools/config_kernels.py <config> [batch] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth

cfg = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
p = synth.params(cfg)
clouds = synth.frames(cfg, 0, B)
n = clouds.shape[1]
counts = np.full(B, n, np.int32)
with ObstacleProcessor(p, n, max_batch=B) as op:
    for _ in range(3):
        res = op.process_batch(clouds, counts)
    op.enable_kernel_timing(True)
    for _ in range(reps):
        op.process_batch_raw(clouds.ctypes.data, n, counts)
    kt = op.kernel_times()
    st = op.stage_times_us()
r = res[0]
print(f"config {cfg} batch {B}: N={r.n_input} M={r.n_crop} V={r.n_voxel} S={r.n_sor} P={r.n_remaining} C={r.n_clusters} L={r.n_cluster_points}")
tot = sum(v[0] for v in kt.values())
for k, (us, cnt) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    print("%-24s %4d launches/call %9.1f us/call %5.1f %%" % (k, cnt // reps, us / reps, 100 * us / tot))
print("total %.1f us/call; stages (last call): %s" % (tot / reps, {k: round(v) for k, v in st.items()}))
