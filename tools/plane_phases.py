"""Phase breakdown (clock64 of frame 0 / CTA 0, last pass) of the frame-resident plane loop kernel.
Needs a library built with PCOP_NVCC_EXTRA=-DPCOP_PLANE_DEBUG_CLK (python -m pointcloud_obstacle_processing_b200._build --force)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth, load_library

NAMES = ["load slice + hypotheses", "score 8 + push", "cluster barrier", "adaptive-k replay (1 thread)", "(second batch)",
         "moments + push", "cluster barrier", "sum partials + eigen33 (1 thread)", "classify + scan + push",
         "cluster barrier", "write remaining + inliers"]
p = synth.params(2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
clouds = synth.frames(2, 0, B)
lib = load_library()
with ObstacleProcessor(p, clouds.shape[1], max_batch=B) as op:
    for _ in range(3):
        res = op.process_batch(clouds, np.full(B, clouds.shape[1], np.int32))
    out = (C.c_longlong * 16)()
    st = lib.pcop_debug_plane_loop_cycles(out)
    assert st == 0, st
    t = list(out)
    print("frame 0: V =", res[0].n_voxel, "P =", res[0].n_remaining, "passes =", res[0].n_plane_passes)
    for k, name in enumerate(NAMES):
        print("%-36s %9d cycles" % (name, t[k + 1] - t[k]))
    print("%-36s %9d cycles" % ("total", t[len(NAMES)] - t[0]))
