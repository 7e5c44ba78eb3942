"""Phase breakdown (clock64 of block 0) of the fused small-cloud clustering kernel.
Needs a library built with PCOP_NVCC_EXTRA=-DPCOP_ECE_DEBUG_CLK (python -m pointcloud_obstacle_processing_b200._build --force)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth, load_library

NAMES = ["minmax+grid", "hash insert+cell ids", "count+scan+scatter", "union", "sizes+minidx", "roots+rank",
         "offsets+labels", "member lists", "centroid/radius"]
p = synth.params(2)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
clouds = synth.frames(2, 0, B)
lib = load_library()
with ObstacleProcessor(p, clouds.shape[1], max_batch=B) as op:
    for _ in range(3):
        res = op.process_batch(clouds, np.full(B, clouds.shape[1], np.int32))
    out = (C.c_longlong * 16)()
    st = lib.pcop_debug_ece_small_cycles(out)
    assert st == 0, st
    t = list(out)
    print("frame 0: P =", res[0].n_remaining, "C =", res[0].n_clusters)
    for k, name in enumerate(NAMES):
        print("%-24s %9d cycles" % (name, t[k + 1] - t[k]))
    print("%-24s %9d cycles" % ("total", t[len(NAMES)] - t[0]))
    print("warp 0, first pass: lookups %d cycles, neighbour walk %d cycles" % (t[10] - t[3], t[11] - t[10]))
