"""Phase breakdown (clock64 of block 0) of the fused small-cloud clustering kernel.
Needs a library built with PCOP_NVCC_EXTRA=-DPCOP_ECE_DEBUG_CLK (python -m pointcloud_obstacle_processing_b200._build --force)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth, load_library

NAMES = ["minmax+grid", "keys+sort", "unpack+heads", "hash build", "union", "flatten+sizes", "roots compaction",
         "roots sort", "rank+offsets", "member keys", "member sort+write", "centroid/radius"]
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
        print("%-20s %9d cycles" % (name, t[k + 1] - t[k]))
    print("%-20s %9d cycles" % ("total", t[12] - t[0]))
