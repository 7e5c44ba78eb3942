"""Run a few batched calls of the BASELINE configs[1] pipeline (for ncu captures): python tools/run_batch.py [B] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = synth.params(2)
clouds = synth.frames(2, 0, B)
with ObstacleProcessor(p, clouds.shape[1], max_batch=B) as op:
    for _ in range(reps):
        res = op.process_batch(clouds, np.full(B, clouds.shape[1], np.int32))
print("ok", B, reps, res[0].n_remaining, res[0].n_clusters)
