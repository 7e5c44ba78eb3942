"""All single-frame BASELINE configs through the library: p50 single-frame latency (host frame in, results on the
host), batched throughput with the frames resident in HBM, and the frame's counts.
python tools/config_sweep.py [out.json]     (configs 1-4; config 5 = the batch of config 2 that bench.py measures)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # device memory for the resident frames only

from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth

BATCH = {1: 256, 2: 256, 3: 64, 4: 8}
WHAT = {1: "VLP-16 30k, params.yaml (crop, voxel 0.015, SOR, plane loop, ECE, centroid/radius)",
        2: "HDL-64 120k (crop, voxel 0.1, plane loop, ECE, centroid/radius)",
        3: "depth camera 640x480 (crop, voxel 0.02, ECE tol 0.05, centroid/radius)",
        4: "adversarial 1M points, three 300k-point components + 50k-point chain (voxel 0.01, ECE, centroid/radius)"}
rows = []
for cfg in (1, 2, 3, 4):
    p = synth.params(cfg)
    n = synth.points_per_frame(cfg)
    B = BATCH[cfg]
    clouds = synth.frames(cfg, 0, min(B, 16))
    with ObstacleProcessor(p, n, max_batch=1) as op:
        for _ in range(5):
            fr = op.process(clouds[0])
        ts = []
        for _ in range(50 if cfg < 4 else 15):
            a = time.perf_counter()
            op.process(clouds[0])
            ts.append((time.perf_counter() - a) * 1e3)
        ts.sort()
        dev_us = op.last_elapsed_us
        launches = op.last_launch_count
        occ_ms = None
        if cfg in (1, 2):  # the occupancy-grid product of the frame: initial data set + shadows + obstacle marks
            if cfg == 2:
                q = p.copy()
                q.block_size = 0.25
                op.set_params(q)
            m = np.eye(4, dtype=np.float32)
            m[:3, 3] = [p.x_max + 0.8, 0.5 * (p.y_min + p.y_max), 1.2]
            mi = np.linalg.inv(m).astype(np.float32)
            to = []
            for _ in range(30):
                a = time.perf_counter()
                g0, _, _ = op.occupancy_grid(clouds[0])
                op.handle_shadow_casting(g0, fr.remaining_cloud, fr.cluster_offsets, fr.cluster_indices, mi, m)
                to.append((time.perf_counter() - a) * 1e3)
            occ_ms = sorted(to)[len(to) // 2]
    reps = np.concatenate([clouds] * ((B + len(clouds) - 1) // len(clouds)))[:B]
    dev = torch.from_numpy(np.ascontiguousarray(reps)).to("cuda:0")
    counts = np.full(B, n, np.int32)
    with ObstacleProcessor(p, n, max_batch=B) as op:
        for _ in range(2):
            op.process_batch_raw(dev.data_ptr(), n, counts)
        torch.cuda.synchronize()
        a = time.perf_counter()
        K = 4
        for _ in range(K):
            op.process_batch_raw(dev.data_ptr(), n, counts)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - a) / K
    row = {"config": cfg, "what": WHAT[cfg], "points_per_frame": n, "p50_latency_ms": ts[len(ts) // 2],
           "device_us_single_frame": dev_us, "launches_single_frame": int(launches),
           "occupancy_grid_plus_shadows_p50_ms": occ_ms, "batch": B,
           "batched_frames_per_s": B / dt, "batched_points_per_s": B * n / dt,
           "counts": {"n_crop": fr.n_crop, "n_voxel": fr.n_voxel, "n_sor": fr.n_sor, "n_remaining": fr.n_remaining,
                      "n_clusters": fr.n_clusters, "n_cluster_points": fr.n_cluster_points}}
    rows.append(row)
    print("config %d: p50 %.3f ms (device %.0f us, %d launches) | batch %d: %.0f frames/s, %.2f G points/s | %s" % (
        cfg, row["p50_latency_ms"], dev_us, launches, B, row["batched_frames_per_s"], row["batched_points_per_s"] / 1e9,
        row["counts"]))
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
