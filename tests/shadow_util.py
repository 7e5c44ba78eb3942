"""Synthetic scenes for the shadow-casting tests (od.cpp:467-672, 817-833): TEST INFRASTRUCTURE."""
import numpy as np

from pointcloud_obstacle_processing_b200 import synth


def rigid(yaw, pitch, roll, t):
    """row-major 4x4 float32 of a rigid transform and of its inverse (what tf hands pcl_ros::transformPointCloud)"""
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]]) @ \
        np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    m = np.eye(4)
    m[:3, :3] = R
    m[:3, 3] = t
    return m.astype(np.float32), np.linalg.inv(m).astype(np.float32)


def shadow_scene(seed, n_clusters=7, noise=400):
    """a remaining cloud inside the params.yaml crop box: blobs (clusters, ascending member indices, size-ordered CSR
    like the pipeline's output) + unclustered noise; returns (cloud [P, 4], offsets, indices)"""
    p = synth.params(1)
    rng = np.random.default_rng(seed)
    f = np.float32
    pts, members = [], []
    base = 0
    for c in range(n_clusters):
        k = int(rng.integers(2, 60)) if c else 1  # one single-point cluster: skipped by od.cpp:574
        ctr = np.array([rng.uniform(p.x_min + 0.4, p.x_max - 0.4), rng.uniform(p.y_min + 0.4, p.y_max - 0.4),
                        rng.uniform(-0.3, 0.1)])
        blob = ctr + rng.normal(0, [0.08, 0.12, 0.05], (k, 3))
        pts.append(blob)
        members.append(np.arange(base, base + k))
        base += k
    pts.append(np.stack([rng.uniform(p.x_min, p.x_max, noise), rng.uniform(p.y_min, p.y_max, noise),
                         rng.uniform(-0.4, 0.2, noise)], 1))
    cloud = np.concatenate(pts).astype(f)
    perm = rng.permutation(len(cloud))  # scatter the members over the cloud
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    cloud = np.concatenate([cloud[perm], np.ones((len(cloud), 1), f)], 1)
    clusters = sorted((np.sort(inv[m]) for m in members), key=lambda m: (-len(m), m[0]))
    offsets = np.concatenate([[0], np.cumsum([len(m) for m in clusters])]).astype(np.int32)
    indices = np.concatenate(clusters).astype(np.int32)
    return cloud, offsets, indices
