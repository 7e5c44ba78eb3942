"""Randomised parity: random scenes x random parameter sets through the whole pipeline, CUDA library against the CPU
oracle (every index array bit-exact, floats as in parity_util.compare_frames).  The reference ships no edge-case
tests; this sweeps the combinations the hand-written cases do not: stage subsets, leaf / tolerance ratios, tiny and
empty clouds, non-finite coordinates, points on the crop faces, duplicate points."""
import numpy as np
import pytest

import oracle_lib as O
from parity_util import compare_frames
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth
from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi

pytestmark = pytest.mark.gpu


def fuzz_case(seed):
    rng = np.random.default_rng(1000 + seed)
    f = np.float32
    n = int(rng.choice([0, 1, 2, 3, 4, 17, 300, 3000, 12000, 40000]))
    ext = float(rng.choice([2.0, 8.0, 30.0]))
    parts = []
    if n:
        n_ground = int(n * rng.uniform(0.0, 0.7))
        n_blob = int((n - n_ground) * rng.uniform(0.3, 1.0))
        n_noise = n - n_ground - n_blob
        tilt = rng.normal(0, 0.03, 2)
        g = np.stack([rng.uniform(-ext, ext, n_ground), rng.uniform(-ext, ext, n_ground), np.zeros(n_ground)], 1)
        g[:, 2] = g[:, 0] * tilt[0] + g[:, 1] * tilt[1] + rng.normal(0, 0.01 * ext / 8, n_ground) - 0.5
        parts.append(g)
        k = max(1, int(rng.integers(1, 25)))
        ctr = np.stack([rng.uniform(-ext, ext, k), rng.uniform(-ext, ext, k), rng.uniform(-0.4, 1.0, k)], 1)
        which = rng.integers(0, k, n_blob)
        parts.append(ctr[which] + rng.normal(0, rng.uniform(0.02, 0.3) * ext / 8, (n_blob, 3)))
        parts.append(np.stack([rng.uniform(-ext, ext, n_noise), rng.uniform(-ext, ext, n_noise),
                               rng.uniform(-1.0, 2.0, n_noise)], 1))
    pts = (np.concatenate(parts) if parts else np.zeros((0, 3))).astype(f)
    rng.shuffle(pts)
    p = synth.params(1)
    lim = ext * rng.uniform(0.5, 1.1)
    p.x_min, p.x_max = -lim, lim * rng.uniform(0.6, 1.0)
    p.y_min, p.y_max = -lim * rng.uniform(0.6, 1.0), lim
    p.z_min, p.z_max = -1.2, 1.5
    if len(pts) >= 16:  # special values: non-finite coordinates, points on the crop faces, exact duplicates
        m = len(pts)
        for col in range(3):
            pts[rng.integers(0, m, max(1, m // 200)), col] = np.nan
        pts[rng.integers(0, m, 3), rng.integers(0, 3, 3)] = np.inf
        pts[rng.integers(0, m, 2), 0] = -np.inf
        pts[rng.integers(0, m)] = [p.x_max, p.y_max, p.z_max]
        pts[rng.integers(0, m)] = [p.x_min, p.y_min, p.z_min]
        dup = rng.integers(0, m, max(2, m // 50))
        pts[dup] = pts[rng.integers(0, m, len(dup))]
    cloud = np.concatenate([pts, np.ones((len(pts), 1), f)], 1)
    p.downsample_size = float(rng.choice([0.02, 0.05, 0.1, 0.25, 0.5])) * ext / 8
    p.euc_cluster_tolerance = p.downsample_size * float(rng.choice([1.2, 2.0, 3.5, 6.0]))
    p.euc_min_cluster_size = int(rng.choice([1, 2, 5, 20]))
    p.euc_max_cluster_size = int(rng.choice([30, 500, 100000]))
    p.plane_segment_dist_thres = float(rng.choice([0.02, 0.05, 0.2])) * ext / 8
    p.statistical_outlier_meanK = int(rng.choice([2, 8, 15]))
    p.statistical_outlier_stdDevThres = float(rng.choice([0.5, 1.0, 4.0]))
    p.enable_crop = int(rng.random() < 0.8)
    p.enable_voxel = int(rng.random() < 0.85)
    if not p.enable_crop and not p.enable_voxel:
        p.enable_crop = 1  # (non-finite points reach the k-d stages only through a crop-less, voxel-less chain)
    p.enable_sor = int(rng.random() < 0.4 and n <= 12000)
    p.enable_plane = int(rng.random() < 0.7)
    p.enable_cluster = 1
    p.outputs = abi.OUT_ALL
    return p, cloud


@pytest.mark.parametrize("seed", range(64))
def test_fuzz_pipeline(seed):
    p, cloud = fuzz_case(seed)
    with ObstacleProcessor(p, max(len(cloud), 1)) as op:
        g = op.process(cloud)
    o = O.process(p, cloud)
    compare_frames(g, o, p, f"seed {seed} (n={len(cloud)}, crop={p.enable_crop} voxel={p.enable_voxel} sor={p.enable_sor} "
                            f"plane={p.enable_plane}, leaf={p.downsample_size:.3f}, tol={p.euc_cluster_tolerance:.3f}): ")
