"""GPU parity: the CUDA library (through the C ABI) against the CPU oracle on identical inputs."""
import numpy as np
import pytest

import oracle_lib as O
from parity_util import assert_bits_equal, assert_close, compare_frames
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth
from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi

pytestmark = pytest.mark.gpu


def all_outputs(p):
    p = p.copy()
    p.outputs = abi.OUT_ALL
    return p


@pytest.fixture(scope="module")
def frames():
    return {c: synth.frame(c, 0) for c in (1, 2, 3)}


@pytest.mark.parametrize("config", [1, 2, 3])
def test_stage_crop(config, frames):
    p = synth.params(config)
    cloud = frames[config]
    with ObstacleProcessor(p, len(cloud)) as op:
        g_pts, g_idx = op.crop(cloud)
    o_pts, o_idx = O.crop(p, cloud)
    assert_bits_equal(g_idx, o_idx, "kept_idx")
    assert_bits_equal(g_pts, o_pts, "cropped points")


@pytest.mark.parametrize("config", [1, 2, 3])
def test_stage_voxel(config, frames):
    p = synth.params(config)
    cloud, _ = O.crop(p, frames[config])
    with ObstacleProcessor(p, len(cloud)) as op:
        g_pts, g_keys, g_w = op.downsample_cloud(cloud)
    o_pts, o_keys, o_w = O.voxel(p, cloud)
    assert g_w == o_w
    assert_bits_equal(g_keys, o_keys, "voxel keys")
    assert_bits_equal(g_pts, o_pts, "voxel centroids")


def test_stage_sor(frames):
    p = synth.params(1)
    cloud, _ = O.crop(p, frames[1])
    cloud, _, _ = O.voxel(p, cloud)
    with ObstacleProcessor(p, len(cloud)) as op:
        g_pts, g_idx, g_w = op.remove_statistical_outliers(cloud)
    o_pts, o_idx, o_w, dist, thr = O.sor(p, cloud)
    print("SOR margin min|d - thr| =", np.min(np.abs(dist.astype(np.float64) - thr)), "removed", len(cloud) - len(o_idx))
    assert g_w == o_w
    assert_bits_equal(g_idx, o_idx, "sor kept_idx")
    assert_bits_equal(g_pts, o_pts, "sor points")


@pytest.mark.parametrize("config", [1, 2, 3])
def test_stage_plane(config, frames):
    p = synth.params(config)
    cloud, _ = O.crop(p, frames[config])
    cloud, _, _ = O.voxel(p, cloud)
    with ObstacleProcessor(p, len(cloud)) as op:
        g = op.segment_plane_and_extract_indices(cloud)
    o = O.plane(p, cloud)
    assert g["n_passes"] == o["n_passes"]
    assert_bits_equal(g["pass_points"], o["pass_points"], "pass_points")
    assert_bits_equal(g["pass_inliers"], o["pass_inliers"], "pass_inliers")
    assert_close(g["pass_coeff"], o["pass_coeff"], "pass_coeff")
    assert_close(g["last_coeff"], o["last_coeff"], "last_coeff")
    assert_bits_equal(g["inliers"], o["inliers"], "last inliers")
    assert_bits_equal(g["src"], o["src"], "remaining src idx")
    assert_bits_equal(g["remaining"], o["remaining"], "remaining cloud")
    assert g["warnings"] == o["warnings"]


@pytest.mark.parametrize("config", [1, 2, 3])
def test_stage_cluster_and_centroid(config, frames):
    p = synth.params(config)
    of = O.process(p, frames[config])
    cloud = of.remaining_cloud
    with ObstacleProcessor(p, max(len(cloud), 1)) as op:
        g_off, g_idx = op.extract_euclidian_clusters(cloud)
        g_obs = op.centroid_radius(cloud, of.cluster_offsets, of.cluster_indices)
    assert_bits_equal(g_off, of.cluster_offsets, "cluster offsets")
    assert_bits_equal(g_idx, of.cluster_indices, "cluster indices")
    assert_close(g_obs, of.obstacles, "obstacles")


@pytest.mark.parametrize("config", [1, 2, 3, 4])
def test_pipeline(config, frames):
    p = all_outputs(synth.params(config))
    cloud = frames[config] if config in frames else synth.frame(config, 0)
    with ObstacleProcessor(p, len(cloud)) as op:
        g = op.process(cloud)
    o = O.process(p, cloud)
    compare_frames(g, o, p, f"config{config}: ")
    assert o.n_clusters > 0


def test_batch_mixed_sizes():
    """frames of different sizes (ragged batch), more frames than one wave"""
    p = all_outputs(synth.params(2))
    n = synth.points_per_frame(2)
    B = 5
    clouds = synth.frames(2, 10, B)
    counts = np.array([n, n - 1777, 50000, 0, 3], np.int32)
    with ObstacleProcessor(p, n, max_batch=2) as op:
        res = op.process_batch(clouds, counts)
    for f in range(B):
        o = O.process(p, clouds[f, :counts[f]])
        compare_frames(res[f], o, p, f"frame{f}: ")
