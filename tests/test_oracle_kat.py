"""CPU tests of the oracle: known-answer vectors pinned in SURVEY.md 8c (computed independently of this repo:
std::mt19937 stream, float bit patterns), independent numpy / scipy restatements, and O(n^2) brute force.
The reference ships no tests or golden vectors, so these are the strongest pins available ("parity unpinned")."""
import math

import numpy as np
import pytest

import oracle_lib as O
from pointcloud_obstacle_processing_b200 import synth
from pointcloud_obstacle_processing_b200._ctypes_abi import (WARN_PLANE_BREAK, WARN_SOR_TOO_FEW_POINTS,
                                                            WARN_VOXEL_OVERFLOW_FALLBACK)


def f32bits(x):
    return int(np.float32(x).view(np.uint32))


def cloud_of(xyz):
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    return np.concatenate([xyz, np.ones((len(xyz), 1), np.float32)], axis=1)


# ---- RNG / sampling known answers ---------------------------------
def test_mt19937_stream_kat():
    raw, rnd = O.rng_raw(12345, 6)
    assert raw.tolist() == [3992670690, 3823185381, 1358822685, 561383553, 789925284, 170765737]
    assert rnd.tolist() == [1996335345, 1911592690, 679411342, 280691776, 394962642, 85382868]


def test_mt19937_matches_numpy_generator():
    # numpy's legacy MT19937 seeded by init_genrand(seed) produces the same raw stream
    rs = np.random.RandomState(12345)
    ref = rs.randint(0, 2**32, size=2000, dtype=np.uint64).astype(np.uint32)
    raw, _ = O.rng_raw(12345, 2000)
    assert np.array_equal(raw, ref)


@pytest.mark.parametrize("n,expect", [
    (1000, [[345, 197, 888], [776, 197, 976], [855, 614, 458]]),
    (30000, [[15345, 26412, 16640], [11776, 25808, 8562], [29855, 17240, 1496]]),
    (120000, [[15345, 8621, 102666], [11776, 45934, 64292], [89855, 92869, 532]]),
])
def test_draw_index_sample_kat(n, expect):
    assert O.draw_samples(12345, n, 3).tolist() == expect


# ---- float bit patterns -----------------------------------------
@pytest.mark.parametrize("tol,bits", [(0.4, 0x3E23D70B), (0.05, 0x3B23D70B), (0.3, 0x3DB851EC)])
def test_radius2_bits(tol, bits):
    assert f32bits(O.lib().pcop_oracle_radius2(np.float32(tol))) == bits


@pytest.mark.parametrize("leaf,bits", [(0.015, 0x42855556), (0.1, 0x41200000), (0.02, 0x42480000)])
def test_inverse_leaf_bits(leaf, bits):
    assert f32bits(O.lib().pcop_oracle_inverse_leaf(np.float32(leaf))) == bits


# ---- deterministic elementary functions ---------------------------------------------------
def test_det_math_accuracy():
    L = O.lib()
    rng = np.random.default_rng(1)
    for x in np.concatenate([rng.uniform(1e-12, 1.0, 500), rng.uniform(1.0, 1e6, 200), [0.01, 0.5, 1.0, 2.0]]):
        assert abs(L.pcop_oracle_det_log(x) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
    for t in rng.uniform(0, math.pi / 3, 500):
        assert abs(L.pcop_oracle_det_sin(t) - math.sin(t)) < 3e-16
        assert abs(L.pcop_oracle_det_cos(t) - math.cos(t)) < 3e-16
    for y, x in zip(rng.uniform(0, 10, 500), rng.uniform(-10, 10, 500)):
        assert abs(L.pcop_oracle_det_atan2_ypos(y, x) - math.atan2(y, x)) < 1e-15


def test_tree_sum_matches_fsum_and_is_chunked():
    rng = np.random.default_rng(2)
    for n in (0, 1, 31, 256, 2047, 2048, 2049, 10000):
        v = rng.normal(size=n) * 1e3
        s = O.tree_sum(v)
        assert abs(s - math.fsum(v)) <= 1e-9 * max(1.0, np.abs(v).sum())
    # shape check: lanes of one chunk are added in xor-butterfly order, chunks sequentially
    v = np.zeros(4096)
    v[0], v[16], v[2048] = 1.0, 2.0 ** -53, 1.0
    assert O.tree_sum(v) == 2.0  # (1 + 2^-53) rounds to 1 inside the chunk; an exact sum would give 2 + 2^-53 -> 2.0 too
    w = np.full(2048, 0.1)
    lane = 0.0
    for _ in range(8):
        lane += 0.1
    part = lane
    for _ in range(5):
        part = part + part
    expect = part
    for _ in range(7):
        expect += part
    assert O.tree_sum(w) == expect


def test_eigen33_against_numpy():
    rng = np.random.default_rng(3)
    for _ in range(200):
        a = rng.normal(size=(3, 3)) * rng.uniform(0.01, 10)
        m = a @ a.T + np.diag(rng.uniform(1e-6, 1e-3, 3))
        ev, vec = O.eigen33_smallest(m.reshape(-1))
        w, v = np.linalg.eigh(m)
        assert abs(ev - w[0]) <= 1e-9 * w[2]
        if w[1] - w[0] > 1e-6 * w[2]:
            assert abs(abs(vec @ v[:, 0]) - 1.0) < 1e-6


# ---- crop (od.cpp:195-215) -----------------------------------------------------------------
def test_crop_literal_predicate():
    p = synth.params(1)
    nan = np.nan
    pts = cloud_of([
        [0.0, 0.0, -0.5],      # inclusive lower bounds: kept
        [4.5, 3.78, 0.25],     # inclusive upper bounds: kept
        [4.5000005, 1.0, 0.0], # just outside x
        [nan, 1.0, 0.0],       # NaN x: dropped
        [1.0, nan, 0.0],       # NaN y with valid x: kept (quirk: only x is NaN-tested)
        [1.0, 1.0, nan],       # NaN z with valid x: kept
        [1.0, -0.0001, 0.0],   # below y_min
        [2.0, 2.0, 0.3],       # above z_max
    ])
    out, kept = O.crop(p, pts)
    assert kept.tolist() == [0, 1, 4, 5]
    assert np.array_equal(out.view(np.uint32), pts[kept].view(np.uint32))
    out, kept = O.crop(p, np.zeros((0, 4), np.float32))
    assert len(kept) == 0


def test_crop_matches_numpy_on_synthetic():
    p = synth.params(1)
    c = synth.frame(1, 3)
    x, y, z = c[:, 0], c[:, 1], c[:, 2]
    with np.errstate(invalid="ignore"):
        drop = np.isnan(x) | (x < p.x_min) | (x > p.x_max) | (z < p.z_min) | (z > p.z_max) | (y < p.y_min) | (y > p.y_max)
    _, kept = O.crop(p, c)
    assert np.array_equal(kept, np.nonzero(~drop)[0])
    assert 0.4 * len(c) < len(kept)


# ---- VoxelGrid -------------------------------------------------------------------------------
def numpy_voxel_keys(p, c):
    inv = np.float32(1.0) / np.float32(p.downsample_size)
    mn = c[:, :3].min(axis=0)
    mx = c[:, :3].max(axis=0)
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div = (max_b - min_b + 1).astype(np.int64)
    ijk = (np.floor(c[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int32).astype(np.int64)
    return (ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]).astype(np.uint32), div


def test_voxel_extents_params_yaml():
    # the params.yaml crop box spans 301 x 253 x 51 voxels of 0.015 m
    p = synth.params(1)
    corners = cloud_of([[p.x_min, p.y_min, p.z_min], [p.x_max, p.y_max, p.z_max]])
    keys = O.voxel_keys(p, corners)
    _, div = numpy_voxel_keys(p, corners)
    assert div.tolist() == [301, 253, 51]
    assert keys.tolist() == [0, 300 + 252 * 301 + 50 * 301 * 253]


@pytest.mark.parametrize("config", [1, 2, 3])
def test_voxel_against_numpy(config):
    p = synth.params(config)
    c, _ = O.crop(p, synth.frame(config, 1))
    keys_np, _ = numpy_voxel_keys(p, c)
    assert np.array_equal(O.voxel_keys(p, c), keys_np)
    out, keys, w = O.voxel(p, c)
    assert w == 0
    uk, inverse, counts = np.unique(keys_np, return_inverse=True, return_counts=True)
    assert np.array_equal(keys, uk)  # ascending key order
    sums = np.zeros((len(uk), 3), np.float64)
    np.add.at(sums, inverse, c[:, :3].astype(np.float64))
    cen = sums / counts[:, None]
    np.testing.assert_allclose(out[:, :3], cen, rtol=1e-5, atol=1e-6)
    assert np.all(out[:, 3] == 1.0)
    # bit-exact check of the sequential float sum on the most populated voxels
    order = np.argsort(-counts)[:50]
    for v in order:
        s = np.zeros(3, np.float32)
        for i in np.nonzero(keys_np == uk[v])[0]:
            s = (s + c[i, :3]).astype(np.float32)
        assert np.array_equal((s / np.float32(counts[v])).view(np.uint32), out[v, :3].view(np.uint32))


def test_voxel_overflow_fallback_and_empty():
    p = synth.params(1)
    p.downsample_size = 1e-4
    c = cloud_of([[0, 0, 0], [100, 100, 100], [5, 5, 5]])
    out, keys, w = O.voxel(p, c)
    assert w == WARN_VOXEL_OVERFLOW_FALLBACK
    assert np.array_equal(out.view(np.uint32), c.view(np.uint32))
    out, keys, w = O.voxel(p, np.zeros((0, 4), np.float32))
    assert len(out) == 0 and w == 0


# ---- StatisticalOutlierRemoval -------------------------------------------------------------------
def test_sor_kdtree_equals_bruteforce_and_scipy():
    from scipy.spatial import cKDTree
    p = synth.params(1)
    rng = np.random.default_rng(5)
    c = cloud_of(np.concatenate([rng.normal(size=(1500, 3)) * 0.1, rng.uniform(-2, 2, size=(60, 3))]))
    out, kept, w, dist, thr = O.sor(p, c)
    assert w == 0
    assert np.array_equal(dist.view(np.uint32), O.sor_distances_bruteforce(c, 15).view(np.uint32))
    d, _ = cKDTree(c[:, :3].astype(np.float64)).query(c[:, :3].astype(np.float64), k=16)
    np.testing.assert_allclose(dist, d[:, 1:].mean(axis=1), rtol=2e-5)
    mean, std = dist.astype(np.float64).mean(), dist.astype(np.float64).std(ddof=1)
    assert abs(thr - (mean + 4.0 * std)) < 1e-6 * thr  # PCL's sq_sum uses float products: loose
    assert np.array_equal(kept, np.nonzero(~(dist.astype(np.float64) > thr))[0])
    assert 0 < len(c) - len(kept) < 60


def test_sor_too_few_points_passthrough():
    p = synth.params(1)
    c = cloud_of(np.random.default_rng(6).normal(size=(15, 3)))
    out, kept, w, _, _ = O.sor(p, c)
    assert w == WARN_SOR_TOO_FEW_POINTS and kept.tolist() == list(range(15))
    out, kept, w, _, _ = O.sor(p, np.zeros((0, 4), np.float32))
    assert w == 0 and len(kept) == 0


# ---- RANSAC plane loop -----------------------------------------------------------------------------
def plane_scene(rng, n_plane=4000, n_out=800, noise=0.005):
    xy = rng.uniform(-2, 2, size=(n_plane, 2))
    z = 0.1 * xy[:, 0] - 0.05 * xy[:, 1] + 0.3 + rng.normal(size=n_plane) * noise
    plane = np.column_stack([xy, z])
    outl = rng.uniform(-2, 2, size=(n_out, 3)) + [0, 0, 1.5]
    pts = np.concatenate([plane, outl])
    rng.shuffle(pts)
    return cloud_of(pts)


def test_segment_recovers_plane():
    p = synth.params(1)
    c = plane_scene(np.random.default_rng(7))
    s = O.segment_once(p, c)
    assert s["ok"]
    n = np.array([0.1, -0.05, -1.0])
    n /= np.linalg.norm(n)
    got = s["refined_coeff"][:3].astype(np.float64)
    assert abs(abs(got @ n) - 1.0) < 1e-4
    assert abs(abs(s["refined_coeff"][3]) - 0.3 / np.linalg.norm([0.1, -0.05, -1.0])) < 2e-3
    assert s["n_refined_inliers"] >= 3990
    # independent check of the refinement: numpy eigh of the inlier covariance
    d = np.abs(c[:, :3].astype(np.float64) @ s["ransac_coeff"][:3].astype(np.float64) + float(s["ransac_coeff"][3]))
    inl = c[d < float(np.float32(p.plane_segment_dist_thres)) - 1e-6, :3].astype(np.float64)
    w, v = np.linalg.eigh(np.cov(inl.T, bias=True))
    assert abs(abs(v[:, 0] @ got) - 1.0) < 1e-6


def test_ransac_first_hypothesis_uses_kat_sample():
    # with probability 0.99 and a 100 % inlier cloud RANSAC stops after the first sample (k < 1 after it)
    p = synth.params(1)
    p.optimize_coefficients = 0
    rng = np.random.default_rng(8)
    xy = rng.uniform(-1, 1, size=(1000, 2))
    c = cloud_of(np.column_stack([xy, np.zeros(1000)]))
    s = O.segment_once(p, c)
    assert s["iterations"] == 1 and s["n_ransac_inliers"] == 1000
    i0, i1, i2 = 345, 197, 888  # SURVEY 8a-4.3, n = 1000
    nrm = np.cross(c[i1, :3] - c[i0, :3], c[i2, :3] - c[i0, :3]).astype(np.float64)
    nrm /= np.linalg.norm(nrm)
    np.testing.assert_allclose(s["ransac_coeff"][:3], nrm, atol=1e-6)


def test_plane_loop_rule_and_break():
    p = synth.params(1)
    c = plane_scene(np.random.default_rng(9), n_plane=3000, n_out=2000)
    o = O.plane(p, c)
    assert o["n_passes"] >= 1
    assert o["pass_points"][0] == len(c)
    rem = len(c)
    for k in range(o["n_passes"]):
        assert rem > 0.3 * len(c)        # od.cpp:379: the pass only ran because > 30 % was left
        assert o["pass_points"][k] == rem
        rem -= o["pass_inliers"][k]
    assert rem == len(o["remaining"]) and (rem <= 0.3 * len(c) or o["warnings"] & WARN_PLANE_BREAK)
    assert np.array_equal(o["remaining"].view(np.uint32), c[o["src"]].view(np.uint32))
    assert np.all(np.diff(o["src"]) > 0)  # ExtractIndices keeps order
    # fewer than 3 points: segment() fails, loop breaks (od.cpp:383-387)
    o = O.plane(p, c[:2])
    assert o["n_passes"] == 0 and o["warnings"] & WARN_PLANE_BREAK and len(o["remaining"]) == 2
    o = O.plane(p, c[:0])
    assert o["n_passes"] == 0 and o["warnings"] == 0


# ---- Euclidean clustering --------------------------------------------------------------------------------
def random_cluster_scene(rng, n):
    centers = rng.uniform(-3, 3, size=(12, 3))
    pts = centers[rng.integers(0, 12, n)] + rng.normal(size=(n, 3)) * 0.15
    pts = np.concatenate([pts, rng.uniform(-4, 4, size=(n // 10, 3))])
    return cloud_of(pts)


@pytest.mark.parametrize("seed,tol", [(10, 0.05), (11, 0.1), (12, 0.2), (13, 0.4)])
def test_cluster_kdtree_equals_bruteforce(seed, tol):
    p = synth.params(1)
    p.euc_cluster_tolerance = tol
    p.euc_min_cluster_size = 3
    p.euc_max_cluster_size = 700
    c = random_cluster_scene(np.random.default_rng(seed), 1500)
    o1, i1 = O.cluster(p, c)
    o2, i2 = O.cluster_bruteforce(p, c)
    assert np.array_equal(o1, o2) and np.array_equal(i1, i2)
    sizes = np.diff(o1)
    assert np.all(sizes[:-1] >= sizes[1:]) and np.all(sizes >= 3) and np.all(sizes <= 700)
    for k in range(len(sizes)):  # indices ascending inside a cluster; ties ordered by smallest index
        seg = i1[o1[k]:o1[k + 1]]
        assert np.all(np.diff(seg) > 0)
        if k and sizes[k] == sizes[k - 1]:
            assert i1[o1[k - 1]] < seg[0]


def test_cluster_against_scipy_components():
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    p = synth.params(1)
    p.euc_cluster_tolerance = 0.12
    p.euc_min_cluster_size = 1
    p.euc_max_cluster_size = 10**6
    c = random_cluster_scene(np.random.default_rng(14), 3000)
    offs, idx = O.cluster(p, c)
    xyz = c[:, :3].astype(np.float64)
    pairs = cKDTree(xyz).sparse_distance_matrix(cKDTree(xyz), 0.12 * (1 - 1e-6), output_type="coo_matrix")
    ncomp, lab = connected_components(pairs, directed=False)
    assert ncomp == len(offs) - 1
    for k in range(len(offs) - 1):
        assert len(set(lab[idx[offs[k]:offs[k + 1]]])) == 1
    assert offs[-1] == len(c)


def test_cluster_oversize_dropped_whole():
    p = synth.params(1)
    p.euc_cluster_tolerance = 0.1
    p.euc_min_cluster_size = 2
    p.euc_max_cluster_size = 10
    line = np.column_stack([np.arange(30) * 0.05, np.zeros(30), np.zeros(30)])   # one 30-point chain: dropped
    pair = np.array([[10, 0, 0], [10.05, 0, 0]])                                 # kept
    single = np.array([[20, 0, 0]])                                              # below min: dropped
    offs, idx = O.cluster(p, cloud_of(np.concatenate([line, pair, single])))
    assert offs.tolist() == [0, 2] and idx.tolist() == [30, 31]


def test_centroid_radius_against_numpy():
    p = synth.params(2)
    o = O.process(p, synth.frame(2, 0))
    assert o.n_clusters > 5
    for k in range(o.n_clusters):
        m = o.remaining_cloud[o.cluster_indices[o.cluster_offsets[k]:o.cluster_offsets[k + 1]], :3].astype(np.float64)
        cen = m.mean(axis=0)
        np.testing.assert_allclose(o.obstacles[k, :3], cen, rtol=1e-6, atol=1e-6)
        assert abs(o.obstacles[k, 3] - np.sqrt(((m - cen) ** 2).sum(axis=1)).max()) < 1e-5


# ---- whole pipeline: stage chaining ------------------------------------------------------------------------
@pytest.mark.parametrize("config", [1, 2, 3])
def test_pipeline_equals_chained_stages(config):
    p = synth.params(config)
    c = synth.frame(config, 2)
    o = O.process(p, c)
    a, kept = O.crop(p, c)
    assert np.array_equal(kept, o.crop_kept_idx)
    b, keys, _ = O.voxel(p, a)
    assert np.array_equal(keys, o.voxel_keys)
    if p.enable_sor:
        b, sk, _, _, _ = O.sor(p, b)
        assert np.array_equal(sk, o.sor_kept_idx)
    pl = O.plane(p, b)
    assert np.array_equal(pl["remaining"].view(np.uint32), o.remaining_cloud.view(np.uint32))
    offs, idx = O.cluster(p, pl["remaining"])
    assert np.array_equal(offs, o.cluster_offsets) and np.array_equal(idx, o.cluster_indices)
    assert o.n_clusters >= 3


def _rigid(yaw, pitch, roll, t):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                  [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    m = np.eye(4)
    m[:3, :3] = R
    m[:3, 3] = t
    return m.astype(np.float32)


def test_transform_matches_pcl_coefficient_formula():
    """od.cpp:696 pcl_ros::transformPointCloud -> pcl::transformPointCloud(Matrix4f): out.x = m00*x + m01*y + m02*z +
    m03 in float, left to right; independent numpy float32 restatement, bit for bit"""
    rng = np.random.default_rng(5)
    pts = rng.uniform(-5, 5, size=(4000, 3)).astype(np.float32)
    cloud = np.concatenate([pts, np.ones((len(pts), 1), np.float32)], axis=1)
    m = _rigid(0.3, -0.2, 1.1, [0.5, -1.25, 0.75])
    out = O.transform(cloud, m, is_dense=True)
    f = np.float32
    for r in range(3):
        ref = (m[r, 0] * pts[:, 0]).astype(f)
        ref = (ref + (m[r, 1] * pts[:, 1]).astype(f)).astype(f)
        ref = (ref + (m[r, 2] * pts[:, 2]).astype(f)).astype(f)
        ref = (ref + m[r, 3]).astype(f)
        assert np.array_equal(out[:, r].view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(out[:, 3], cloud[:, 3])
    # identity leaves every bit alone
    assert np.array_equal(O.transform(cloud, np.eye(4, dtype=np.float32), is_dense=True).view(np.uint32), cloud.view(np.uint32))


def test_transform_non_dense_cloud_copies_non_finite_points():
    """PCL skips points with a non-finite coordinate when cloud.is_dense is false (Kinect clouds), transforms them
    (-> NaN) when it is true"""
    cloud = np.array([[1, 2, 3, 1], [np.nan, 2, 3, 1], [1, np.inf, 3, 1], [4, 5, 6, 1]], np.float32)
    m = _rigid(0.5, 0.1, -0.3, [1, 2, 3])
    out = O.transform(cloud, m, is_dense=False)
    assert np.array_equal(out[1].view(np.uint32), cloud[1].view(np.uint32))
    assert np.array_equal(out[2].view(np.uint32), cloud[2].view(np.uint32))
    assert not np.array_equal(out[0], cloud[0]) and np.isfinite(out[[0, 3]]).all()
    dense = O.transform(cloud, m, is_dense=True)
    assert np.isnan(dense[1, :3]).all() and not np.isfinite(dense[2, :3]).all()


def _make_pointcloud2(xyz, point_step, offs, seed=3):
    """a sensor_msgs/PointCloud2-like payload: FLOAT32 x, y, z at `offs` inside point_step-byte records, other bytes
    random (rgb, intensity, padding)"""
    n = len(xyz)
    rng = np.random.default_rng(seed)
    buf = rng.integers(0, 256, size=(n, point_step), dtype=np.uint8)
    for a, off in enumerate(offs):
        buf[:, off:off + 4] = np.ascontiguousarray(xyz[:, a], np.float32).view(np.uint8).reshape(n, 4)
    return buf.reshape(-1)


@pytest.mark.parametrize("point_step,offs", [(16, (0, 4, 8)), (32, (0, 4, 8)), (22, (1, 9, 14)), (12, (8, 0, 4))])
def test_pointcloud2_field_extraction(point_step, offs):
    """od.cpp:688-689 toPCL + fromPCLPointCloud2<PointXYZ>: strided FLOAT32 field copy, padding float 1.0f, bits kept
    (NaN payloads included)"""
    rng = np.random.default_rng(11)
    xyz = rng.normal(size=(1000, 3)).astype(np.float32)
    xyz[5] = np.nan
    xyz[6, 1] = np.inf
    buf = _make_pointcloud2(xyz, point_step, offs)
    out = O.pointcloud2_to_xyz(buf, len(xyz), point_step, *offs)
    assert np.array_equal(out[:, :3].view(np.uint32), xyz.view(np.uint32))
    assert (out[:, 3] == 1.0).all()


def test_occupancy_grid_dims_params_yaml():
    """od.cpp:958-959 with params.yaml: 101 x 120 cells (SURVEY 8c known answer)"""
    import ctypes as C
    p = synth.params(1)
    w, h = C.c_int32(), C.c_int32()
    assert O.lib().pcop_oracle_occupancy_dims(C.byref(p), C.byref(w), C.byref(h)) == 0
    assert (w.value, h.value) == (101, 120)


def test_occupancy_grid_against_literal_python_loops():
    """od.cpp:134-157 + 175-269 restated independently in Python with numpy float32 scalars (the while-loops as written)"""
    p = synth.params(1)
    f = np.float32
    rng = np.random.default_rng(17)
    n = 3000
    pts = np.stack([rng.uniform(-0.2, 4.7, n), rng.uniform(-0.2, 4.0, n), rng.uniform(-0.6, 0.3, n)], 1).astype(f)
    pts[::97, 0] = np.nan
    pts[5] = [p.x_max, p.y_max, 0.0]   # corner cases on the box faces
    pts[6] = [p.x_min, p.y_min, 0.0]
    cloud = np.concatenate([pts, np.ones((n, 1), f)], 1)
    grid, counts, avg = O.occupancy_grid(p, cloud)
    H, W = grid.shape
    bs, x_min, x_max, y_min, y_max = f(p.block_size), f(p.x_min), f(p.x_max), f(p.y_min), f(p.y_max)
    ref = np.zeros(H * W, np.int64)
    for x, y, z in pts:
        if np.isnan(x) or x < x_min or x > x_max or z < f(p.z_min) or z > f(p.z_max) or y < y_min or y > y_max:
            continue
        xc = 0
        while f(y_min + f(f(xc + 1) * bs)) < y:
            xc += 1
        yc = 0
        while f(x_max - f(f(yc + 1) * bs)) > x:
            yc += 1
        idx = yc * W + xc
        if idx < H * W:
            ref[idx] += 1
    assert np.array_equal(counts.ravel(), ref)
    ravg = ref.reshape(H, W).sum(1) // W
    assert np.array_equal(avg, ravg)
    thr = (ravg.astype(f) * f(f(1.0) - f(p.dev_percent))).astype(f)
    want = np.where(ref.reshape(H, W).astype(f) < thr[:, None], 100, 0).astype(np.int8)
    assert np.array_equal(grid, want)
    assert counts.sum() > 1000


# ---- occupancy grid: shadow casting + obstacle marks (od.cpp:467-672, 817-833) ------------------------------------
from shadow_util import rigid as rigid_pair, shadow_scene  # noqa: E402


def literal_shadows(p, grid, cloud, offsets, indices, ws, sw):
    """handle_shadow_casting / calculate_shadow_cast / traceShadow / the marking loop restated from od.cpp with numpy
    float32 scalars and libm's asin / tan (double overloads), independent of oracle/pcop_oracle.cpp"""
    import math
    f = np.float32
    H, W = grid.shape
    g = grid.copy().ravel()
    size = H * W
    bs, y_min, x_max = f(p.block_size), f(p.y_min), f(p.x_max)

    def xf(m, q):
        return [f(f(f(f(m[r, 0] * q[0]) + f(m[r, 1] * q[1])) + f(m[r, 2] * q[2])) + m[r, 3]) for r in range(3)]

    def grid_xy(x, y, x_mn, y_mx):
        xc = 0
        while f(x_mn + f(f(xc + 1) * bs)) < x:
            xc += 1
        yc = 0
        while f(y_mx - f(f(yc + 1) * bs)) > y:
            yc += 1
        return xc, yc

    def trace(v1, v2):
        x0, x1, y0, y1 = int(v1[0]), int(v2[0]), int(v1[1]), int(v2[1])
        steep = abs(y1 - y0) > abs(x1 - x0)
        if steep:
            x0, y0, x1, y1 = y0, x0, y1, x1
        if x0 > x1:
            x0, x1, y0, y1 = x1, x0, y1, y0
        dx, dy = f(x1 - x0), f(y1 - y0)
        with np.errstate(divide="ignore", invalid="ignore"):
            gradient = f(dy / dx)
        if dx == 0.0:
            gradient = f(1)
        iy = f(y0)
        for x in range(x0, x1 + 1):
            fl = int(math.floor(iy))
            idx = x * W + fl if steep else fl * W + x
            for k in (idx, idx + 1):
                if -1 < k < size:
                    g[k] = p.grid_opacity
            iy = f(iy + gradient)

    records = []
    for c in range(len(offsets) - 1):
        mem = indices[offsets[c]:offsets[c + 1]]
        if len(mem) < 2:
            records.append([0] * 6)
            continue
        tp = [xf(ws, cloud[i]) for i in mem]
        vmin, vmax, hmin, hmax = tp[0], tp[0][0], tp[0][1], tp[0][1]
        for q in tp[1:]:
            if q[0] < vmin[0]:
                vmin = q
            if q[0] > vmax:
                vmax = q[0]
            if q[1] < hmin:
                hmin = q[1]
            if q[1] > hmax:
                hmax = q[1]
        width = f(abs(f(hmax - hmin)))
        a, b = vmin[2], f(abs(vmin[0]))
        cc = f(math.sqrt(f(f(a * a) + f(b * b))))
        e = f(abs(float(vmax)) - abs(float(vmin[0])) + 0.04)
        D = f(math.asin(f(a / cc)))
        d = f(math.tan(float(D)) * float(e) + 0.25)
        v_len = f(math.sqrt(f(f(f(vmin[0] * vmin[0]) + f(vmin[1] * vmin[1])) + f(vmin[2] * vmin[2]))))
        end = [f(f(f(vmin[k] / v_len) * d) + vmin[k]) for k in range(3)]
        we = xf(sw, end)
        ex, ey = grid_xy(we[1], we[0], y_min, x_max)
        wst = xf(sw, vmin)
        sx, sy = grid_xy(wst[1], wst[0], y_min, x_max)
        wb = f(width / bs)
        shift = math.ceil(f(wb / f(2)))
        sx, ex = int(sx + shift), int(ex + shift)
        n_lines = int(math.ceil(wb) + 3)
        records.append([sx, sy, ex, ey, n_lines, 0])
        for i in range(n_lines):
            trace((f(sx - i), f(sy)), (f(ex - i), f(ey)))
    for q in cloud:
        if np.isnan(q[0]):
            continue
        xc, yc = grid_xy(q[1], q[0], y_min, x_max)
        if yc * W + xc < size:
            g[yc * W + xc] = 100
    return g.reshape(H, W), np.array(records, np.int32).reshape(-1, 6)


@pytest.mark.parametrize("seed,opacity", [(1, 0), (2, 50), (3, 77)])
def test_occupancy_shadows_against_literal_python(seed, opacity):
    """the oracle's shadow casting against an independent literal restatement of od.cpp:467-672, 817-833"""
    p = synth.params(1)
    p.grid_opacity = opacity
    cloud, offsets, indices = shadow_scene(seed)
    # sensor above and behind the arena, pitched down (any rigid pose will do: the matrices are inputs)
    sw, ws = rigid_pair(0.3 * seed, 0.5, 0.1, [5.2, 1.9, 1.1])
    grid0, _, _ = O.occupancy_grid(p, cloud)
    grid, rec, warn = O.occupancy_shadows(p, grid0, cloud, offsets, indices, ws, sw)
    want, wrec = literal_shadows(p, grid0, cloud, offsets, indices, ws, sw)
    assert warn == 0
    assert np.array_equal(rec, wrec)
    assert np.array_equal(grid, want)
    assert (grid == 100).sum() > 100 and (rec[:, 4] > 0).sum() >= 3
    if opacity:
        assert (grid == opacity).sum() > 0  # some shadow cells survive the obstacle marks


def test_occupancy_shadows_degenerate_inputs():
    """NaN members, a far-away pose (cell search cap), no clusters: defined results, no hang"""
    p = synth.params(1)
    p.grid_opacity = 33
    cloud, offsets, indices = shadow_scene(5)
    grid0, _, _ = O.occupancy_grid(p, cloud)
    sw, ws = rigid_pair(0.2, 0.4, 0.0, [5.0, 2.0, 1.0])
    g_none, rec, warn = O.occupancy_shadows(p, grid0, cloud, np.zeros(1, np.int32), np.zeros(0, np.int32), ws, sw)
    assert rec.shape == (0, 6) and warn == 0 and (g_none == 100).sum() > 0 and (g_none == 33).sum() == 0
    bad = cloud.copy()
    bad[indices[offsets[0]], 1] = np.nan  # first member of the largest cluster: its NaN y is never replaced (od.cpp:608-620)
    g_nan, rec_nan, warn = O.occupancy_shadows(p, grid0, bad, offsets, indices, ws, sw)
    assert rec_nan[0, 4] == 0  # width = NaN -> the line-count comparison is false at once (od.cpp:645)
    far_sw, _ = rigid_pair(0.0, 0.0, 0.0, [-3.0e6, 3.0e6, 0.0])  # both cell searches run into the 2^20 step cap
    g_far, rec_far, warn = O.occupancy_shadows(p, grid0, cloud, offsets, indices, np.eye(4, dtype=np.float32), far_sw)
    assert (rec_far[rec_far[:, 4] > 0, 1] == 1 << 20).all() and (g_far == 33).sum() == 0
    # a member almost on the sensor's z axis: asin -> pi/2, tan explodes, the end point lands ~1e6 cells away:
    # the fan is not drawn and the warning is raised
    eye = np.eye(4, dtype=np.float32)
    tall = np.array([[1e-6, -1.0, 1.0, 1.0], [0.5, 1.2, 0.2, 1.0], [0.6, 1.1, 0.1, 1.0]], np.float32)
    g_t, rec_t, warn = O.occupancy_shadows(p, grid0, tall, np.array([0, 3], np.int32), np.arange(3, dtype=np.int32), eye, eye)
    assert warn == 16 and rec_t[0, 5] == 1 and (g_t == 33).sum() == 0


def test_det_asin_tan_against_libm():
    import ctypes as C
    import math
    L = O.lib()
    for name in ("det_asin", "det_tan"):
        fn = getattr(L, "pcop_oracle_" + name)
        fn.restype = C.c_double
        fn.argtypes = [C.c_double]
    rng = np.random.default_rng(9)
    for q in np.concatenate([rng.uniform(-1, 1, 2000), [0.0, 1.0, -1.0, 1e-9, 0.999999, -0.5]]):
        assert abs(L.pcop_oracle_det_asin(q) - math.asin(q)) <= 4e-16 * max(1.0, abs(math.asin(q)))
    assert math.isnan(L.pcop_oracle_det_asin(1.0000001)) and math.isnan(L.pcop_oracle_det_asin(float("nan")))
    for x in np.concatenate([rng.uniform(-1.5, 1.5, 2000), [0.0, 0.7, -1.2]]):
        assert abs(L.pcop_oracle_det_tan(x) - math.tan(x)) <= 1e-14 * max(1.0, abs(math.tan(x)))


def test_pointcloud2_egress_layout_and_round_trip():
    """pcl::toROSMsg of a PointXYZ cloud (od.cpp:290-294): the reference layout (16, 0, 4, 8) is the record array
    itself (padding float included); other layouts carry the three fields and zeros, and the ingest restatement
    (od.cpp:689) reads them back bit for bit"""
    rng = np.random.default_rng(17)
    cloud = rng.normal(size=(257, 4)).astype(np.float32)
    cloud[3, 1] = np.nan
    cloud[:, 3] = 1.0
    assert O.xyz_to_pointcloud2(cloud, 16, 0, 4, 8).tobytes() == cloud.tobytes()
    for step, offs in ((32, (0, 4, 8)), (22, (1, 9, 14)), (12, (8, 0, 4))):
        data = O.xyz_to_pointcloud2(cloud, step, *offs)
        assert data.shape == (257 * step,)
        rec = data.reshape(257, step)
        covered = np.zeros(step, bool)
        for o in offs:
            covered[o:o + 4] = True
        assert not rec[:, ~covered].any()
        back = O.pointcloud2_to_xyz(data, 257, step, *offs)
        assert back.view(np.uint32).tolist() == cloud.view(np.uint32).tolist()
