"""ctypes binding of oracle/_build/libpcop_oracle.so — TEST INFRASTRUCTURE ONLY (never imported by the package)."""
import ctypes as C
import os
import subprocess

import numpy as np

from pointcloud_obstacle_processing_b200._ctypes_abi import FrameResult, Params, MAX_PASSES
from pointcloud_obstacle_processing_b200.result import Frame

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "_build", "libpcop_oracle.so")
_lib = None

_fp = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_ROOT, "oracle", "pcop_oracle.cpp")
        if (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB_PATH)
        L.pcop_oracle_process.argtypes = [C.POINTER(Params), _fp, C.c_int32, C.POINTER(FrameResult)]
        L.pcop_oracle_free_result.argtypes = [C.POINTER(FrameResult)]
        L.pcop_oracle_radius2.restype = C.c_float
        L.pcop_oracle_radius2.argtypes = [C.c_float]
        L.pcop_oracle_inverse_leaf.restype = C.c_float
        L.pcop_oracle_inverse_leaf.argtypes = [C.c_float]
        for name in ("det_log", "det_sin", "det_cos"):
            f = getattr(L, "pcop_oracle_" + name)
            f.restype = C.c_double
            f.argtypes = [C.c_double]
        L.pcop_oracle_det_atan2_ypos.restype = C.c_double
        L.pcop_oracle_det_atan2_ypos.argtypes = [C.c_double, C.c_double]
        L.pcop_oracle_tree_sum.restype = C.c_double
        L.pcop_oracle_tree_sum.argtypes = [_fp, C.c_int32]
        _lib = L
    return _lib


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p)


def process(params: Params, cloud: np.ndarray) -> Frame:
    cloud, cp = _c(cloud, np.float32)
    r = FrameResult()
    st = lib().pcop_oracle_process(C.byref(params), cp, cloud.shape[0], C.byref(r))
    assert st == 0, st
    f = Frame.from_c(r)
    lib().pcop_oracle_free_result(C.byref(r))
    return f


def crop(params, cloud):
    cloud, cp = _c(cloud, np.float32)
    n = cloud.shape[0]
    out = np.empty((max(n, 1), 4), np.float32)
    kept = np.empty(max(n, 1), np.int32)
    m = C.c_int32()
    lib().pcop_oracle_crop(C.byref(params), cp, n, out.ctypes.data_as(_fp), kept.ctypes.data_as(_fp), C.byref(m))
    return out[:m.value].copy(), kept[:m.value].copy()


def voxel(params, cloud):
    cloud, cp = _c(cloud, np.float32)
    n = cloud.shape[0]
    out = np.empty((max(n, 1), 4), np.float32)
    keys = np.empty(max(n, 1), np.uint32)
    v, w = C.c_int32(), C.c_uint32()
    lib().pcop_oracle_voxel(C.byref(params), cp, n, out.ctypes.data_as(_fp), keys.ctypes.data_as(_fp), C.byref(v),
                            C.byref(w))
    return out[:v.value].copy(), keys[:v.value].copy(), w.value


def voxel_keys(params, cloud):
    cloud, cp = _c(cloud, np.float32)
    keys = np.empty(max(cloud.shape[0], 1), np.uint32)
    lib().pcop_oracle_voxel_keys(C.byref(params), cp, cloud.shape[0], keys.ctypes.data_as(_fp))
    return keys[:cloud.shape[0]].copy()


def sor(params, cloud):
    cloud, cp = _c(cloud, np.float32)
    n = cloud.shape[0]
    out = np.empty((max(n, 1), 4), np.float32)
    kept = np.empty(max(n, 1), np.int32)
    dist = np.empty(max(n, 1), np.float32)
    s, w, thr = C.c_int32(), C.c_uint32(), C.c_double()
    st = lib().pcop_oracle_sor(C.byref(params), cp, n, out.ctypes.data_as(_fp), kept.ctypes.data_as(_fp), C.byref(s),
                               C.byref(w), dist.ctypes.data_as(_fp), C.byref(thr))
    assert st == 0
    return out[:s.value].copy(), kept[:s.value].copy(), w.value, dist[:n].copy(), thr.value


def sor_distances_bruteforce(cloud, meanK):
    cloud, cp = _c(cloud, np.float32)
    dist = np.empty(cloud.shape[0], np.float32)
    st = lib().pcop_oracle_sor_distances_bruteforce(cp, cloud.shape[0], meanK, dist.ctypes.data_as(_fp))
    assert st == 0
    return dist


def plane(params, cloud):
    cloud, cp = _c(cloud, np.float32)
    n = cloud.shape[0]
    rem = np.empty((max(n, 1), 4), np.float32)
    src = np.empty(max(n, 1), np.int32)
    inl = np.empty(max(n, 1), np.int32)
    pp = np.zeros(MAX_PASSES, np.int32)
    pi = np.zeros(MAX_PASSES, np.int32)
    pc = np.zeros((MAX_PASSES, 4), np.float32)
    lc = np.zeros(4, np.float32)
    p, npass, ninl, w = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint32()
    lib().pcop_oracle_plane(C.byref(params), cp, n, rem.ctypes.data_as(_fp), src.ctypes.data_as(_fp), C.byref(p),
                            C.byref(npass), pp.ctypes.data_as(_fp), pi.ctypes.data_as(_fp), pc.ctypes.data_as(_fp),
                            lc.ctypes.data_as(_fp), inl.ctypes.data_as(_fp), C.byref(ninl), C.byref(w))
    return dict(remaining=rem[:p.value].copy(), src=src[:p.value].copy(), n_passes=npass.value, pass_points=pp,
                pass_inliers=pi, pass_coeff=pc, last_coeff=lc, inliers=inl[:ninl.value].copy(), warnings=w.value)


def _cluster(fn, params, cloud):
    cloud, cp = _c(cloud, np.float32)
    n = cloud.shape[0]
    offs = np.zeros(n + 2, np.int32)
    idx = np.zeros(n + 1, np.int32)
    c, l = C.c_int32(), C.c_int32()
    fn(C.byref(params), cp, n, offs.ctypes.data_as(_fp), idx.ctypes.data_as(_fp), C.byref(c), C.byref(l))
    return offs[:c.value + 1].copy(), idx[:l.value].copy()


def cluster(params, cloud):
    return _cluster(lib().pcop_oracle_cluster, params, cloud)


def cluster_bruteforce(params, cloud):
    return _cluster(lib().pcop_oracle_cluster_bruteforce, params, cloud)


def centroid_radius(cloud, offsets, indices):
    cloud, cp = _c(cloud, np.float32)
    offsets, op = _c(offsets, np.int32)
    indices, ip = _c(indices, np.int32)
    c = len(offsets) - 1
    out = np.zeros((max(c, 1), 4), np.float32)
    lib().pcop_oracle_centroid_radius(cp, cloud.shape[0], op, ip, c, out.ctypes.data_as(_fp))
    return out[:c].copy()


def rng_raw(seed, count):
    raw = np.empty(count, np.uint32)
    rnd = np.empty(count, np.int32)
    lib().pcop_oracle_rng_raw(C.c_uint32(seed), count, raw.ctypes.data_as(_fp), rnd.ctypes.data_as(_fp))
    return raw, rnd


def draw_samples(seed, n_points, n_samples):
    out = np.empty((n_samples, 3), np.int32)
    lib().pcop_oracle_draw_samples(C.c_uint32(seed), n_points, n_samples, out.ctypes.data_as(_fp))
    return out


def segment_once(params, cloud):
    cloud, cp = _c(cloud, np.float32)
    rc = np.zeros(4, np.float32)
    fc = np.zeros(4, np.float32)
    a, b, it = C.c_int32(), C.c_int32(), C.c_int32()
    st = lib().pcop_oracle_segment_once(C.byref(params), cp, cloud.shape[0], rc.ctypes.data_as(_fp),
                                        fc.ctypes.data_as(_fp), C.byref(a), C.byref(b), C.byref(it))
    return dict(ok=(st == 0), ransac_coeff=rc, refined_coeff=fc, n_ransac_inliers=a.value,
                n_refined_inliers=b.value, iterations=it.value)


def tree_sum(v):
    v, vp = _c(v, np.float64)
    return lib().pcop_oracle_tree_sum(vp, len(v))


def eigen33_smallest(m):
    m, mp = _c(m, np.float64)
    ev = C.c_double()
    vec = np.zeros(3, np.float64)
    lib().pcop_oracle_eigen33_smallest(mp, C.byref(ev), vec.ctypes.data_as(_fp))
    return ev.value, vec


def transform(cloud, m, is_dense=False):
    """pcl_ros::transformPointCloud restatement (od.cpp:696); m: 4x4 float, row-major"""
    cloud, cp = _c(cloud, np.float32)
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty_like(cloud)
    st = lib().pcop_oracle_transform(cp, cloud.shape[0], m.ctypes.data_as(_fp), 1 if is_dense else 0,
                                     out.ctypes.data_as(_fp))
    assert st == 0
    return out


def xyz_to_pointcloud2(cloud, point_step, off_x, off_y, off_z):
    """pcl::toROSMsg restatement (od.cpp:290-294)"""
    cloud, cp = _c(cloud, np.float32)
    n = cloud.shape[0]
    out = np.full(max(n * point_step, 1), 0xAB, np.uint8)
    st = lib().pcop_oracle_xyz_to_pointcloud2(cp, n, point_step, off_x, off_y, off_z, out.ctypes.data_as(_fp))
    assert st == 0, st
    return out[:n * point_step]


def pointcloud2_to_xyz(data, n_points, point_step, off_x, off_y, off_z):
    """pcl::fromPCLPointCloud2<PointXYZ> restatement (od.cpp:689)"""
    buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data, np.uint8)
    out = np.empty((max(n_points, 1), 4), np.float32)
    st = lib().pcop_oracle_pointcloud2_to_xyz(buf.ctypes.data_as(_fp), n_points, point_step, off_x, off_y, off_z,
                                              out.ctypes.data_as(_fp))
    assert st == 0, st
    return out[:n_points].copy()


def occupancy_grid(params, cloud):
    """build_initial_occupancy_grid_dataset restatement (od.cpp:175-269): (grid int8 [H, W], counts, row_avg)"""
    cloud, cp = _c(cloud, np.float32)
    w, h = C.c_int32(), C.c_int32()
    assert lib().pcop_oracle_occupancy_dims(C.byref(params), C.byref(w), C.byref(h)) == 0
    grid = np.empty((h.value, w.value), np.int8)
    counts = np.empty((h.value, w.value), np.int64)
    avg = np.empty(h.value, np.int64)
    st = lib().pcop_oracle_occupancy_grid(C.byref(params), cp, cloud.shape[0], grid.ctypes.data_as(_fp),
                                          counts.ctypes.data_as(_fp), avg.ctypes.data_as(_fp))
    assert st == 0, st
    return grid, counts, avg


def occupancy_shadows(params, grid, remaining, offsets, indices, world_to_sensor, sensor_to_world):
    """handle_shadow_casting per cluster + obstacle marks (od.cpp:584-672, 817-833): (grid, records [C, 6], warnings)"""
    grid = np.array(grid, dtype=np.int8, order="C", copy=True)
    cloud = np.ascontiguousarray(remaining, np.float32).reshape(-1, 4)
    off = np.ascontiguousarray(offsets, np.int32)
    idx = np.ascontiguousarray(indices, np.int32)
    nc = max(len(off) - 1, 0)
    ws = np.ascontiguousarray(world_to_sensor, np.float32).reshape(16)
    sw = np.ascontiguousarray(sensor_to_world, np.float32).reshape(16)
    rec = np.zeros((max(nc, 1), 6), np.int32)
    warn = C.c_uint32(0)
    L = lib()
    L.pcop_oracle_occupancy_shadows.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    st = L.pcop_oracle_occupancy_shadows(C.byref(params), cloud.ctypes.data_as(C.c_void_p), cloud.shape[0],
                                         off.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), nc,
                                         ws.ctypes.data_as(C.c_void_p), sw.ctypes.data_as(C.c_void_p),
                                         grid.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p), C.byref(warn))
    assert st == 0, st
    return grid, rec[:nc].copy(), warn.value
