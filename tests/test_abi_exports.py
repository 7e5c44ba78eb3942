"""CPU checks of the drop-in boundary: libpcop.so loads without a GPU, exports every symbol include/pcop.h declares,
struct layouts match the header, parameter initialisers carry the reference's values, and nothing computes without a
CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import pointcloud_obstacle_processing_b200 as pkg
from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcop.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcop_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = pkg.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pcop.h but not exported by libpcop.so"
    assert set(pkg.api.EXPORTS) <= set(names)
    assert lib.pcop_abi_version() == 1


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "pcop.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(pcop_params),sizeof(pcop_frame_result),offsetof(pcop_params,plane_keep_fraction),'
                   'offsetof(pcop_frame_result,crop_kept_idx),offsetof(pcop_frame_result,obstacles));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(abi.Params), C.sizeof(abi.FrameResult), abi.Params.plane_keep_fraction.offset,
                   abi.FrameResult.crop_kept_idx.offset, abi.FrameResult.obstacles.offset]


def test_params_yaml_values():
    p = pkg.params_yaml()  # minibot_cr18/params.yaml:2-31
    assert (p.x_min, p.x_max, p.y_min, round(p.y_max, 5), p.z_min, p.z_max) == (0.0, 4.5, 0.0, 3.78, -0.5, 0.25)
    assert round(p.downsample_size, 6) == 0.015 and p.statistical_outlier_meanK == 15
    assert p.statistical_outlier_stdDevThres == 4.0 and round(p.plane_segment_dist_thres, 6) == 0.04
    assert p.plane_segment_angle == 20 and round(p.euc_cluster_tolerance, 6) == 0.4
    assert (p.euc_min_cluster_size, p.euc_max_cluster_size, p.accumulate_count) == (5, 20000, 200)
    d = pkg.params_code_defaults()  # od.cpp:940-975
    assert (d.x_min, d.x_max, d.z_min, d.z_max, d.accumulate_count) == (-1.0, 1.0, 0.0, -0.5, 2)
    assert d.statistical_outlier_stdDevThres == 1.0 and d.plane_keep_fraction == 0.3 and d.ransac_seed == 12345
    assert list(d.plane_axis) == [0.0, 0.0, 1.0] and d.plane_max_iterations == 50 and d.plane_probability == 0.99


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.PcopError) as e:
        pkg.ObstacleProcessor(pkg.params_yaml(), 1000)
    assert e.value.status == abi.ERR_CUDA


def test_package_does_not_import_oracle():
    """the product never includes, imports, links or loads anything under oracle/ (comments may mention it)"""
    pkg_dir = os.path.join(ROOT, "pointcloud_obstacle_processing_b200")
    bad = re.compile(r'#include\s*[<"][^>"]*oracle|import\s+oracle|from\s+oracle|oracle_lib|libpcop_oracle|pcop_oracle_\w+\s*\(')
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(text), f"{f} uses the oracle"
    out = subprocess.run(["ldd", pkg.api.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_cpp_host_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """examples/pcop_host_demo.cpp (the C++ shim's shape): pcop.h is valid C++11 under -Wall -Wextra -pedantic -Werror,
    the program links against libpcop.so, and on a box without a CUDA device it stops with the no-fallback error"""
    import torch
    exe = tmp_path / "demo"
    pkg_dir = os.path.join(ROOT, "pointcloud_obstacle_processing_b200")
    pkg.load_library()  # (builds libpcop.so if missing)
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "pcop_host_demo.cpp"), "-L", pkg_dir, "-lpcop", "-o", str(exe)])
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    env = dict(os.environ, LD_LIBRARY_PATH=pkg_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def _build_cpp(tmp_path, name):
    exe = tmp_path / name
    pkg_dir = os.path.join(ROOT, "pointcloud_obstacle_processing_b200")
    pkg.load_library()  # (builds libpcop.so if missing)
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".cpp"), "-L", pkg_dir, "-lpcop", "-o", str(exe)])
    env = dict(os.environ, LD_LIBRARY_PATH=pkg_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return exe, env


def test_cpp_digest_host_compiles(tmp_path):
    _build_cpp(tmp_path, "pcop_host_digest")


@pytest.mark.gpu
@pytest.mark.parametrize("config,frames", [(1, (0, 7)), (2, (0, 5))])
def test_cpp_host_on_the_gpu_matches_the_golden_fixtures(tmp_path, config, frames):
    """the node's host side is C++: a pure C++ caller of the C ABI (examples/pcop_host_digest.cpp, no Python between it and
    libpcop.so) processes recorded frames on the GPU -- frame by frame and as one batch -- and the digests of every result
    array must equal the oracle's golden fixtures: counts and CRC-32 of every index array and of the float arrays' bits"""
    import json
    import numpy as np
    from pointcloud_obstacle_processing_b200 import synth
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "pipeline_golden.json")))["cases"]
    exe, env = _build_cpp(tmp_path, "pcop_host_digest")
    n = synth.points_per_frame(config)
    clouds = np.stack([synth.frame(config, f) for f in frames])
    path = tmp_path / "frames.bin"
    clouds.astype(np.float32).tofile(path)
    r = subprocess.run([str(exe), str(path), str(n), str(len(frames)), str(config)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == 2 * len(frames)
    crc_keys = ("crop_kept_idx", "voxel_keys", "voxel_centroids_bits", "sor_kept_idx", "plane_inlier_idx", "remaining_src_idx",
                "remaining_cloud_bits", "cluster_offsets", "cluster_indices")
    cnt_keys = ("n_input", "n_crop", "n_voxel", "n_sor", "n_remaining", "n_clusters", "n_cluster_points", "n_plane_passes",
                "n_plane_inliers", "warnings")
    for l in lines:
        g = gold[f"config{config}_frame{frames[int(l[1])]}"]
        counts = [int(x) for x in l[4:14]]
        assert counts == [g["counts"][k] for k in cnt_keys], (l[:3], counts)
        crcs = [int(x) for x in l[16:25]]
        want = [g["crc"][k] for k in crc_keys]
        if config == 2:  # SOR disabled: the library returns no array, the oracle reports the identity
            crcs[3] = want[3]
        assert crcs == want, (l[:3], crcs, want)
        obs = np.array([float(x) for x in l[27:]], np.float64).reshape(-1, 4)
        np.testing.assert_allclose(obs, np.array(g["obstacles"]).reshape(-1, 4), rtol=1e-5, atol=1e-5)
