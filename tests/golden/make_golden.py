"""Regenerates tests/golden/pipeline_golden.json.

The reference cannot run here (needs ROS + PCL), and ships no golden vectors, so these fixtures are outputs of the
CPU oracle (oracle/pcop_oracle.cpp) on the synthetic frames of the BASELINE configs: counts plus CRC32 checksums of
every index array and (for float arrays) of the raw bit patterns.  They pin the oracle against silent drift and give
the GPU tests a second, file-based target.  Run: python tests/golden/make_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_lib as O  # noqa: E402
from pointcloud_obstacle_processing_b200 import synth  # noqa: E402


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def summarize(f):
    return {
        "counts": {k: int(getattr(f, k)) for k in ("n_input", "n_crop", "n_voxel", "n_sor", "n_remaining", "n_clusters",
                                                    "n_cluster_points", "n_plane_passes", "n_plane_inliers", "warnings")},
        "plane_pass_points": [int(x) for x in f.plane_pass_points],
        "plane_pass_inliers": [int(x) for x in f.plane_pass_inliers],
        "crc": {
            "crop_kept_idx": crc(f.crop_kept_idx), "voxel_keys": crc(f.voxel_keys),
            "voxel_centroids_bits": crc(f.voxel_centroids), "sor_kept_idx": crc(f.sor_kept_idx),
            "plane_inlier_idx": crc(f.plane_inlier_idx), "remaining_src_idx": crc(f.remaining_src_idx),
            "remaining_cloud_bits": crc(f.remaining_cloud), "cluster_offsets": crc(f.cluster_offsets),
            "cluster_indices": crc(f.cluster_indices),
        },
        "cluster_sizes": np.diff(f.cluster_offsets).astype(int).tolist(),
        "plane_pass_coeff": np.asarray(f.plane_pass_coeff, np.float64).round(7).tolist(),
        "obstacles": np.asarray(f.obstacles, np.float64).round(6).tolist(),
    }


def main():
    out = {"generator": "tests/golden/make_golden.py", "source": "CPU oracle (parity unpinned by the reference)",
           "cases": {}}
    for config, frame in ((1, 0), (1, 7), (2, 0), (2, 5), (3, 0)):
        f = O.process(synth.params(config), synth.frame(config, frame))
        out["cases"][f"config{config}_frame{frame}"] = summarize(f)
    with open(os.path.join(HERE, "pipeline_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
