"""CPU cross-checks of the oracle on the random scenes of tests/test_gpu_fuzz.py: the crop predicate against numpy,
VoxelGrid keys / order / sequential-sum centroids against an independent numpy restatement, Euclidean clustering
(k-d tree) against the O(n^2) exact-predicate brute force, the pipeline against its chained stages.  Together with
the GPU sweep this ties the CUDA path, the oracle and the independent restatements on the same inputs."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_fuzz import fuzz_case  # scene generator only (the module's GPU tests are deselected on a CPU box)
from test_oracle_kat import numpy_voxel_keys

SEEDS = [s for s in range(64)]


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_stages_on_fuzz_scene(seed):
    p, cloud = fuzz_case(seed)
    fr = O.process(p, cloud)
    # crop: literal predicate (od.cpp:197-199), numpy restatement
    x, y, z = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    with np.errstate(invalid="ignore"):
        drop = np.isnan(x) | (x < np.float32(p.x_min)) | (x > np.float32(p.x_max)) | (z < np.float32(p.z_min)) | \
            (z > np.float32(p.z_max)) | (y < np.float32(p.y_min)) | (y > np.float32(p.y_max))
    kept = np.nonzero(~drop)[0] if p.enable_crop else np.arange(len(cloud))
    assert fr.n_crop == len(kept)
    if p.enable_crop:
        assert np.array_equal(fr.crop_kept_idx, kept)
    c = cloud[kept]
    finite = np.isfinite(c[:, :3]).all() if len(c) else True
    # voxel: keys, ascending order, sequential float sums (only where every coordinate is finite: the key arithmetic
    # of non-finite points is the oracle's cvttss2si choice, covered by the GPU sweep)
    if p.enable_voxel and len(c) and finite and not (fr.warnings & 1):
        keys_np, _ = numpy_voxel_keys(p, c)
        uk, counts = np.unique(keys_np, return_counts=True)
        assert fr.n_voxel == len(uk) and np.array_equal(fr.voxel_keys, uk)
        for v in np.argsort(-counts)[:20]:
            s = np.zeros(3, np.float32)
            for i in np.nonzero(keys_np == uk[v])[0]:
                s = (s + c[i, :3]).astype(np.float32)
            assert np.array_equal((s / np.float32(counts[v])).view(np.uint32), fr.voxel_centroids[v, :3].view(np.uint32))
    # clustering: k-d tree path against the exact-predicate brute force on the cloud the indices refer to
    rem = fr.remaining_cloud
    if 0 < len(rem) <= 3000:
        o2, i2 = O.cluster_bruteforce(p, rem)
        assert np.array_equal(fr.cluster_offsets, o2) and np.array_equal(fr.cluster_indices, i2)
        sizes = np.diff(fr.cluster_offsets)
        assert np.all(sizes >= p.euc_min_cluster_size) and np.all(sizes <= p.euc_max_cluster_size)
        assert np.all(sizes[:-1] >= sizes[1:])
    # remaining cloud = voxel (or crop) cloud minus the plane inliers, original order kept
    assert np.all(np.diff(fr.remaining_src_idx) > 0) if len(rem) > 1 else True
