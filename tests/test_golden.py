"""Committed golden fixtures (tests/golden/pipeline_golden.json, made by tests/golden/make_golden.py from the CPU
oracle) checked against the oracle on CPU and against the CUDA path on the GPU."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from golden.make_golden import summarize
from pointcloud_obstacle_processing_b200 import synth
from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pipeline_golden.json")))
CASES = sorted(GOLD["cases"])


def _parse(case):
    c, f = case.split("_")
    return int(c[len("config"):]), int(f[len("frame"):])


def _check(summary, gold):
    assert summary["counts"] == gold["counts"]
    assert summary["plane_pass_points"] == gold["plane_pass_points"]
    assert summary["plane_pass_inliers"] == gold["plane_pass_inliers"]
    assert summary["cluster_sizes"] == gold["cluster_sizes"]
    assert summary["crc"] == gold["crc"]
    np.testing.assert_allclose(summary["plane_pass_coeff"], gold["plane_pass_coeff"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(summary["obstacles"], gold["obstacles"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(case):
    config, frame = _parse(case)
    _check(summarize(O.process(synth.params(config), synth.frame(config, frame))), GOLD["cases"][case])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_golden(case):
    from pointcloud_obstacle_processing_b200 import ObstacleProcessor
    config, frame = _parse(case)
    p = synth.params(config)
    p.outputs = abi.OUT_ALL
    cloud = synth.frame(config, frame)
    with ObstacleProcessor(p, len(cloud)) as op:
        g = op.process(cloud)
    if g.sor_kept_idx is None:  # stage disabled: the oracle reports the identity
        g.sor_kept_idx = np.arange(g.n_sor, dtype=np.int32)
    _check(summarize(g), GOLD["cases"][case])
