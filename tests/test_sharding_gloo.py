"""Host-side multi-GPU logic on CPU: world_size-2 gloo.  Each rank produces the results of its frame shard
(here from the CPU oracle, standing in for its GPU) and rank 0 gathers; must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointcloud_obstacle_processing_b200 import sharding, synth

F = 5  # deliberately not divisible by 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frames_for(lo, hi):
    import oracle_lib as O
    p = synth.params(1)
    return [O.process(p, synth.frame(1, 100 + k)) for k in range(lo, hi)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(F, rank, world)
    mine = _frames_for(lo, hi)
    g = sharding.gather_results(mine, F)
    if rank == 0:
        q.put((g.n_clusters.tolist(), [o.tolist() for o in g.cluster_offsets],
               [i.tolist() for i in g.cluster_indices], [b.tolist() for b in g.obstacles]))
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    for total in (0, 1, 5, 8, 4096):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                got += list(range(lo, hi))
            assert got == list(range(total))


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _frames_for(0, F)
    assert got[0] == [f.n_clusters for f in ref]
    for k, f in enumerate(ref):
        assert got[1][k] == f.cluster_offsets.tolist()
        assert got[2][k] == f.cluster_indices.tolist()
        assert np.array_equal(np.array(got[3][k], np.float32).reshape(-1, 4), f.obstacles)


def test_gather_single_process():
    ref = _frames_for(0, 2)
    g = sharding.gather_results(ref, 2)
    assert g.n_clusters.tolist() == [f.n_clusters for f in ref]
    assert all(np.array_equal(a, f.cluster_indices) for a, f in zip(g.cluster_indices, ref))


def _raw_results(frames):
    """a ctypes pcop_frame_result array over oracle results (host pointers, as process_batch_raw returns them)"""
    import ctypes as C
    from pointcloud_obstacle_processing_b200._ctypes_abi import FrameResult
    res = (FrameResult * len(frames))()
    keep = []
    # arrays of consecutive frames adjacent in memory (as in the library's pinned result buffer) for the first two
    # frames, separate allocations for the others: the gather must stage both layouts
    for k, f in enumerate(frames):
        res[k].n_clusters = f.n_clusters
        res[k].n_cluster_points = f.n_cluster_points
    def place(field, arrays, ctype):
        joint = np.concatenate([np.ascontiguousarray(a).reshape(-1) for a in arrays[:2]]) if len(arrays) >= 2 else None
        o = 0
        for k, a in enumerate(arrays):
            a = np.ascontiguousarray(a).reshape(-1)
            if k < 2 and joint is not None:
                view = joint[o:o + a.size]
                o += a.size
                keep.append(joint)
            else:
                view = a.copy()
                keep.append(view)
            setattr(res[k], field, C.cast(view.ctypes.data, C.POINTER(ctype)) if view.size else C.cast(keep[0].ctypes.data, C.POINTER(ctype)))
    place("cluster_offsets", [f.cluster_offsets for f in frames], C.c_int32)
    place("cluster_indices", [f.cluster_indices for f in frames], C.c_int32)
    place("obstacles", [f.obstacles for f in frames], C.c_float)
    return res, keep


def _worker_rg(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = 3
    mine = _frames_for(rank * per, rank * per + per)
    rg = sharding.ResultGather(per, torch.device("cpu"), ints_per_frame=40000)
    for _ in range(3):  # several steps: the two exchange slots are reused
        res, keep = _raw_results(mine)
        rg.submit(res)
    g = rg.last()
    if rank == 0:
        q.put((g.n_clusters.tolist(), [o.tolist() for o in g.cluster_offsets],
               [i.tolist() for i in g.cluster_indices], [b.tolist() for b in g.obstacles]))
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


def test_result_gather_world2_gloo():
    """the per-step gather bench.py runs over NCCL (cluster_offsets, cluster_indices, obstacles of every frame to rank 0),
    here over gloo from raw ctypes result arrays"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rg, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _frames_for(0, 6)
    assert got[0] == [f.n_clusters for f in ref]
    for k, f in enumerate(ref):
        assert got[1][k] == f.cluster_offsets.tolist()
        assert got[2][k] == f.cluster_indices.tolist()
        assert np.array_equal(np.array(got[3][k], np.float32).reshape(-1, 4), f.obstacles)


def _worker_nccl(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pointcloud_obstacle_processing_b200 import ObstacleProcessor
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    per = 6
    p = synth.params(2)
    clouds = synth.frames(2, 40 + rank * per, per)
    rg = sharding.ResultGather(per, torch.device("cuda", rank))
    with ObstacleProcessor(p, clouds.shape[1], max_batch=per, device=rank) as op:
        for _ in range(3):
            res = op.process_batch_raw(clouds.ctypes.data, clouds.shape[1], np.full(per, clouds.shape[1], np.int32))
            rg.submit(res)
        g = rg.last()
    if rank == 0:
        q.put((g.n_clusters.tolist(), [o.tolist() for o in g.cluster_offsets],
               [i.tolist() for i in g.cluster_indices], [b.tolist() for b in g.obstacles]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_result_gather_world2_nccl_matches_single_process_results():
    """two ranks, one GPU each: every rank runs its frame shard through the CUDA library, rank 0 receives the CSR
    clusters and obstacle records of all frames over NCCL; must equal the oracle on the concatenated frames"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import oracle_lib as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_nccl, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    prm = synth.params(2)
    ref = [O.process(prm, synth.frame(2, 40 + k)) for k in range(12)]
    assert got[0] == [f.n_clusters for f in ref]
    for k, f in enumerate(ref):
        assert got[1][k] == f.cluster_offsets.tolist()
        assert got[2][k] == f.cluster_indices.tolist()
        np.testing.assert_allclose(np.array(got[3][k], np.float32).reshape(-1, 4), f.obstacles, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_result_gather_from_device_resident_results():
    """outputs | OUT_DEVICE: the result arrays stay in HBM and ResultGather stages them straight from the library's device
    result buffer (what bench.py's `value_results_left_in_hbm_gathered` runs at N > 1); one rank here, against the oracle"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import oracle_lib as O
    from pointcloud_obstacle_processing_b200 import ObstacleProcessor
    from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi
    per = 20  # (more than one wave per lane would need a larger batch; 20 frames = one wave)
    p = synth.params(2)
    p.outputs = abi.OUT_DEFAULT | abi.OUT_DEVICE
    clouds = synth.frames(2, 60, per)
    rg = sharding.ResultGather(per, torch.device("cuda", 0))
    with ObstacleProcessor(p, clouds.shape[1], max_batch=per, device=0) as op:
        for _ in range(3):
            res = op.process_batch_raw(clouds.ctypes.data, clouds.shape[1], np.full(per, clouds.shape[1], np.int32))
            rg.submit(res)
            rg.wait_staged()  # the device result buffer is reused by the next call
        g = rg.last()
    prm = synth.params(2)
    ref = [O.process(prm, synth.frame(2, 60 + k)) for k in range(per)]
    assert g.n_clusters.tolist() == [f.n_clusters for f in ref]
    for k, f in enumerate(ref):
        assert g.cluster_offsets[k].tolist() == f.cluster_offsets.tolist()
        assert g.cluster_indices[k].tolist() == f.cluster_indices.tolist()
        np.testing.assert_allclose(np.asarray(g.obstacles[k], np.float32).reshape(-1, 4), f.obstacles, rtol=1e-5, atol=1e-5)
