"""Host-side multi-GPU logic on CPU: world_size-2 gloo.  Each rank produces the results of its frame shard
(here from the CPU oracle, standing in for its GPU) and rank 0 gathers; must equal the single-process result."""
import os
import socket

import numpy as np
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
