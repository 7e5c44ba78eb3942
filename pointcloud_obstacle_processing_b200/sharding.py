"""Multi-GPU plumbing: frames are independent, so a batch is split into contiguous frame blocks, one per rank
(one process per GPU), with no collective inside a frame.  The only exchange is the final result gather
(SURVEY 8e): per-frame cluster counts first, then the variable-length CSR / obstacle payloads, padded to the
longest rank.  Works on any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_frames: int, rank: int, world: int):
    """Contiguous block of ceil(F/G) frames for `rank` (the last ranks may get fewer or none)."""
    per = -(-total_frames // world) if world > 0 else total_frames
    lo = min(rank * per, total_frames)
    hi = min(lo + per, total_frames)
    return lo, hi


@dataclass
class GatheredResults:
    """What rank 0 holds after the gather: per global frame, its clusters (CSR) and obstacle records."""
    n_clusters: np.ndarray        # [F]
    n_cluster_points: np.ndarray  # [F]
    cluster_offsets: List[np.ndarray]
    cluster_indices: List[np.ndarray]
    obstacles: List[np.ndarray]


def _pad_cat(arrays: Sequence[np.ndarray], dtype, cols=None):
    if arrays:
        a = np.concatenate([np.asarray(x, dtype=dtype).reshape(-1) for x in arrays])
    else:
        a = np.zeros(0, dtype)
    return a


def gather_results(frames, total_frames: int, device: Optional[torch.device] = None, group=None, dst: int = 0):
    """frames: this rank's list of result.Frame (its shard, in frame order).  Returns GatheredResults on `dst`,
    None elsewhere.  Two rounds: all_gather of fixed-size per-rank headers, then all_gather of padded payloads."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    device = device or torch.device("cpu")
    nloc = len(frames)
    for f in frames:
        if f.cluster_offsets is None or f.cluster_indices is None or f.obstacles is None:
            raise ValueError("gather_results needs frames processed with outputs | OUT_CLUSTERS | OUT_OBSTACLES")
    c = np.array([f.n_clusters for f in frames], np.int64)
    l = np.array([f.n_cluster_points for f in frames], np.int64)
    offs = _pad_cat([f.cluster_offsets for f in frames], np.int32)
    idx = _pad_cat([f.cluster_indices for f in frames], np.int32)
    obs = _pad_cat([f.obstacles for f in frames], np.float32)
    if world == 1:
        return _unpack(c, l, offs, idx, obs)
    per = -(-total_frames // world)
    head = torch.zeros(3 + 2 * per, dtype=torch.int64, device=device)
    head[0], head[1], head[2] = nloc, len(offs), len(idx)
    head[3:3 + nloc] = torch.from_numpy(c).to(device)
    head[3 + per:3 + per + nloc] = torch.from_numpy(l).to(device)
    heads = [torch.empty_like(head) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    heads = [h.cpu().numpy() for h in heads]
    max_offs = max(int(h[1]) for h in heads)
    max_idx = max(int(h[2]) for h in heads)
    max_obs = max(int(h[3:3 + per][:int(h[0])].sum()) for h in heads) * 4

    def padded(a, n, dtype):
        t = torch.zeros(max(n, 1), dtype=dtype, device=device)
        if len(a):
            t[:len(a)] = torch.from_numpy(a).to(device)
        return t

    payloads = []
    for a, n, dt in ((offs, max_offs, torch.int32), (idx, max_idx, torch.int32), (obs, max_obs, torch.float32)):
        t = padded(a, n, dt)
        out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, out, dst=dst, group=group)
        payloads.append(out)
    if rank != dst:
        return None
    cs, ls, os_, is_, bs = [], [], [], [], []
    for r, h in enumerate(heads):
        n_r = int(h[0])
        c_r = h[3:3 + n_r]
        l_r = h[3 + per:3 + per + n_r]
        cs.append(c_r)
        ls.append(l_r)
        os_.append(payloads[0][r].cpu().numpy()[:int(h[1])])
        is_.append(payloads[1][r].cpu().numpy()[:int(h[2])])
        bs.append(payloads[2][r].cpu().numpy()[:int(c_r.sum()) * 4])
    return _unpack(np.concatenate(cs), np.concatenate(ls), np.concatenate(os_), np.concatenate(is_),
                   np.concatenate(bs))


def _unpack(c, l, offs, idx, obs):
    co, ci, ob = [], [], []
    po = pi = pb = 0
    for k in range(len(c)):
        co.append(offs[po:po + c[k] + 1].copy())
        po += c[k] + 1
        ci.append(idx[pi:pi + l[k]].copy())
        pi += l[k]
        ob.append(obs[pb:pb + 4 * c[k]].reshape(-1, 4).copy())
        pb += 4 * c[k]
    return GatheredResults(np.asarray(c), np.asarray(l), co, ci, ob)
