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


_cudart = None


def _cuda_memcpy_h2d_async(dst_dev: int, src_host: int, nbytes: int, stream: int):
    """cudaMemcpyAsync(-> device, cudaMemcpyDefault) on raw addresses: the result arrays are library-owned pinned host
    memory, or device memory when the frames were processed with outputs | OUT_DEVICE (unified addressing tells)"""
    global _cudart
    import ctypes as C
    if _cudart is None:
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart is None:
            import glob
            import os
            hits = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
            _cudart = C.CDLL(hits[0])
        _cudart.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        _cudart.cudaMemcpyAsync.restype = C.c_int
    st = _cudart.cudaMemcpyAsync(C.c_void_p(dst_dev), C.c_void_p(src_host), C.c_size_t(nbytes), 4, C.c_void_p(stream))
    if st != 0:
        raise RuntimeError(f"cudaMemcpyAsync failed with status {st}")


class ResultGather:
    """Per-step result gather for batched runs, straight from the library's result arrays.

    Every step each rank hands in its ctypes result array (`process_batch_raw`); the per-frame `cluster_offsets`,
    `cluster_indices` and `obstacles` arrays -- adjacent in the library's pinned result buffer wave by wave, so they are
    staged run by run with a handful of memmoves -- go to `dst` in two collectives: an all-gather of the per-frame
    (C, L) counts and a gather of one int32 payload per rank ([offsets | indices | obstacle records], padded to a fixed
    capacity so that no rank has to wait for another's sizes).  Asynchronous and double-buffered: the collectives of
    step k run on the backend's stream while step k + 1 computes; `flush()` waits for the outstanding ones,
    `last()` unpacks the most recent completed exchange on `dst` (GatheredResults, frames in rank-major order).
    """

    def __init__(self, frames_per_rank: int, device: torch.device, group=None, dst: int = 0,
                 ints_per_frame: int = 8192):
        import ctypes as C
        from ._ctypes_abi import FrameResult
        self._C = C
        self.B = frames_per_rank
        self.device = device
        self.group = group
        self.dst = dst
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cap = frames_per_rank * ints_per_frame  # int32 elements per rank and step
        self.rec = C.sizeof(FrameResult)
        self.off = {k: getattr(FrameResult, k).offset for k in
                    ("n_clusters", "n_cluster_points", "cluster_offsets", "cluster_indices", "obstacles")}
        pin = device.type == "cuda"
        self.slots = []
        for _ in range(2):
            stage = torch.empty(self.cap, dtype=torch.int32)
            head = torch.empty(2 * frames_per_rank + 4, dtype=torch.int32)
            self.slots.append({
                "stage": stage.pin_memory() if pin else stage, "head": head.pin_memory() if pin else head,
                "dev_head": torch.empty_like(head, device=device), "all_head": torch.empty((self.world, head.numel()), dtype=torch.int32, device=device),
                "pad": torch.zeros(self.cap, dtype=torch.int32, device=device),
                "out": torch.empty((self.world, self.cap), dtype=torch.int32, device=device) if self.rank == dst else None,
                "work": [], "used": False})
        self.step = 0
        self.bytes_per_step = 0

    def _runs(self, ptrs, sizes):
        """(first frame index, bytes) of every maximal run of frames whose arrays are adjacent in memory"""
        live = np.flatnonzero(sizes > 0)
        if len(live) == 0:
            return []
        p, s = ptrs[live], sizes[live].astype(np.uint64)
        brk = np.flatnonzero(p[1:] != p[:-1] + s[:-1]) + 1
        starts = np.concatenate([[0], brk])
        stops = np.concatenate([brk, [len(live)]])
        return [(int(p[a]), int(s[a:b].sum())) for a, b in zip(starts, stops)]

    def submit(self, res):
        C = self._C
        sl = self.slots[self.step & 1]
        self.step += 1
        for w in sl["work"]:  # the slot's buffers are free again once its previous exchange has completed
            w.wait()
        nf = len(res)
        assert nf <= self.B
        raw = np.frombuffer(res, dtype=np.uint8).reshape(nf, self.rec)
        col32 = lambda k: raw[:, self.off[k]:self.off[k] + 4].copy().view(np.int32).ravel()
        col64 = lambda k: raw[:, self.off[k]:self.off[k] + 8].copy().view(np.uint64).ravel()
        c, l = col32("n_clusters"), col32("n_cluster_points")
        parts = (("cluster_offsets", (c.astype(np.int64) + 1) * 4), ("cluster_indices", l.astype(np.int64) * 4),
                 ("obstacles", c.astype(np.int64) * 16))
        totals = [int(sz.sum()) // 4 for _, sz in parts]
        if sum(totals) > self.cap:
            raise ValueError(f"ResultGather: {sum(totals)} int32 per step exceed the exchange capacity {self.cap} "
                             f"(raise ints_per_frame)")
        # CUDA: the arrays go straight from the library's pinned result buffer into the device exchange buffer
        # (asynchronous copies on the current stream; the library keeps a call's results valid while the next call
        # runs).  CPU (gloo tests): staged with memmove.
        on_gpu = self.device.type == "cuda"
        base = sl["pad"].data_ptr() if on_gpu else sl["stage"].data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream if on_gpu else None
        o = 0
        for key, sz in parts:
            if (col64(key)[sz > 0] == 0).any():
                raise ValueError("ResultGather needs frames processed with outputs | OUT_CLUSTERS | OUT_OBSTACLES")
            for ptr, nbytes in self._runs(col64(key), sz):
                if on_gpu:
                    _cuda_memcpy_h2d_async(base + o, ptr, nbytes, stream)
                else:
                    C.memmove(base + o, ptr, nbytes)
                o += nbytes
        used = o // 4
        if on_gpu:  # wait_staged(): the result arrays have been read
            self._staged = torch.cuda.Event()
            self._staged.record(torch.cuda.current_stream(self.device))
        h = sl["head"]
        h[0], h[1], h[2], h[3] = nf, totals[0], totals[1], totals[2]
        h[4:4 + nf] = torch.from_numpy(c)
        h[4 + self.B:4 + self.B + nf] = torch.from_numpy(l)
        sl["dev_head"].copy_(h, non_blocking=True)
        if not on_gpu:
            sl["pad"][:used].copy_(sl["stage"][:used], non_blocking=True)
        self.bytes_per_step = 4 * (used + h.numel())
        if self.world == 1:
            sl["all_head"][0].copy_(sl["dev_head"])
            sl["out"][0].copy_(sl["pad"])
            sl["work"] = []
        else:
            w1 = dist.all_gather_into_tensor(sl["all_head"].view(-1), sl["dev_head"], group=self.group, async_op=True)
            w2 = dist.gather(sl["pad"], list(sl["out"].unbind(0)) if self.rank == self.dst else None, dst=self.dst,
                             group=self.group, async_op=True)
            sl["work"] = [w1, w2]
        sl["used"] = True

    def wait_staged(self):
        """blocks until the last submit()'s copies out of the library's result buffer have run (call it before the
        call after next overwrites that buffer)"""
        ev = getattr(self, "_staged", None)
        if ev is not None:
            ev.synchronize()

    def flush(self):
        for sl in self.slots:
            for w in sl["work"]:
                w.wait()
            sl["work"] = []

    def last(self):
        """GatheredResults of the most recent exchange (on `dst`; None elsewhere)"""
        self.flush()
        if self.rank != self.dst or self.step == 0:
            return None
        sl = self.slots[(self.step - 1) & 1]
        heads = sl["all_head"].cpu().numpy()
        outs = sl["out"].cpu().numpy()
        cs, ls, os_, is_, bs = [], [], [], [], []
        for r in range(self.world):
            h = heads[r]
            nf, to, ti, tb = int(h[0]), int(h[1]), int(h[2]), int(h[3])
            cs.append(h[4:4 + nf].astype(np.int64))
            ls.append(h[4 + self.B:4 + self.B + nf].astype(np.int64))
            os_.append(outs[r][:to])
            is_.append(outs[r][to:to + ti])
            bs.append(outs[r][to + ti:to + ti + tb].view(np.float32))
        return _unpack(np.concatenate(cs), np.concatenate(ls), np.concatenate(os_), np.concatenate(is_), np.concatenate(bs))
