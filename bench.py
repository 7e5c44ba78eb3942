#!/usr/bin/env python
"""Benchmark of the per-frame perception hot path (BASELINE.json: points/s & frames/s per B200).

A "step" = one pass of the whole hot path (crop -> VoxelGrid -> RANSAC plane loop -> Euclidean clustering ->
centroid/radius) over one batch of synthetic HDL-64-style 120 000-point frames (BASELINE.json configs[1],
parameters of SURVEY.md 8d config 2), `--batch` frames per GPU per step.

  value     points/s, whole job, frames already resident in HBM (the batch, 1.97 GB at 1024 frames, is larger
            than the 126 MB L2, so every step re-reads its input from HBM; no explicit L2 flush)
  e2e       the same through the C ABI with HOST (pinned) frames: H2D of the frames and D2H of the results
            inside the timed region
  roofline  the dominant kernel, timed live with CUDA-event pairs on the library's stream during the timed steps
  cpu_baseline  the CPU oracle (PCL-semantics restatement, oracle/) on a bounded sample of the same frames

`--impl reference` times that CPU restatement with all host threads (the reference itself needs ROS + PCL and
cannot be built in this image).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = 2  # BASELINE.json configs[1]: HDL-64-style 120k-point frame
METRIC = "points/sec per B200 job (HDL-64 120k-point frames: crop + 0.1 m voxel + RANSAC ground removal + Euclidean clustering + centroid/radius)"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [x.strip() for x in out.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = []
        reasons = set()
        mx = None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_frames_per_sec(params, frames, threads):
    """CPU oracle over `frames` ([F, n, 4]) with `threads` host threads; returns (frames/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    O.lib()
    t0 = time.perf_counter()
    if threads <= 1:
        for f in range(len(frames)):
            O.process(params, frames[f])
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(lambda f: O.process(params, frames[f]), range(len(frames))))
    dt = time.perf_counter() - t0
    return len(frames) / dt, dt


def run_reference(args, rank, world, emit):
    """--impl reference: the CPU restatement (oracle port) with all host threads, same metric/config."""
    if rank != 0:
        return
    from pointcloud_obstacle_processing_b200 import synth
    n = synth.points_per_frame(CONFIG)
    threads = os.cpu_count() or 1
    sample = max(threads, min(args.batch, 8 * threads))
    frames = synth.frames(CONFIG, 0, sample)
    params = synth.params(CONFIG)
    for _ in range(args.warmup):
        oracle_frames_per_sec(params, frames[:threads], threads)
    times = []
    for _ in range(args.steps):
        _, dt = oracle_frames_per_sec(params, frames, threads)
        times.append(dt)
    total = sum(times)
    fps = args.steps * sample / total
    pps = fps * n
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: HDL-64-style 120k-point frame, crop + 0.1 m voxel + ground-plane "
                               "RANSAC + Euclidean clustering", "points_per_frame": n, "frames_per_step": sample},
        "frames_per_sec": fps,
        "cpu_baseline": {"value": pps, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} synthetic config-2 frames per step, frame-parallel over {threads} host "
                                   f"threads; PCL-semantics CPU restatement (oracle/), not PCL itself"},
        "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    # the contract is ONE JSON line on stdout: everything else that libraries print there (NCCL's version banner,
    # ...) is sent to stderr; emit() writes to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--latency-reps", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=128, help="frames of the CPU-baseline sample")
    ap.add_argument("--no-kernel-timing", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = synth.points_per_frame(CONFIG)
    B = args.batch
    params = synth.params(CONFIG)
    # distinct frames per rank (weak scaling: per-GPU work fixed)
    host = torch.empty((B, n, 4), dtype=torch.float32).pin_memory()
    synth.frames(CONFIG, rank * B, B, out=host.numpy())
    dev = host.to(f"cuda:{local_rank}", non_blocking=False)
    counts = np.full(B, n, np.int32)
    from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi
    # `value`: frames resident in HBM and results LEFT in HBM (outputs | OUT_DEVICE: the result pointers are device
    # pointers for a GPU-side consumer -- here the NCCL result gather reads them in place); counts, warnings and the
    # plane record still come back to the host every step.  `value_results_to_host` and `e2e` copy every result
    # array to pinned host memory.
    params_dev = params.copy()
    params_dev.outputs = abi.OUT_DEFAULT | abi.OUT_DEVICE
    op = ObstacleProcessor(params_dev, n, max_batch=B, device=local_rank)

    def step_device():
        return op.process_batch_raw(dev.data_ptr(), n, counts)

    def step_host():
        return op.process_batch_raw(host.data_ptr(), n, counts)

    # ---- final result gather, once per step: per-frame obstacle counts (all-gather), then the obstacle
    # records padded to the largest rank total (gather to rank 0), over NCCL.  The per-frame fields are read as numpy
    # views over the ctypes result array; the records of consecutive frames of a wave are adjacent in the library's
    # pinned result buffer, so they are staged run by run (a handful of memmoves per step).
    # The exchange is asynchronous and double-buffered: the collectives of step k run on NCCL's stream while step k+1
    # computes; gather_flush() waits for the outstanding ones before a timed region ends.
    from pointcloud_obstacle_processing_b200._ctypes_abi import FrameResult
    PADCAP = B * 128  # obstacle records per rank and step (fixed: no size negotiation, no host synchronisation)
    slots = []
    if world > 1:
        for _ in range(2):
            slots.append({
                "obs_stage": torch.empty((PADCAP, 4), dtype=torch.float32).pin_memory(),
                "cnt_stage": torch.empty(B + 1, dtype=torch.int32).pin_memory(),
                "dev_cnt": torch.empty(B + 1, dtype=torch.int32, device=dev.device),
                "all_cnt": torch.empty((world, B + 1), dtype=torch.int32, device=dev.device),
                "pad": torch.zeros((PADCAP, 4), dtype=torch.float32, device=dev.device),
                "out": torch.empty((world, PADCAP, 4), dtype=torch.float32, device=dev.device) if rank == 0 else None,
                "work": []})
    gather_step = [0]
    results_on_device = [True]
    off_c, off_p, rec = FrameResult.n_clusters.offset, FrameResult.obstacles.offset, C.sizeof(FrameResult)

    def gather_flush():
        for sl in slots:
            for w in sl["work"]:
                w.wait()
            sl["work"] = []

    def gather_results(res):
        if world == 1 or os.environ.get("PCOP_BENCH_NO_GATHER"):
            return
        sl = slots[gather_step[0] & 1]
        gather_step[0] += 1
        for w in sl["work"]:  # the slot's buffers are free again once its previous exchange has completed
            w.wait()
        raw = np.frombuffer(res, dtype=np.uint8).reshape(len(res), rec)
        ns = raw[:, off_c:off_c + 4].copy().view(np.int32).ravel()
        ptrs = raw[:, off_p:off_p + 8].copy().view(np.uint64).ravel()
        tot = int(ns.sum())
        assert tot <= PADCAP, "more obstacle records than the exchange buffer holds"
        sl["cnt_stage"][:B].copy_(torch.from_numpy(ns))
        sl["cnt_stage"][B] = tot
        sl["dev_cnt"].copy_(sl["cnt_stage"], non_blocking=True)
        w1 = dist.all_gather_into_tensor(sl["all_cnt"].view(-1), sl["dev_cnt"], async_op=True)
        if tot:
            live = np.flatnonzero(ns > 0)
            ends = ptrs[live] + 16 * ns[live].astype(np.uint64)
            brk = np.flatnonzero(ptrs[live][1:] != ends[:-1]) + 1  # a new run starts where the records are not adjacent
            starts = np.concatenate([[0], brk])
            stops = np.concatenate([brk, [len(live)]])
            o = 0
            for a, b_ in zip(starts, stops):
                nrec = int(ns[live[a:b_]].sum())
                if results_on_device[0]:  # device -> device, straight out of the library's result buffer
                    op.copy_device(sl["pad"].data_ptr() + 16 * o, int(ptrs[live[a]]), 16 * nrec)
                else:
                    C.memmove(sl["obs_stage"].data_ptr() + 16 * o, int(ptrs[live[a]]), 16 * nrec)
                o += nrec
            if not results_on_device[0]:
                sl["pad"][:tot].copy_(sl["obs_stage"][:tot], non_blocking=True)
        w2 = dist.gather(sl["pad"], list(sl["out"].unbind(0)) if rank == 0 else None, dst=0, async_op=True)
        sl["work"] = [w1, w2]

    # ---- device-resident run ---------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        gather_results(step_device())
    stage_acc = {}
    launches = 0
    alg_bytes = 0.0
    dev_us = 0.0
    with ClockSampler(local_rank) as clocks:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step_device()
            gather_results(res)
            dev_us += op.last_elapsed_us
            launches += op.last_launch_count
            alg_bytes += op.last_algorithmic_bytes
        gather_flush()
        barrier()
        wall = time.perf_counter() - t0
    # ---- the same K steps with every result array copied to pinned host memory -----------------------------
    op.set_params(params)
    results_on_device[0] = False
    for _ in range(2):
        gather_results(step_device())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gather_results(step_device())
    gather_flush()
    barrier()
    wall_host_results = time.perf_counter() - t0
    # (the passes below keep the host-result mode: the single lane of the instrumented pass runs four waves)
    # ---- instrumented pass: the same K steps with a CUDA-event pair around every launch on the library's streams.
    # Per-kernel timing needs each kernel alone on the GPU, so the library serialises its lanes here; this pass
    # feeds `roofline`, `kernels` and `stage_ms_per_step` only, never `value`.
    kernel_times = {}
    sort_keys = 0
    wall_instr = None
    if not args.no_kernel_timing:
        op.enable_kernel_timing(True)
        step_device()  # untimed: the single lane of this pass grows its pinned result buffer once
        op.enable_kernel_timing(True)  # (re-enabling clears the accumulated totals)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_device()
            sort_keys += op.last_sort_pass_keys
            for k, v in op.stage_times_us().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
        barrier()
        wall_instr = time.perf_counter() - t0
        kernel_times = op.kernel_times()
        op.enable_kernel_timing(False)
    counts_sum = {k: sum(getattr(r, k) for r in res) for k in
                  ("n_input", "n_crop", "n_voxel", "n_remaining", "n_clusters", "n_cluster_points")}
    d2h_bytes = sum(r.n_remaining * 20 + (r.n_clusters + 1) * 4 + r.n_cluster_points * 4 + r.n_clusters * 16
                    for r in res) + B * 600

    # ---- end-to-end run (host frames in, results out) ------------------------------------------------
    op.set_params(params)
    results_on_device[0] = False
    for _ in range(2):
        step_host()
    barrier()
    t1 = time.perf_counter()
    d2h_exact = 0.0
    for _ in range(args.steps):
        gather_results(step_host())
        d2h_exact += op.last_d2h_bytes
    gather_flush()
    barrier()
    wall_e2e = time.perf_counter() - t1
    d2h_bytes = d2h_exact / args.steps  # counted by the library from the copies it issued

    # ---- max over ranks -----------------------------------------------------------------------------
    t = torch.tensor([wall, wall_e2e, dev_us * 1e-6, wall_host_results], dtype=torch.float64, device=dev.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, wall_e2e, dev_s, wall_host_results = [float(x) for x in t.tolist()]

    # ---- single-frame latency (rank 0) -----------------------------------------------------------------
    lat = None
    if rank == 0 and args.latency_reps > 0:
        op1 = ObstacleProcessor(params, n, max_batch=1, device=local_rank)
        one = host[0].numpy()
        for _ in range(10):
            op1.process_batch_raw(host.data_ptr(), n, counts[:1])
        ts = []
        for i in range(args.latency_reps):
            a = time.perf_counter()
            op1.process_batch_raw(host.data_ptr() + (i % B) * n * 16, n, counts[:1])
            ts.append((time.perf_counter() - a) * 1e3)
        ts.sort()
        lat = {"p50_ms": ts[len(ts) // 2], "p90_ms": ts[int(len(ts) * 0.9)], "min_ms": ts[0], "reps": len(ts),
               "what": "one 120k-point frame, host pinned in -> results on host, wall clock"}
        op1.close()

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        sample = min(args.cpu_sample, B)
        fps1, dt1 = oracle_frames_per_sec(params, host.numpy()[:sample], 1)
        threads = os.cpu_count() or 1
        fpsn, dtn = oracle_frames_per_sec(params, host.numpy()[:min(B, max(sample, 2 * threads))], threads)
        cpu = {"value": fps1 * n, "unit": "points/s", "cores": 1, "kind": "port",
               "sample": f"first {sample} frames of the step's batch, single thread ({dt1:.1f} s); PCL-semantics CPU "
                         f"restatement (oracle/), the reference node is single-threaded (od.cpp:1014)",
               "frames_per_sec": fps1,
               "all_cores": {"value": fpsn * n, "unit": "points/s", "cores": threads, "frames_per_sec": fpsn}}

    if rank == 0:
        total_frames = world * args.steps * B
        total_points = total_frames * n
        peak, peak_kind = measured_peak_gbs()
        value = total_points / wall
        # dominant kernel, timed live
        roof = None
        ktable = {}
        if kernel_times:
            tot_us = sum(v[0] for v in kernel_times.values())
            for k, (us, cnt) in sorted(kernel_times.items(), key=lambda kv: -kv[1][0]):
                ktable[k] = {"total_us": round(us, 1), "launches": cnt, "share": round(us / tot_us, 4)}
            top = max(kernel_times.items(), key=lambda kv: kv[1][0])
            name, (us, cnt) = top
            per_frame = {k: v / B for k, v in counts_sum.items()}
            N_, M_, V_, P_ = (counts_sum[k] * args.steps for k in ("n_input", "n_crop", "n_voxel", "n_remaining"))
            L_, C_ = (counts_sum[k] * args.steps for k in ("n_cluster_points", "n_clusters"))
            alg_all = {  # algorithmic bytes of ALL launches of a kernel over the instrumented steps (DESIGN.md "kernels")
                "k_vf_sort_pass": 16.0 * sort_keys,              # 8-byte (key, index) element read + written once per pass
                "k_vf_crop_key": 16.0 * N_ + 8.0 * M_,           # input read once, (key, index) of the survivors written
                "k_vf_reduce": 8.0 * M_ + 16.0 * M_ + 20.0 * V_,  # sorted pairs + point gather in, voxels + keys out
                "k_sort_pass": 16.0 * sort_keys,
                "k_crop": 16.0 * N_ + 20.0 * M_,                 # SURVEY 8d crop row
                "k_voxel_keys": 16.0 * M_ + 4.0 * M_,
                "k_voxel_centroid": 16.0 * M_ + 8.0 * M_ + 20.0 * V_,
                "k_plane_score": 16.0 * V_,
                "k_plane_moments": 16.0 * V_,
                "k_plane_extract": 16.0 * V_ + 20.0 * P_,
                "k_ece_small": 32.0 * P_ + 8.0 * L_ + 20.0 * C_,  # SURVEY 8d ECE + centroid/radius rows (fused kernel)
                "k_pack": 2.0 * (20.0 * P_ + 4.0 * L_ + 20.0 * C_),
            }
            for k in ktable:
                if k in alg_all and kernel_times[k][0] > 0:
                    ktable[k]["achieved_GBps"] = round(alg_all[k] / (kernel_times[k][0] * 1e-6) / 1e9, 1)
                    ktable[k]["frac_of_hbm_peak"] = round(ktable[k]["achieved_GBps"] / peak, 4)
            alg = alg_all.get(name)
            # DRAM bytes per launch from the committed `ncu --set full` capture (profiles/ncu_full_r01f_summary.csv:
            # dram__bytes_read.sum + dram__bytes_write.sum at 256 frames per launch), scaled by the units per launch
            ncu_dram_bytes_per_unit = {
                "k_vf_sort_pass": ((0.2129e9 + 182.2e6) / 26.35e6, sort_keys),  # per key moved (mean of the 4 passes)
                "k_vf_reduce": ((1.0779e9 + 207.2e6) / 26.35e6, M_),            # per sorted pair
                "k_vf_crop_key": ((0.4921e9 + 177.0e6) / 30.72e6, N_),          # per input point
            }
            if alg is not None and us > 0:
                ach = alg / (us * 1e-6) / 1e9
                tr = ncu_dram_bytes_per_unit.get(name)
                roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": (tr[0] * tr[1] / cnt) if tr else None,
                        "traffic_source": "profiles/ncu_full_r01f_summary.csv (ncu --set full, 256 frames per launch), "
                                          "scaled to this run's units per launch" if tr else None,
                        "peak_source": peak_kind,
                        "launches": cnt, "avg_launch_us": us / cnt,
                        "timed": "second pass of the same K steps with a CUDA-event pair around every launch on the "
                                 "library's stream (lanes serialised so that each kernel runs alone)",
                        "algorithmic_bytes_per_launch": alg / cnt,
                        "share_of_kernel_time": us / tot_us}
            else:
                roof = {"bound": "hbm", "kernel": name, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                        "traffic": None, "peak_source": peak_kind}
        pipe_gbs = world * alg_bytes / wall / 1e9 if wall > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: HDL-64-style 120k-point frame, crop + 0.1 m voxel + ground-plane "
                                   "RANSAC + Euclidean clustering + centroid/radius (SURVEY 8d config 2 parameters)",
                       "points_per_frame": n, "frames_per_gpu_per_step": B,
                       "results": "left in HBM (PCOP_OUT_DEVICE; counts and plane records on the host); "
                                  "value_results_to_host and e2e copy them to pinned host memory",
                       "l2": "no flush: the per-step input batch (%.0f MB) exceeds the 126 MB L2" % (B * n * 16 / 1e6),
                       "parallelism": f"frames sharded over {world} GPU(s), no intra-frame collective"},
            "frames_per_sec": total_frames / wall,
            "value_results_to_host": {"value": total_points / wall_host_results, "unit": "points/s",
                                      "ms_per_step": 1000.0 * wall_host_results / args.steps,
                                      "what": "same K steps, frames resident in HBM, every result array copied to pinned "
                                              "host memory inside the step"},
            "device_ms_per_step": 1000.0 * dev_s / args.steps,
            "instrumented_ms_per_step": 1000.0 * wall_instr / args.steps if wall_instr else None,
            "lanes": int(os.environ.get("PCOP_LANES", str(2 if (os.cpu_count() or 1) < 8 * torch.cuda.device_count() else min(4, max(2, B // 256))))),
            "e2e": {"value": total_points / wall_e2e, "unit": "points/s", "h2d_bytes_per_step": B * n * 16 + B * 4,
                    "d2h_bytes_per_step": int(d2h_bytes), "frames_per_sec": total_frames / wall_e2e,
                    "ms_per_step": 1000.0 * wall_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roof,
            "pipeline_roofline": {"algorithmic_bytes_per_frame": alg_bytes / (args.steps * B),
                                  "achieved_GBps": pipe_gbs, "peak": peak * world, "frac": pipe_gbs / (peak * world) if pipe_gbs else None,
                                  "what": "sum of SURVEY 8d stage bytes over all frames of all GPUs / elapsed (whole pipeline); peak = measured HBM peak x GPUs"},
            "stage_ms_per_step": {k: round(v / args.steps / 1000.0, 3) for k, v in stage_acc.items()},
            "kernels": ktable,
            "counts_per_frame": {k: v / B for k, v in counts_sum.items()},
            "latency": lat,
            "cpu_baseline": cpu,
        }
        emit(line)
    op.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
