"""Barrier-aligned multi-GPU probe of the library alone (no result gather): every rank runs the same 1024-frame calls at
the same time; prints per-rank call times, SM clocks / power under load (NVML) and, on rank 0, the wave trace.
torchrun --nproc-per-node N tools/n8_sync_probe.py [device|host] [reps]"""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
mode = sys.argv[1] if len(sys.argv) > 1 else "device"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(lr)
ncpu = os.cpu_count() or 1
per = max(1, ncpu // max(world, 1))
if world > 1 and not os.environ.get("PROBE_NO_PIN"):
    os.sched_setaffinity(0, set(range(lr * per, min(ncpu, (lr + 1) * per))))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank == 0 and os.environ.get("PROBE_TRACE"):
    os.environ["PCOP_TRACE"] = "1"
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth
from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi

B = 1024; n = synth.points_per_frame(2); p = synth.params(2)
if mode == "device":
    p.outputs |= abi.OUT_DEVICE
host = torch.empty((B, n, 4), dtype=torch.float32).pin_memory()
synth.frames(2, rank * B, B, out=host.numpy())
dev = host.cuda()
counts = np.full(B, n, np.int32)
op = ObstacleProcessor(p, n, max_batch=B, device=lr)
for _ in range(3):
    op.process_batch_raw(dev.data_ptr(), n, counts)

rows = []
stop = threading.Event()
def sample():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    while not stop.is_set():
        rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                     pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        stop.wait(0.003)
t = threading.Thread(target=sample, daemon=True); t.start(); time.sleep(0.3); rows.clear()

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
barrier()
ts = []
t0 = time.perf_counter()
for _ in range(reps):
    a = time.perf_counter()
    op.process_batch_raw(dev.data_ptr(), n, counts)
    ts.append((time.perf_counter() - a) * 1e3)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / reps
stop.set(); t.join()
sm = [r[0] for r in rows]; pw = [r[1] for r in rows]; rs = 0
for r in rows: rs |= r[2]
print(f"[{mode} N={world}] rank {rank}: {wall:.3f} ms per call (min {min(ts):.3f} median {np.median(ts):.3f} max {max(ts):.3f}); "
      f"SM MHz median {np.median(sm) if sm else None} min {min(sm) if sm else None}; power W median {np.median(pw) if pw else None:.0f} max {max(pw) if pw else None:.0f}; "
      f"throttle mask {rs:#x}; samples {len(sm)}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
