"""A/B helper for the wave engine: python tools/wave_ab.py [frames] [reps]  (env knobs are read by pcop_create).
Prints the mean / min wall-clock ms per call with device-resident frames and results delivered to the host."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointcloud_obstacle_processing_b200 import ObstacleProcessor, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = synth.points_per_frame(2)
p = synth.params(2)
if os.environ.get('WAVE_AB_DEVICE'):
    from pointcloud_obstacle_processing_b200 import _ctypes_abi as abi
    p.outputs |= abi.OUT_DEVICE
host = torch.empty((B, n, 4), dtype=torch.float32).pin_memory()
synth.frames(2, 0, B, out=host.numpy())
dev = host.cuda()
counts = np.full(B, n, np.int32)
op = ObstacleProcessor(p, n, max_batch=min(B, 1024))
for _ in range(3):
    op.process_batch_raw(dev.data_ptr(), n, counts)
ts = []
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    op.process_batch_raw(dev.data_ptr(), n, counts)
    ts.append((time.perf_counter() - t0) * 1e3)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PCOP_") or k.startswith("WAVE_AB"))
print(f"[{tag}] B={B} mean {np.mean(ts):.3f} ms  min {np.min(ts):.3f} ms  median {np.median(ts):.3f} ms")
