"""Host-path probe: D2H / H2D copy rate of this process's GPU while the other GPUs of the box do the same.
python tools/pcie_probe.py <start_epoch_s>   (all processes start their timed loops at the same wall-clock second)"""
import sys, time
import torch

t_start = float(sys.argv[1])
MB = 1 << 20
n = 128 * MB
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()

def run(kind, piece, secs=2.0):
    reps = 0
    with torch.cuda.stream(s):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < secs:
            for o in range(0, n, piece):
                if kind == "d2h":
                    host[o:o + piece].copy_(dev[o:o + piece], non_blocking=True)
                else:
                    dev[o:o + piece].copy_(host[o:o + piece], non_blocking=True)
            s.synchronize()
            reps += 1
        dt = time.perf_counter() - t0
    return reps * n / dt / 1e9

out = []
for kind, piece in (("d2h", n), ("d2h", MB), ("d2h", 4 * MB), ("h2d", n)):
    while time.time() < t_start:
        time.sleep(0.001)
    out.append("%s piece %4d MB: %6.1f GB/s" % (kind, piece // MB, run(kind, piece)))
    t_start += 4.0
print(" | ".join(out))
