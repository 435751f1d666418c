"""D2H / H2D bandwidth of the box with pinned memory (development aid for the sink's ceiling)."""
import torch, time
n = 1 << 30
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
d2h = timed(lambda: host.copy_(dev, non_blocking=True))
h2d = timed(lambda: dev.copy_(host, non_blocking=True))
print(f"D2H 1 GiB: {d2h:.2f} ms = {n/d2h/1e6:.1f} GB/s;  H2D: {h2d:.2f} ms = {n/h2d/1e6:.1f} GB/s")
# 24.9 MB frames, back to back on one stream
f = 3840*2160*3
d2 = timed(lambda: [host[k*f:(k+1)*f].copy_(dev[k*f:(k+1)*f], non_blocking=True) for k in range(40)])
print(f"40 x 24.9 MB frames D2H: {d2:.2f} ms = {40*f/d2/1e6:.1f} GB/s")
