"""Fixed cost and bandwidth of the host<->device copies fiss_plan_grid_host issues (pinned buffers), on this box."""
import time
import torch

dev = torch.device("cuda", 0)
s = torch.cuda.current_stream()
def timed(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
def dev_timed(fn, n=200):
    for _ in range(10): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for nbytes in (10 * 1024, 24 * 1024, 410880, 821760, 1643520, 3287040, 6574080):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d2h = lambda: h.copy_(d, non_blocking=True)
    h2d = lambda: d.copy_(h, non_blocking=True)
    def d2h_sync():
        h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print(f"{nbytes:9d} B  D2H back-to-back {dev_timed(d2h):7.1f} us  ({nbytes / dev_timed(d2h) / 1e3:6.1f} GB/s)   H2D {dev_timed(h2d):7.1f} us   D2H + sync wall {timed(d2h_sync):7.1f} us")
