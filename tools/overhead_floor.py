"""Floor of a blocking single-frame call: pinned H2D of one 752x480 frame + one trivial kernel + D2H of 63 KB + sync (torch)."""
import time, numpy as np, torch
h = torch.empty(480 * 752, dtype=torch.uint8, pin_memory=True); d = torch.empty_like(h, device='cuda')
o = torch.empty(63 * 1024, dtype=torch.uint8, device='cuda'); ho = torch.empty(63 * 1024, dtype=torch.uint8, pin_memory=True)
def step():
    d.copy_(h, non_blocking=True); o.add_(1); ho.copy_(o, non_blocking=True); torch.cuda.synchronize()
for _ in range(50): step()
lat = []
for _ in range(2000):
    t0 = time.perf_counter(); step(); lat.append(time.perf_counter() - t0)
lat = 1e3 * np.array(lat)
print("H2D 361 KB + 1 kernel + D2H 63 KB + sync: p50 %.4f ms  p10 %.4f  min %.4f" % (np.median(lat), np.percentile(lat, 10), lat.min()))
