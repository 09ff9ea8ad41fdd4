"""Single-frame latency: p50 of the blocking call (CUDA-graph path) and the per-stage device times
of the same frame through plain launches (profiling mode)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from oracle import oracle as O
img = O.blocks_v1(752, 480, 1, 0)
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
for _ in range(30): ex(img)
lat = []
for _ in range(300):
    t0 = time.perf_counter(); ex(img); lat.append(time.perf_counter() - t0)
print("p50 %.4f ms  p10 %.4f  p90 %.4f" % (1e3 * np.median(lat), 1e3 * np.percentile(lat, 10), 1e3 * np.percentile(lat, 90)))
ex.set_profiling(True); ex.stage_times()
for _ in range(50): ex(img)
ms, ch = ex.stage_times()
print({k: round(1e3 * v / ch, 1) for k, v in ms.items()}, "us per stage; sum %.1f us" % (1e3 * sum(ms.values()) / ch))
