"""p50 of the blocking single-frame C-ABI call with a PINNED host frame (what bench.py's p50 leg times)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
fr = P.synth_frames("blocks", 1, 752, 480, seed=1)
h = torch.empty((480, 752), dtype=torch.uint8, pin_memory=True); h.copy_(fr[0]); torch.cuda.synchronize()
img = h.numpy()
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
k1, d1 = np.empty(ex.max_keypoints() + 64, P.KP_DTYPE), np.empty((ex.max_keypoints() + 64, 32), np.uint8)
for _ in range(50): ex.extract_into(img, k1, d1)
lat = []
for _ in range(2000):
    t0 = time.perf_counter(); nm, n = ex.extract_into(img, k1, d1); lat.append(time.perf_counter() - t0)
lat = 1e3 * np.array(lat)
print("pinned frame: p50 %.4f ms  p10 %.4f  p90 %.4f  min %.4f  (n = %d keypoints)" % (np.median(lat), np.percentile(lat, 10), np.percentile(lat, 90), lat.min(), n))
