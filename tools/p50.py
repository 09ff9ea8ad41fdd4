"""p50 of the blocking single-frame call (config 1), like bench.py's p50 leg."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from oracle import oracle as O
img = O.blocks_v1(752, 480, 1, 0)
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
ref = O.Extractor(1000, 1.2, 8, 20, 7, trig=O.TRIG_CR)(img)
for _ in range(30): out = ex(img)
assert out[0] == ref[0] and out[1].tobytes() == ref[1].tobytes() and np.array_equal(out[2], ref[2])
lat = []
for _ in range(500):
    t0 = time.perf_counter(); out = ex(img); lat.append(time.perf_counter() - t0)
assert out[1].tobytes() == ref[1].tobytes() and np.array_equal(out[2], ref[2])
print("p50 %.4f ms  p10 %.4f  p90 %.4f" % (1e3 * np.median(lat), 1e3 * np.percentile(lat, 10), 1e3 * np.percentile(lat, 90)))
k1, d1 = np.empty(ex.max_keypoints(), P.KP_DTYPE), np.empty((ex.max_keypoints(), 32), np.uint8)
lat = []
for _ in range(1000):
    t0 = time.perf_counter(); nm, n = ex.extract_into(img, k1, d1); lat.append(time.perf_counter() - t0)
assert k1[:n].tobytes() == ref[1].tobytes() and np.array_equal(d1[:n], ref[2])
print("bare C-ABI call: p50 %.4f ms  p10 %.4f  p90 %.4f" % (1e3 * np.median(lat), 1e3 * np.percentile(lat, 10), 1e3 * np.percentile(lat, 90)))
