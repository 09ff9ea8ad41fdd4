"""p50 of the whole-function projection matchers in the dense case of bench.py's tracker leg (one window per keypoint)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from oracle import oracle as O
W, H = 752, 480
img = O.blocks_v1(W, H, 1, 0)
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
_, k0, d0 = ex(img)
n0 = len(k0)
rng = np.random.default_rng(0)
sfac = np.float32(1.2) ** np.arange(8, dtype=np.float32)
qw = np.zeros(n0, P.WQ_DTYPE)
qw["u"] = k0["x"] + rng.normal(0, 3, n0).astype(np.float32)
qw["v"] = k0["y"] + rng.normal(0, 3, n0).astype(np.float32)
qw["r"] = (np.float32(12.0) * sfac[np.clip(k0["octave"], 0, 7)]).astype(np.float32)
qw["min_level"], qw["max_level"] = k0["octave"] - 1, k0["octave"]
geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
m = P.ORBmatcher(0.8, True)
def p50(fn, reps=200):
    for _ in range(20): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))
a = m.SearchByProjection(k0, d0, geom, qw, d0, None, None, None, None, 100, 0.8)
w = O.search_by_projection(k0, d0, geom, qw, d0, None, None, None, None, 100, 0.8)
assert a[0] == w[0] and np.array_equal(a[1], w[1])
print("SearchByProjection(Frame, MapPoints): %.0f us (%d matches)" % (p50(lambda: m.SearchByProjection(k0, d0, geom, qw, d0, None, None, None, None, 100, 0.8)), a[0]))
print("SearchByProjection(Current, Last):    %.0f us" % p50(lambda: m.SearchByProjectionLast(k0, d0, geom, qw, d0, k0["angle"], None, None, None, None, 100, True)))
print("window search only:                   %.0f us" % p50(lambda: m.window_search(k0, d0, geom, qw, d0)))
