import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
fr = P.synth_frames("blocks", 1, 752, 480, seed=1)
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
_, k0, d0 = ex(fr[0].cpu().numpy())
n0 = len(k0)
rng = np.random.default_rng(0)
sfac = np.float32(1.2) ** np.arange(8, dtype=np.float32)
qw = np.zeros(n0, P.WQ_DTYPE)
qw["u"] = k0["x"] + rng.normal(0, 3, n0).astype(np.float32)
qw["v"] = k0["y"] + rng.normal(0, 3, n0).astype(np.float32)
qw["r"] = (np.float32(12.0) * sfac[np.clip(k0["octave"], 0, 7)]).astype(np.float32)
qw["min_level"], qw["max_level"] = k0["octave"] - 1, k0["octave"]
geom = (0.0, 0.0, np.float32(64) / np.float32(752), np.float32(48) / np.float32(480), 64, 48)
m = P.ORBmatcher()
f = lambda: m.window_search(k0, d0, geom, qw, d0)
for _ in range(20): f()
ts = []
for _ in range(300):
    t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
print("window_search %d queries: p50 %.1f us" % (n0, 1e6 * np.median(ts)))
