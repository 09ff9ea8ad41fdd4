"""The tracker's whole-function projection calls of bench.py (one window per keypoint of one frame), for an ncu launch
list (per-kernel durations) and a p50:  python tools/tracker_phases.py [reps]"""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
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
fa = lambda: m.SearchByProjection(k0, d0, geom, qw, d0, None, None, None, None, 100, 0.8)
fb = lambda: m.SearchByProjectionLast(k0, d0, geom, qw, d0, k0["angle"], None, None, None, None, 100, True)
for name, fn in (("SearchByProjection", fa), ("SearchByProjectionLast", fb)):
    for _ in range(min(20, reps)):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    print("%s: p50 %.1f us  min %.1f  matches %d of %d windows" % (name, 1e6 * np.median(ts), 1e6 * min(ts), r[0], n0))
