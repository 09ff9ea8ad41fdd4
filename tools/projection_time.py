"""Time of one whole SearchByProjection call (host buffers through the C ABI) vs the CPU paths, 1000 keypoints x 900 map points."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import orb_slam_fusion_b200 as P
from oracle import oracle as O, ref as R
from test_oracle_vs_ref_frame import _frame_and_points, projection_windows, W, H
kps, desc, pts, qdesc, src, rng = _frame_and_points(O, 1, nq=900)
geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
sf = O.Extractor(1000).tables()["scale"]
keep, q = projection_windows(O, sf, pts, 3.0, False, 40.0)
m = P.ORBmatcher(0.8, True)
for _ in range(20): nm, got = m.SearchByProjection(kps, desc, geom, q, qdesc[keep], None)
t = []
for _ in range(200):
    t0 = time.perf_counter(); nm, got = m.SearchByProjection(kps, desc, geom, q, qdesc[keep], None); t.append(time.perf_counter() - t0)
print("GPU whole function, host buffers: p50 %.1f us (%d windows, %d keypoints, %d matches)" % (1e6 * np.median(t), len(q), len(kps), nm))
t = []
for _ in range(200):
    t0 = time.perf_counter(); r = m.window_search(kps, desc, geom, q, qdesc[keep], None); t.append(time.perf_counter() - t0)
print("GPU window search only: p50 %.1f us" % (1e6 * np.median(t)))
if R.frame_available():
    t = []
    for _ in range(50):
        t0 = time.perf_counter(); R.search_by_projection(kps, desc, (0.0, float(W), 0.0, float(H)), sf, pts, qdesc, None, None, 3.0, 0.8, False, 40.0); t.append(time.perf_counter() - t0)
    print("reference lines on one core (incl. building the Frame grid and the MapPoint objects): p50 %.1f us" % (1e6 * np.median(t)))
