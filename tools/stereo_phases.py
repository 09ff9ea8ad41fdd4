"""Config 2 (stereo pair 2 x 752x480 x 1200 + ComputeStereoMatches) as bench.py runs it: p50 of the stereo frame and of
its two halves; under `ncu --metrics gpu__time_duration.sum` the per-kernel durations.  python tools/stereo_phases.py [reps]"""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
W, H = 752, 480
m = P.ORBmatcher()
exl, exr = P.OrbExtractor(1200, 1.2, 8, 20, 7, max_batch=1), P.OrbExtractor(1200, 1.2, 8, 20, 7, max_batch=1)
hl = P.synth_frames("blocks", 1, W, H, seed=1, first_frame=3)[0].cpu().numpy()
hr = P.synth_frames("blocks", 1, W, H, seed=1, first_frame=3, shift_x=12, noise_seed=2)[0].cpu().numpy()
bf, mb = np.float32(47.90639384423901), np.float32(0.11)
cap = 1400
kl, dl, kr, dr = np.empty(cap, P.KP_DTYPE), np.empty((cap, 32), np.uint8), np.empty(cap, P.KP_DTYPE), np.empty((cap, 32), np.uint8)
ur, dp = np.empty(cap, np.float32), np.empty(cap, np.float32)
ta, tb = [], []
for it in range(reps + min(20, reps)):
    t0 = time.perf_counter()
    exl.extract_begin(hl); exr.extract_begin(hr)
    exl.extract_end(kl, dl); exr.extract_end(kr, dr)
    t1 = time.perf_counter()
    m.stereo_matches_last(exl, exr, bf, mb, ur, dp)
    t2 = time.perf_counter()
    if it >= min(20, reps):
        ta.append(t1 - t0); tb.append(t2 - t1)
print("two extractions in flight: p50 %.1f us; ComputeStereoMatches: p50 %.1f us; stereo frame %.1f us" %
      (1e6 * np.median(ta), 1e6 * np.median(tb), 1e6 * np.median(np.array(ta) + np.array(tb))))
