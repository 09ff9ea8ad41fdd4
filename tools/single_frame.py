"""One blocking single-frame extraction (config 1) repeated a few times: target for `ncu -k regex:k_octree`."""
import sys
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
import torch
frames = P.synth_frames("blocks", 1, 752, 480, seed=1)
img = frames[0].cpu().numpy()
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
for _ in range(6):
    ex(img)
print("ok")
