"""e2e (pinned host frames in, keypoints + descriptors out, HOST_ASYNC pipelined calls) for several chunk sizes."""
import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from orb_slam_fusion_b200 import _abi as A
W, H, B = 752, 480, 512
frames = P.synth_frames("blocks", B, W, H, seed=1)
h_frames = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
for mb in (int(a) for a in (sys.argv[1:] or (32, 64, 128, 256))):
    ex = P.OrbExtractor(1000, 1.2, 8, 20, 7, max_batch=mb)
    cap = ex.max_keypoints() + 8
    hk = torch.empty((B, cap, 7), dtype=torch.float32, pin_memory=True); hd = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
    hn = torch.empty(B, dtype=torch.int32, pin_memory=True); hm = torch.empty(B, dtype=torch.int32, pin_memory=True)
    def step(): ex.extract_batch_into(h_frames.data_ptr(), B, W, H, W, W * H, A.MEM_HOST_ASYNC, (0, 0), hk.data_ptr(), hd.data_ptr(), cap, hn.data_ptr(), hm.data_ptr(), None)
    for _ in range(3): step()
    ex.sync()
    t0 = time.perf_counter()
    for _ in range(20): step()
    t_enq = time.perf_counter() - t0
    ex.sync()
    dt = time.perf_counter() - t0
    print("chunk %4d: e2e %.0f frames/s (%.3f ms per 512 frames; host enqueue %.3f ms per step); H2D at that rate %.1f GB/s" %
          (mb, B * 20 / dt, 1e3 * dt / 20, 1e3 * t_enq / 20, B * W * H * 20 / dt / 1e9), flush=True)
    del ex
