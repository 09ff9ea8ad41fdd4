"""One warm-up + one extraction of B frames (default 64) of config 1: the process ncu attaches to."""
import sys, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from orb_slam_fusion_b200 import _abi as A
W, H = 752, 480
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = P.synth_frames("blocks", B, W, H, seed=1)
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7, max_batch=B)
cap = ex.max_keypoints() + 8
kps = torch.empty((B, cap, 7), dtype=torch.float32, device='cuda'); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device='cuda')
n = torch.empty(B, dtype=torch.int32, device='cuda'); nm = torch.empty(B, dtype=torch.int32, device='cuda')
st = A.torch_stream(frames.device)
for _ in range(2):
    ex.extract_batch_into(frames.data_ptr(), B, W, H, frames.stride(1), frames.stride(0), A.MEM_DEVICE, (0, 0), kps.data_ptr(), desc.data_ptr(), cap, n.data_ptr(), nm.data_ptr(), st)
torch.cuda.synchronize()
print("keypoints/frame", float(n.float().mean()))
