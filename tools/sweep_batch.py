import time, torch, numpy as np, sys
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from orb_slam_fusion_b200 import _abi as A
W,H=752,480
B=512
frames = P.synth_frames("blocks", B, W, H, seed=1)
h_frames = torch.empty((B,H,W), dtype=torch.uint8, pin_memory=True); h_frames.copy_(frames)
for mb in (32, 64, 128, 256, 512):
    ex = P.OrbExtractor(1000,1.2,8,20,7, max_batch=mb)
    cap = ex.max_keypoints()+8
    kps = torch.empty((B,cap,7), dtype=torch.float32, device='cuda'); desc = torch.empty((B,cap,32), dtype=torch.uint8, device='cuda')
    n = torch.empty(B, dtype=torch.int32, device='cuda'); nm = torch.empty(B, dtype=torch.int32, device='cuda')
    st = A.torch_stream(frames.device)
    def step(): ex.extract_batch_into(frames.data_ptr(), B, W, H, frames.stride(1), frames.stride(0), A.MEM_DEVICE, (0,0), kps.data_ptr(), desc.data_ptr(), cap, n.data_ptr(), nm.data_ptr(), st)
    for _ in range(3): step()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    dev_fps = B*10/(e0.elapsed_time(e1)*1e-3)
    hk = torch.empty((B,cap,7), dtype=torch.float32, pin_memory=True); hd = torch.empty((B,cap,32), dtype=torch.uint8, pin_memory=True)
    hn = torch.empty(B, dtype=torch.int32, pin_memory=True); hm = torch.empty(B, dtype=torch.int32, pin_memory=True)
    def e2e(): ex.extract_batch_into(h_frames.data_ptr(), B, W, H, W, W*H, A.MEM_HOST, (0,0), hk.data_ptr(), hd.data_ptr(), cap, hn.data_ptr(), hm.data_ptr(), None)
    for _ in range(2): e2e()
    t0=time.perf_counter()
    for _ in range(10): e2e()
    dt=time.perf_counter()-t0
    print("max_batch %4d: device %.0f fps, e2e %.0f fps" % (mb, dev_fps, B*10/dt), flush=True)
    del ex
# raw PCIe H2D rate
d = torch.empty_like(frames)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10): d.copy_(h_frames, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t0
print("H2D GB/s", 10*h_frames.numel()/dt/1e9)
