"""Does splitting the 512-frame batch over two streams (two handles, 256 frames each) beat one stream?  Kernels of different
stages would have to share SMs for that (idle issue slots of the quadtree / descriptor kernels against the FAST kernel)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from orb_slam_fusion_b200 import _abi as A
W, H, B = 752, 480, 512
frames = P.synth_frames("blocks", B, W, H, seed=1)

def make(nb):
    ex = P.OrbExtractor(1000, 1.2, 8, 20, 7, max_batch=nb)
    cap = ex.max_keypoints() + 8
    kps = torch.empty((nb, cap, 7), dtype=torch.float32, device='cuda'); desc = torch.empty((nb, cap, 32), dtype=torch.uint8, device='cuda')
    n = torch.empty(nb, dtype=torch.int32, device='cuda'); nm = torch.empty(nb, dtype=torch.int32, device='cuda')
    return ex, cap, kps, desc, n, nm

def run(parts):
    nb = B // parts
    hs = [make(nb) for _ in range(parts)]
    ss = [torch.cuda.Stream() for _ in range(parts)]
    def step():
        for p, (ex, cap, kps, desc, n, nm) in enumerate(hs):
            fr = frames[p * nb:(p + 1) * nb]
            ex.extract_batch_into(fr.data_ptr(), nb, W, H, fr.stride(1), fr.stride(0), A.MEM_DEVICE, (0, 0), kps.data_ptr(), desc.data_ptr(), cap,
                                  n.data_ptr(), nm.data_ptr(), ss[p].cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss: s.wait_event(e0)
    for _ in range(10): step()
    for s in ss: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    print("%d stream(s) x %d frames: %.4f ms per %d frames" % (parts, nb, e0.elapsed_time(e1) / 10, B))

for parts in (1, 2, 4, 1, 2):
    run(parts)
