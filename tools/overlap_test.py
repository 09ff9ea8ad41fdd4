"""Does running two half-batches on two streams (two handles) beat one full batch on one stream?"""
import sys, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
from orb_slam_fusion_b200 import _abi as A
W, H, B = 752, 480, 512
frames = P.synth_frames("blocks", B, W, H, seed=1)
def mk(nb):
    ex = P.OrbExtractor(1000, 1.2, 8, 20, 7, max_batch=nb)
    cap = ex.max_keypoints() + 8
    kps = torch.empty((nb, cap, 7), dtype=torch.float32, device='cuda'); desc = torch.empty((nb, cap, 32), dtype=torch.uint8, device='cuda')
    n = torch.empty(nb, dtype=torch.int32, device='cuda'); nm = torch.empty(nb, dtype=torch.int32, device='cuda')
    return ex, cap, kps, desc, n, nm
def run(ex, cap, kps, desc, n, nm, fr, st):
    nb = fr.shape[0]
    ex.extract_batch_into(fr.data_ptr(), nb, W, H, fr.stride(1), fr.stride(0), A.MEM_DEVICE, (0, 0), kps.data_ptr(), desc.data_ptr(), cap, n.data_ptr(), nm.data_ptr(), st)
def timeit(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
full = mk(B)
st0 = A.torch_stream(frames.device)
print("one stream, 512 frames: %.4f ms" % timeit(lambda: run(*full, frames, st0)))
for parts in (2, 4, 8):
    nb = B // parts
    hs = [mk(nb) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def go():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for i, (h, s) in enumerate(zip(hs, streams)):
            s.wait_event(ev)
            run(*h, frames[i * nb:(i + 1) * nb], s.cuda_stream)
            e = torch.cuda.Event(); e.record(s); cur.wait_event(e)
    print("%d streams x %d frames: %.4f ms" % (parts, nb, timeit(go)))
    # serial small batches on one stream, for comparison
    def go1():
        for i, h in enumerate(hs): run(*h, frames[i * nb:(i + 1) * nb], st0)
    print("1 stream, %d x %d frames: %.4f ms" % (parts, nb, timeit(go1)))
    del hs
