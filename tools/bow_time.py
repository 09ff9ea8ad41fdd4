"""The bag-of-words legs of bench.py alone (512 frames of config 1): transform and SearchByBoW(frame f, frame f + 1), ms per batch."""
import sys, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
import bench as BN
B, W, H = 512, 752, 480
frames = P.synth_frames("blocks", B, W, H, seed=1)
ex = P.OrbExtractor(1000, 1.2, 8, 20, 7, max_batch=B)
n, nm, kps, desc = ex.extract_batch(frames)
torch.cuda.synchronize()
vparent, vleaf, vdesc, vweight = BN.synth_vocabulary(10, 6, seed=7)
voc = P.ORBVocabulary(10, 6, vparent, vleaf, vdesc, vweight)
m = P.ORBmatcher()
dev = frames.device
pairs = torch.stack([torch.arange(B, dtype=torch.int32), (torch.arange(B, dtype=torch.int32) + 1) % B], 1).to(dev)
def timed(fn, reps=10):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
t_bow, bow = timed(lambda: voc.transform_batch(desc, n, 4))
t_sb, (sb_n, _) = timed(lambda: m.SearchByBoW(kps, desc, n, bow, pairs, None, 0.7, True))
print("transform %.4f ms per %d frames; SearchByBoW %.4f ms per %d pairs (%.2f matches per pair)" % (t_bow, B, t_sb, B, float(sb_n.float().mean())))
