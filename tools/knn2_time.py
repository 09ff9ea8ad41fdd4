import sys, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
m = P.ORBmatcher()
db = P.synth_descriptors(0, 10_000_000, 7); q = P.synth_descriptors(0, 1000, 8)
for _ in range(2): m.knn2(q, db)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): m.knn2(q, db)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("knn2 1000 x 10M: %.3f ms, %.3e pairs/s; ham256 register loop peak %.3e/s" % (ms, 1e10 / (ms * 1e-3), P.popc_peak(2)))
