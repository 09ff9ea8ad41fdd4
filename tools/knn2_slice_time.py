"""One rank's share of the sharded search (1000 queries x 10M / 8 rows) on one GPU: orbm_knn2_sharded with comm = NULL."""
import sys, torch
sys.path.insert(0, '/root/repo')
import orb_slam_fusion_b200 as P
m = P.ORBmatcher()
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
db = P.synth_descriptors(0, nd, 7); q = P.synth_descriptors(0, 1000, 8)
for _ in range(3): m.knn2_sharded(None, q, db, 0, 0.7)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): m.knn2_sharded(None, q, db, 0, 0.7)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("knn2_sharded(comm=NULL) 1000 x %d: %.4f ms, %.3e pairs/s" % (nd, ms, 1000.0 * nd / (ms * 1e-3)))
