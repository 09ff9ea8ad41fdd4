"""CPU, world_size 2 and 3 over gloo: the multi-GPU host logic of orb_slam_fusion_b200/sharding.py.
The compute calls are answered by an oracle-backed stand-in (this is a test of the plumbing: slice
bounds, index bases, gather layout, merge order); the real kernels run in tests/test_gpu_match.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleMatcher:
    """ORBmatcher-shaped stand-in on CPU tensors."""

    def __init__(self):
        from oracle import oracle as O
        self.O = O

    def knn2(self, q, db, index_base=0):
        idx, d = self.O.knn2(q.numpy(), db.numpy())
        idx = np.where(idx >= 0, idx + index_base, -1)
        return torch.from_numpy(idx), torch.from_numpy(d)

    def top2_merge(self, gi, gd):
        gi, gd = gi.numpy(), gd.numpy()
        P, nq = gi.shape[0], gi.shape[1]
        oi = np.full((nq, 2), -1, np.int64)
        od = np.full((nq, 2), np.iinfo(np.int32).max, np.int32)
        for qi in range(nq):
            c = sorted((int(gd[p, qi, k]), int(gi[p, qi, k])) for p in range(P) for k in range(2) if gi[p, qi, k] >= 0)
            for k, (d, i) in enumerate(c[:2]):
                oi[qi, k], od[qi, k] = i, d
        return torch.from_numpy(oi), torch.from_numpy(od)

    def ratio_test(self, idx, dd, ratio):
        return torch.from_numpy(self.O.ratio_accept(idx.numpy(), dd.numpy(), ratio))


def worker(rank, world, port, nd, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from orb_slam_fusion_b200 import sharding as S
    q = O.synth_descriptors(0, 50, 5)
    db = (O.synth_descriptors(0, nd, 6) & 3)               # low entropy: many distance ties
    db[nd - 1] = q[7]
    db[0] = q[7]                                           # the same best row in the first and the last shard
    r0, r1 = S.db_slice(nd, rank, world)
    idx, dd, acc = S.sharded_knn2(OracleMatcher(), torch.from_numpy(q), torch.from_numpy(db[r0:r1]), r0, 0.7)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), idx=idx.numpy(), dist=dd.numpy(), acc=acc.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nd", [(2, 4001), (3, 1000), (2, 3)])
def test_sharded_knn2_equals_single(oracle, tmp_path, world, nd):
    port = 29500 + (os.getpid() * 7 + world * 13 + nd) % 2000
    mp.spawn(worker, args=(world, port, nd, str(tmp_path)), nprocs=world, join=True)
    q = oracle.synth_descriptors(0, 50, 5)
    db = oracle.synth_descriptors(0, nd, 6) & 3
    db[nd - 1] = q[7]
    db[0] = q[7]
    wi, wd = oracle.knn2(q, db)
    wacc = oracle.ratio_accept(wi, wd, 0.7)
    assert wi[7, 0] == 0 and wi[7, 1] == nd - 1 and wd[7, 0] == 0      # tie broken by the lower global row
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(g["idx"], wi) and np.array_equal(g["dist"], wd) and np.array_equal(g["acc"], wacc)


def test_frame_sharding_covers_every_frame_once():
    from orb_slam_fusion_b200 import sharding as S
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 64, 4096):
            parts = [S.local_frames(n, r, world) for r in range(world)]
            assert sorted(f for p in parts for f in p) == list(range(n))
            assert all(S.frame_owner(f, world) == r for r, p in enumerate(parts) for f in p)
            assert S.interleave_frames(parts) == list(range(n))
        bounds = [S.db_slice(10_000_000, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == 10_000_000
        assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
