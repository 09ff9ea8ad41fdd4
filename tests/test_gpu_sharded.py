"""GPU: orbm_knn2_sharded through the C ABI -- one rank without NCCL, then real NCCL ranks (one process per GPU,
skipped on a box with fewer than two GPUs): every rank must hold exactly the single-GPU result of the whole
database (frame.cc:1154-1162 order: (distance, row), lowest row wins ties), planted rows in every shard."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(oracle, nq, nd, world):
    q = oracle.synth_descriptors(0, nq, 5)
    db = (oracle.synth_descriptors(0, nd, 6) & 3)                 # low entropy: many distance ties inside and across shards
    rng = np.random.default_rng(9)
    for s in range(world):                                         # a known best row in every shard, two equal ones across shards
        r0, r1 = nd * s // world, nd * (s + 1) // world
        if r1 > r0:
            db[r0 + (r1 - r0) // 2] = q[(3 * s + 1) % nq]
    db[nd - 1] = q[7 % nq]
    db[0] = q[7 % nq]
    flip = q[11 % nq].copy()
    flip[0] ^= 1
    db[nd // 2] = flip                                             # distance-1 decoy next to an exact match of query 11?  (ratio test)
    return q, db, rng


def test_single_rank_equals_knn2_and_ratio(oracle):
    import orb_slam_fusion_b200 as P
    m = P.ORBmatcher()
    q, db, _ = _case(oracle, 300, 20001, 1)
    idx, dist, acc = m.knn2_sharded(None, q, db, 0, 0.7)
    ri, rd = oracle.knn2(q, db)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rd)
    assert np.array_equal(acc, oracle.ratio_accept(ri, rd, 0.7))
    # a slice with a base: global rows come back; empty and one-row slices
    idx2, dist2, _ = m.knn2_sharded(None, q, db[5000:], 5000, 0.7)
    ri2, rd2 = oracle.knn2(q, db[5000:])
    assert np.array_equal(idx2, np.where(ri2 >= 0, ri2 + 5000, -1)) and np.array_equal(dist2, rd2)
    idx3, dist3, acc3 = m.knn2_sharded(None, q, db[:1], 0, 0.7)
    assert (idx3[:, 0] == 0).all() and (idx3[:, 1] == -1).all() and not acc3.any()
    idx4, _, acc4 = m.knn2_sharded(None, q, db[:0], 0, 0.7)
    assert (idx4 == -1).all() and not acc4.any()
    import torch
    ti, td, ta = m.knn2_sharded(None, torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), 0, 0.7)
    torch.cuda.synchronize()
    assert np.array_equal(ti.cpu().numpy(), ri) and np.array_equal(td.cpu().numpy(), rd)
    assert np.array_equal(ta.cpu().numpy().astype(bool), acc)


def _worker(rank, world, port, nq, nd, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import orb_slam_fusion_b200 as P
    from orb_slam_fusion_b200 import sharding as S
    from oracle import oracle as O
    q, db, _ = _case(O, nq, nd, world)
    r0, r1 = S.db_slice(nd, rank, world)
    m = P.ORBmatcher(device=rank)
    tq, tdb = torch.from_numpy(q).cuda(), torch.from_numpy(db[r0:r1].copy()).cuda()
    for _ in range(3):                                             # repeated collectives on the same communicator
        idx, dd, acc = S.sharded_knn2(m, tq, tdb, r0, 0.7)
    torch.cuda.synchronize()
    # host-memory form of the same collective
    hi, hd, ha = m.knn2_sharded(S.nccl_comm(rank).handle, q, db[r0:r1], r0, 0.7)
    assert np.array_equal(hi, idx.cpu().numpy()) and np.array_equal(hd, dd.cpu().numpy())
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), idx=idx.cpu().numpy(), dist=dd.cpu().numpy(), acc=acc.cpu().numpy(),
             launches=m.launch_count())
    dist.barrier()
    S.close_comms()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq,nd", [(257, 40001), (50, 3)])
def test_nccl_ranks_equal_single_gpu(oracle, tmp_path, nq, nd):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, nq, nd, str(tmp_path)), nprocs=world, join=True)
    q, db, _ = _case(oracle, nq, nd, world)
    ri, rd = oracle.knn2(q, db)
    racc = oracle.ratio_accept(ri, rd, 0.7)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % r))
        assert np.array_equal(z["idx"], ri) and np.array_equal(z["dist"], rd), r
        assert np.array_equal(z["acc"].astype(bool), racc), r
        assert z["launches"] > 0
