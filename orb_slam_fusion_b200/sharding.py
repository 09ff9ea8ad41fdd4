"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

* Extraction shards by frame -- independent units, no data-path collective.
* Database search shards by row slice: every rank scans its slice for all queries, the per-rank
  top-2 lists (nq x 2 x 12 bytes) are all-gathered (NCCL over NVLink on GPUs) and merged by the
  lexicographic key (distance, global row), which reproduces knnMatch's tie order exactly and is
  associative, so the sharded result is bit-identical to the single-GPU one.

The compute calls go through an ORBmatcher-shaped object (orb_matcher.ORBmatcher on a GPU); the
CPU tests run the same plumbing over gloo with a stand-in that answers from the oracle.
"""


def frame_owner(frame, world):
    """Frame f is extracted by rank f mod G (BASELINE config 4)."""
    return frame % world


def local_frames(n_frames, rank, world):
    """Frames of `rank`, ascending."""
    return list(range(rank, n_frames, world))


def interleave_frames(parts):
    """Per-rank result lists (rank r holds frames r, r+G, ...) back into frame order."""
    world = len(parts)
    n = sum(len(p) for p in parts)
    out = [None] * n
    for r, p in enumerate(parts):
        for k, v in enumerate(p):
            out[r + k * world] = v
    return out


def db_slice(n_rows, rank, world):
    """Rows [begin, end) of the database held by `rank` (contiguous, balanced)."""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def sharded_knn2(matcher, q, db_local, index_base, ratio=0.7, group=None):
    """2-NN + ratio test of `q` against a database whose rows are sharded over the ranks of `group`.
    `db_local` is this rank's slice and `index_base` its first global row.  Returns
    (idx[nq,2], dist[nq,2], accept[nq]) -- identical on every rank."""
    import torch
    import torch.distributed as dist
    idx, dd = matcher.knn2(q, db_local, index_base)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        nq = idx.shape[0]
        gi = torch.empty((world * nq, 2), dtype=idx.dtype, device=idx.device)   # rank-major concatenation
        gd = torch.empty((world * nq, 2), dtype=dd.dtype, device=dd.device)
        dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
        dist.all_gather_into_tensor(gd, dd.contiguous(), group=group)
        idx, dd = matcher.top2_merge(gi.view(world, nq, 2), gd.view(world, nq, 2))
    return idx, dd, matcher.ratio_test(idx, dd, ratio)
