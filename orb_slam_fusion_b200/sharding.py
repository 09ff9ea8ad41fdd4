"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

* Extraction shards by frame -- independent units, no data-path collective.
* Database search shards by row slice: every rank scans its slice for all queries, the per-rank
  top-2 lists are all-gathered and merged by the lexicographic key (distance, global row), which
  reproduces knnMatch's tie order exactly and is associative, so the sharded result is
  bit-identical to the single-GPU one.  On GPUs the whole search is ONE C-ABI call,
  orbm_knn2_sharded (include/orbx.h): local scan, one ncclAllGather of 16 bytes per query and rank
  on the call's stream, one merge + ratio kernel.  torch.distributed only carries the 128-byte
  NCCL id once (NcclComm); the C++ host does the same with any channel it has.

The CPU tests run the slice / gather / merge plumbing over gloo with an ORBmatcher-shaped stand-in
that answers from the oracle (sharded_knn2's generic path).
"""
import ctypes as C

from . import _abi as A


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


class NcclComm:
    """An ncclComm_t over the ranks of a torch.distributed group, created through the C ABI
    (orbm_nccl_unique_id on rank 0 -> 128 bytes broadcast over the group -> orbm_nccl_comm_create),
    so the communicator belongs to the NCCL the library bound and can be handed to orbm_knn2_sharded."""

    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist
        self._lib = A.lib()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        ident = torch.zeros(A.NCCL_ID_BYTES, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_uint8 * A.NCCL_ID_BYTES)()
            rc = self._lib.orbm_nccl_unique_id(buf)
            if rc:
                raise A.OrbxError(rc, "orbm_nccl_unique_id failed (no libnccl.so.2?)")
            ident = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        on_gpu = dist.get_backend(group) == "nccl"
        if on_gpu:
            ident = ident.to(torch.device("cuda", device))
        dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(ident.cpu().numpy().tobytes())
        self.handle = A.vp()
        rc = self._lib.orbm_nccl_comm_create(raw, self.world, self.rank, int(device), C.byref(self.handle))
        if rc:
            self.handle = None
            raise A.OrbxError(rc, "orbm_nccl_comm_create failed")

    def close(self):
        if getattr(self, "handle", None):
            self._lib.orbm_nccl_comm_destroy(self.handle)
            self.handle = None


_COMMS = {}


def nccl_comm(device, group=None):
    """The cached NcclComm of (device, group)."""
    key = (int(device), id(group))
    if key not in _COMMS:
        _COMMS[key] = NcclComm(device, group)
    return _COMMS[key]


def close_comms():
    """Destroy the cached communicators (before torch.distributed.destroy_process_group)."""
    for c in _COMMS.values():
        c.close()
    _COMMS.clear()


def sharded_knn2(matcher, q, db_local, index_base, ratio=0.7, group=None):
    """2-NN + ratio test of `q` against a database whose rows are sharded over the ranks of `group`.
    `db_local` is this rank's slice and `index_base` its first global row.  Returns
    (idx[nq,2], dist[nq,2], accept[nq]) -- identical on every rank.  CUDA tensors take the C-ABI
    path (orbm_knn2_sharded, one NCCL all-gather inside the call)."""
    import torch
    import torch.distributed as dist
    if getattr(q, "is_cuda", False) and hasattr(matcher, "knn2_sharded"):
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        comm = nccl_comm(q.device.index, group).handle if world > 1 else None
        return matcher.knn2_sharded(comm, q, db_local, index_base, ratio)
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
