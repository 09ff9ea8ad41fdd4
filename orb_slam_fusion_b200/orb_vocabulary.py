"""Host-side mirror of the reference's ORBVocabulary = DBoW2::TemplatedVocabulary<FORB>
(3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h) for the calls on the data path: loadFromTextFile
(:1246-1330) and transform (:1056-1118, 1139-1179), i.e. Frame::ComputeBoW (src/map/frame.cc:761-766),
over the C ABI (include/orbx.h, orbv_*).  The tree lives in HBM; there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

from . import _abi as A


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


class ORBVocabulary:
    """`ORBVocabulary(path=...)` = loadFromTextFile; `ORBVocabulary(k, L, parent, is_leaf, desc, weight)`
    takes the node arrays (indexed by node id, 0 = root) directly."""

    def __init__(self, k=10, L=5, parent=None, is_leaf=None, desc=None, weight=None, scoring=0, weighting=0, path=None,
                 device=0):
        self._lib = A.lib()
        self._v = A.vp()
        self.device = int(device)
        if path is not None:
            rc = self._lib.orbv_load_text(self.device, os.fsencode(path), C.byref(self._v))
        else:
            parent = np.ascontiguousarray(parent, np.int32)
            is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
            desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            weight = np.ascontiguousarray(weight, np.float64)
            assert len(parent) == len(is_leaf) == len(desc) == len(weight)
            rc = self._lib.orbv_create(self.device, k, L, scoring, weighting, len(parent), parent.ctypes.data,
                                       is_leaf.ctypes.data, desc.ctypes.data, weight.ctypes.data, C.byref(self._v))
        if rc:
            self._v = None
            raise A.OrbxError(rc, "vocabulary could not be created (bad file / arrays, or no CUDA device %d)" % self.device)
        vals = [C.c_int() for _ in range(6)]
        self._lib.orbv_info(self._v, *[C.byref(x) for x in vals])
        self.k, self.L, self.scoring, self.weighting, self.n_nodes, self.n_words = [x.value for x in vals]

    # the reference's spelling
    @classmethod
    def loadFromTextFile(cls, path, device=0):
        return cls(path=path, device=device)

    def close(self):
        if getattr(self, "_v", None):
            self._lib.orbv_destroy(self._v)
            self._v = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise A.OrbxError(rc, self._lib.orbv_last_error(self._v).decode())

    def size(self):
        return self.n_words

    def empty(self):
        return self.n_words == 0

    def launch_count(self):
        return self._lib.orbv_launch_count(self._v)

    def sync(self):
        self._check(self._lib.orbv_sync(self._v))

    def features(self, desc, levelsup=4):
        """transform(feature, word id, weight, node id, levelsup) for every row of desc [n,32]."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid, w, nid = np.zeros(n, np.uint32), np.zeros(n, np.float64), np.zeros(n, np.uint32)
        self._check(self._lib.orbv_features(self._v, desc.ctypes.data, n, levelsup, wid.ctypes.data, w.ctypes.data,
                                            nid.ctypes.data, A.MEM_HOST, None))
        return wid, w, nid

    def transform(self, desc, levelsup=4):
        """One frame: (word ids, values) of the BowVector and (node ids, [feature indices per node]) of the
        FeatureVector, both in increasing id like the reference's std::maps."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        r = self.transform_batch(desc[None], None, levelsup)
        nb, nf, tot = int(r["bow_n"][0]), int(r["fv_n"][0]), int(r["fv_total"][0])
        begin = r["fv_begin"][0]
        ends = list(begin[1:nf]) + [tot]
        return (r["bow_ids"][0, :nb].copy(), r["bow_vals"][0, :nb].copy(), r["fv_nodes"][0, :nf].copy(),
                [r["fv_feats"][0, begin[j]:ends[j]].copy() for j in range(nf)])

    def transform_batch(self, desc, n_per_frame=None, levelsup=4):
        """desc [F, cap, 32] (numpy, or a CUDA torch tensor: then every output is a CUDA tensor and the call only
        enqueues on torch's current stream), n_per_frame [F] int32 or None.  Returns a dict of the arrays
        of orbv_transform, shaped [F, cap] / [F]."""
        if _is_torch(desc):
            import torch
            assert desc.is_cuda and desc.dtype == torch.uint8 and desc.is_contiguous() and desc.shape[-1] == 32
            F, cap = desc.shape[0], desc.shape[1]
            dev = desc.device
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
            out = {"bow_ids": mk((F, cap), torch.int32), "bow_vals": mk((F, cap), torch.float64), "bow_n": mk((F,), torch.int32),
                   "fv_nodes": mk((F, cap), torch.int32), "fv_begin": mk((F, cap), torch.int32), "fv_n": mk((F,), torch.int32),
                   "fv_feats": mk((F, cap), torch.int32), "fv_total": mk((F,), torch.int32)}
            mem, stream = A.MEM_DEVICE, A.torch_stream(dev)
            if n_per_frame is not None:
                assert n_per_frame.is_cuda and n_per_frame.dtype == torch.int32
        else:
            desc = np.ascontiguousarray(desc, np.uint8)
            assert desc.ndim == 3 and desc.shape[2] == 32
            F, cap = desc.shape[0], desc.shape[1]
            out = {"bow_ids": np.zeros((F, cap), np.uint32), "bow_vals": np.zeros((F, cap), np.float64),
                   "bow_n": np.zeros(F, np.int32), "fv_nodes": np.zeros((F, cap), np.uint32),
                   "fv_begin": np.zeros((F, cap), np.int32), "fv_n": np.zeros(F, np.int32),
                   "fv_feats": np.zeros((F, cap), np.uint32), "fv_total": np.zeros(F, np.int32)}
            mem, stream = A.MEM_HOST, None
            if n_per_frame is not None:
                n_per_frame = np.ascontiguousarray(n_per_frame, np.int32)
        self._check(self._lib.orbv_transform(
            self._v, A.ptr(desc), cap, A.ptr(n_per_frame), F, levelsup, A.ptr(out["bow_ids"]), A.ptr(out["bow_vals"]),
            A.ptr(out["bow_n"]), A.ptr(out["fv_nodes"]), A.ptr(out["fv_begin"]), A.ptr(out["fv_n"]), A.ptr(out["fv_feats"]),
            A.ptr(out["fv_total"]), mem, stream))
        return out
