"""Host-side mirror of ORB_SLAM_FUSION::ORBmatcher's data-parallel kernels
(include/cam/orb_feature/orb_matcher.h:36-129) over the C ABI: DescriptorDistance
(orb_matcher.cc:1877-1891), knnMatch(k=2) + ratio test (frame.cc:1154-1162), the stereo row-band
search (frame.cc:836-900) and the best / second-best window search of SearchByProjection
(orb_matcher.cc:66-113).  The greedy, pointer-chasing parts of the Search* methods stay with the
caller, exactly as SURVEY.md section 3.3 scopes them."""
import ctypes as C

import numpy as np

from . import _abi as A


def _is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


class ORBmatcher:
    TH_LOW = 50        # orb_matcher.cc:35-37
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, check_ori=True, device=0):
        self.nnratio, self.check_ori, self.device = float(nnratio), bool(check_ori), int(device)
        self._lib = A.lib()
        self._m = A.vp()
        rc = self._lib.orbm_create(self.device, C.byref(self._m))
        if rc:
            self._m = None
            raise A.OrbxError(rc, "orbm_create failed (no CUDA device %d)" % self.device)

    def close(self):
        if getattr(self, "_m", None):
            self._lib.orbm_destroy(self._m)
            self._m = None

    __del__ = close

    def _check(self, rc):
        if rc:
            raise A.OrbxError(rc, self._lib.orbm_last_error(self._m).decode())

    def _stream(self, t):
        return A.torch_stream(t.device)

    def launch_count(self):
        return self._lib.orbm_launch_count(self._m)

    def sync(self):
        self._check(self._lib.orbm_sync(self._m))

    # ---- DescriptorDistance
    def DescriptorDistance(self, a, b):
        """Hamming distance of two 32-byte descriptors, or row-wise of two [n,32] arrays."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.empty(len(a), np.int32)
        self._check(self._lib.orbm_hamming_pairs(self._m, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data,
                                                 A.MEM_HOST, None))
        return int(out[0]) if len(out) == 1 else out

    # ---- brute force 2-NN + ratio
    def knn2(self, q, db, index_base=0):
        """(idx[nq,2] int64, dist[nq,2] int32) ordered by (distance, index); numpy in -> numpy out,
        CUDA tensors in -> CUDA tensors out (asynchronous on the current stream)."""
        if _is_torch(q):
            import torch
            nq, nd = q.shape[0], db.shape[0]
            idx = torch.empty((nq, 2), dtype=torch.int64, device=q.device)
            dist = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
            self._check(self._lib.orbm_knn2(self._m, q.data_ptr(), nq, db.data_ptr(), nd, index_base, idx.data_ptr(),
                                            dist.data_ptr(), A.MEM_DEVICE, self._stream(q)))
            return idx, dist
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        db = np.ascontiguousarray(db, np.uint8).reshape(-1, 32)
        idx = np.empty((len(q), 2), np.int64)
        dist = np.empty((len(q), 2), np.int32)
        self._check(self._lib.orbm_knn2(self._m, q.ctypes.data, len(q), db.ctypes.data, len(db), index_base,
                                        idx.ctypes.data, dist.ctypes.data, A.MEM_HOST, None))
        return idx, dist

    def knn2_sharded(self, comm, q, db_local, index_base=0, ratio=0.7):
        """orbm_knn2_sharded: collective 2-NN + ratio test over a database sharded by rows over the ranks of the
        NCCL communicator `comm` (an ncclComm_t address, e.g. sharding.NcclComm.handle; None = one rank).
        Returns (idx[nq,2] int64 global rows, dist[nq,2] int32, accept[nq]) -- identical on every rank."""
        if _is_torch(q):
            import torch
            assert q.dtype == torch.uint8 and db_local.dtype == torch.uint8 and q.is_contiguous() and db_local.is_contiguous()
            assert q.device == db_local.device and q.device.index == self.device, "tensors must live on the matcher's GPU"
            nq, nd = q.shape[0], db_local.shape[0]
            idx = torch.empty((nq, 2), dtype=torch.int64, device=q.device)
            dist = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
            acc = torch.empty(nq, dtype=torch.uint8, device=q.device)
            self._check(self._lib.orbm_knn2_sharded(self._m, comm, q.data_ptr(), nq, db_local.data_ptr(), nd, int(index_base),
                                                    float(ratio), idx.data_ptr(), dist.data_ptr(), acc.data_ptr(), A.MEM_DEVICE,
                                                    self._stream(q)))
            return idx, dist, acc
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        db_local = np.ascontiguousarray(db_local, np.uint8).reshape(-1, 32)
        idx = np.empty((len(q), 2), np.int64)
        dist = np.empty((len(q), 2), np.int32)
        acc = np.empty(len(q), np.uint8)
        self._check(self._lib.orbm_knn2_sharded(self._m, comm, q.ctypes.data, len(q), db_local.ctypes.data, len(db_local),
                                                int(index_base), float(ratio), idx.ctypes.data, dist.ctypes.data, acc.ctypes.data,
                                                A.MEM_HOST, None))
        return idx, dist, acc.astype(bool)

    def set_option(self, option, value):
        self._check(self._lib.orbm_set_option(self._m, int(option), int(value)))

    def top2_merge(self, idx_parts, dist_parts):
        """Merge [P,nq,2] partial top-2 lists into the global top-2."""
        if _is_torch(idx_parts):
            import torch
            P, nq = idx_parts.shape[0], idx_parts.shape[1]
            idx = torch.empty((nq, 2), dtype=torch.int64, device=idx_parts.device)
            dist = torch.empty((nq, 2), dtype=torch.int32, device=idx_parts.device)
            self._check(self._lib.orbm_top2_merge(self._m, idx_parts.data_ptr(), dist_parts.data_ptr(), P, nq,
                                                  idx.data_ptr(), dist.data_ptr(), A.MEM_DEVICE, self._stream(idx_parts)))
            return idx, dist
        idx_parts = np.ascontiguousarray(idx_parts, np.int64)
        dist_parts = np.ascontiguousarray(dist_parts, np.int32)
        P, nq = idx_parts.shape[0], idx_parts.shape[1]
        idx = np.empty((nq, 2), np.int64)
        dist = np.empty((nq, 2), np.int32)
        self._check(self._lib.orbm_top2_merge(self._m, idx_parts.ctypes.data, dist_parts.ctypes.data, P, nq,
                                              idx.ctypes.data, dist.ctypes.data, A.MEM_HOST, None))
        return idx, dist

    def ratio_test(self, idx, dist, ratio=0.7):
        if _is_torch(idx):
            import torch
            acc = torch.empty(idx.shape[0], dtype=torch.uint8, device=idx.device)
            self._check(self._lib.orbm_ratio_test(self._m, idx.data_ptr(), dist.data_ptr(), idx.shape[0], ratio,
                                                  acc.data_ptr(), A.MEM_DEVICE, self._stream(idx)))
            return acc
        idx = np.ascontiguousarray(idx, np.int64)
        dist = np.ascontiguousarray(dist, np.int32)
        acc = np.empty(len(idx), np.uint8)
        self._check(self._lib.orbm_ratio_test(self._m, idx.ctypes.data, dist.ctypes.data, len(idx), ratio,
                                              acc.ctypes.data, A.MEM_HOST, None))
        return acc.astype(bool)

    # ---- stereo row band
    def stereo_rowband(self, kl, dl, kr, dr, scale_factors, n_rows, min_d, max_d):
        kl = np.ascontiguousarray(kl, A.KP_DTYPE)
        kr = np.ascontiguousarray(kr, A.KP_DTYPE)
        dl = np.ascontiguousarray(dl, np.uint8)
        dr = np.ascontiguousarray(dr, np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        bi = np.empty(len(kl), np.int32)
        bd = np.empty(len(kl), np.int32)
        self._check(self._lib.orbm_stereo_rowband(self._m, kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data,
                                                  dr.ctypes.data, len(kr), sf.ctypes.data, len(sf), int(n_rows),
                                                  float(min_d), float(max_d), bi.ctypes.data, bd.ctypes.data,
                                                  A.MEM_HOST, None))
        return bi, bd

    def stereo_refine(self, ex_left, ex_right, kl, kr, best_idx, best_dist, min_d, max_d, bf, th_orb_dist=None):
        """frame.cc:903-985 on the two extractors' last frames: (u_right, depth, sad) per left keypoint."""
        th = (self.TH_HIGH + self.TH_LOW) // 2 if th_orb_dist is None else int(th_orb_dist)   # frame.cc:832
        kl = np.ascontiguousarray(kl, A.KP_DTYPE)
        kr = np.ascontiguousarray(kr, A.KP_DTYPE)
        bi = np.ascontiguousarray(best_idx, np.int32)
        bd = np.ascontiguousarray(best_dist, np.int32)
        ur = np.empty(len(kl), np.float32)
        dp = np.empty(len(kl), np.float32)
        sad = np.empty(len(kl), np.int32)
        self._check(self._lib.orbm_stereo_refine(self._m, ex_left._h, ex_right._h, kl.ctypes.data, len(kl), kr.ctypes.data,
                                                 len(kr), bi.ctypes.data, bd.ctypes.data, th, float(min_d), float(max_d),
                                                 float(bf), ur.ctypes.data, dp.ctypes.data, sad.ctypes.data, A.MEM_HOST, None))
        return ur, dp, sad

    def ComputeStereoMatches(self, ex_left, ex_right, kl, dl, kr, dr, bf, mb):
        """Frame::ComputeStereoMatches (frame.cc:828-986) for the extractors' last frames:
        returns (mvuRight, mvDepth).  minZ = mb, minD = 0, maxD = bf / minZ (:853-856)."""
        n_rows = ex_left.pyramid_level_size(0)[1]
        min_d, max_d = 0.0, float(np.float32(bf) / np.float32(mb))
        bi, bd = self.stereo_rowband(kl, dl, kr, dr, ex_left.GetScaleFactors(), n_rows, min_d, max_d)
        ur, dp, _ = self.stereo_refine(ex_left, ex_right, kl, kr, bi, bd, min_d, max_d, bf)
        return ur, dp

    def stereo_matches_last(self, ex_left, ex_right, bf, mb, u_right=None, depth=None):
        """Frame::ComputeStereoMatches (frame.cc:828-986) as one call on what the two extractors left on the device with
        their last single-frame call (orbm_stereo_matches_last): nothing is uploaded.  Returns (mvuRight, mvDepth)."""
        cap = ex_left.max_keypoints() if u_right is None else len(u_right)
        ur = np.empty(cap, np.float32) if u_right is None else u_right
        dp = np.empty(cap, np.float32) if depth is None else depth
        nl = C.c_int(0)
        self._check(self._lib.orbm_stereo_matches_last(self._m, ex_left._h, ex_right._h, float(bf), float(mb), ur.ctypes.data,
                                                       dp.ctypes.data, cap, C.byref(nl)))
        return ur[:nl.value], dp[:nl.value]

    def ComputeDistinctiveDescriptors(self, desc, offsets):
        """mappoint.cc:365-428 for a batch of map points: (best row per point, its median distance)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        offsets = np.ascontiguousarray(offsets, np.int32)
        n = len(offsets) - 1
        bi = np.empty(n, np.int32)
        bm = np.empty(n, np.int32)
        max_rows = int(np.diff(offsets).max()) if n else 0
        self._check(self._lib.orbm_distinctive(self._m, desc.ctypes.data, offsets.ctypes.data, n, max_rows, bi.ctypes.data,
                                               bm.ctypes.data, A.MEM_HOST, None))
        return bi, bm

    def window_search_fuse(self, kps, desc, geom, queries, qdesc, inv_level_sigma2, kp_u_right=None, q_u_right=None):
        """The search of ORBmatcher::Fuse (orb_matcher.cc:1130-1187): nearest keypoint per projected map point under the
        chi-square reprojection gate (7.8 with a right coordinate, 5.99 without)."""
        kps = np.ascontiguousarray(kps, A.KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8)
        queries = np.ascontiguousarray(queries, A.WQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
        ur = qr = None
        if kp_u_right is not None:
            ur, qr = np.ascontiguousarray(kp_u_right, np.float32), np.ascontiguousarray(q_u_right, np.float32)
        out = np.empty(len(queries), A.WR_DTYPE)
        g = A.GridGeom(*geom)
        self._check(self._lib.orbm_window_search_fuse(self._m, kps.ctypes.data, desc.ctypes.data, len(kps), C.byref(g), queries.ctypes.data,
                                                      qdesc.ctypes.data, len(queries), A.ptr(ur), A.ptr(qr), inv.ctypes.data, len(inv),
                                                      out.ctypes.data, A.MEM_HOST, None))
        return out

    def SearchByProjection(self, kps, desc, geom, queries, qdesc, skip=None, kp_u_right=None, q_u_right=None, q_max_err=None,
                           th_high=None, nnratio=None):
        """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, ...) (orb_matcher.cc:42-134, Nleft == -1)
        as a whole, greedy claim included: queries = the windows of the map points that passed :50-57, in order.
        Returns (nmatches, assigned[n]): assigned[i] = query whose map point frame keypoint i receives, -1 = untouched."""
        kps = np.ascontiguousarray(kps, A.KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8)
        queries = np.ascontiguousarray(queries, A.WQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
        ur = qr = qe = None
        if kp_u_right is not None:
            ur, qr, qe = (np.ascontiguousarray(a, np.float32) for a in (kp_u_right, q_u_right, q_max_err))
            assert len(ur) == len(kps) and len(qr) == len(qe) == len(queries)
        assigned = np.empty(max(len(kps), 1), np.int32)
        nm = C.c_int32(0)
        g = A.GridGeom(*geom)
        self._check(self._lib.orbm_search_by_projection(
            self._m, kps.ctypes.data, desc.ctypes.data, len(kps), C.byref(g), queries.ctypes.data, qdesc.ctypes.data, len(queries),
            A.ptr(sk), A.ptr(ur), A.ptr(qr), A.ptr(qe), self.TH_HIGH if th_high is None else int(th_high),
            self.nnratio if nnratio is None else float(nnratio), assigned.ctypes.data, C.addressof(nm), A.MEM_HOST, None))
        return nm.value, assigned[:len(kps)].copy()

    def SearchByProjectionLast(self, kps, desc, geom, queries, qdesc, q_angle, skip=None, kp_u_right=None, q_u_right=None,
                               q_max_err=None, th_high=None, check_orientation=None):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (orb_matcher.cc:1518-1728, Nleft == -1) after
        the projection: queries = the windows of the last frame's map points that project into the frame, in order,
        q_angle their keypoint angles.  Returns (nmatches, assigned[n])."""
        kps = np.ascontiguousarray(kps, A.KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8)
        queries = np.ascontiguousarray(queries, A.WQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        qa = np.ascontiguousarray(q_angle, np.float32)
        sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
        ur = qr = qe = None
        if kp_u_right is not None:
            ur, qr, qe = (np.ascontiguousarray(a, np.float32) for a in (kp_u_right, q_u_right, q_max_err))
        assigned = np.empty(max(len(kps), 1), np.int32)
        nm = C.c_int32(0)
        g = A.GridGeom(*geom)
        self._check(self._lib.orbm_search_by_projection_last(
            self._m, kps.ctypes.data, desc.ctypes.data, len(kps), C.byref(g), queries.ctypes.data, qdesc.ctypes.data, qa.ctypes.data,
            len(queries), A.ptr(sk), A.ptr(ur), A.ptr(qr), A.ptr(qe), self.TH_HIGH if th_high is None else int(th_high),
            int(self.check_ori if check_orientation is None else check_orientation), assigned.ctypes.data, C.addressof(nm),
            A.MEM_HOST, None))
        return nm.value, assigned[:len(kps)].copy()

    # ---- bag-of-words guided matching
    def SearchByBoWKeyFrames(self, kps, desc, n_per_frame, fv, pairs, has_point=None, nnratio=0.7, check_orientation=True):
        """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cc:697-815, loop closing): same pool and
        pair layout as SearchByBoW; has_point gates both sides; match[p, i] = feature of key frame 2 matched to
        feature i of key frame 1, -1 = none."""
        return self.SearchByBoW(kps, desc, n_per_frame, fv, pairs, has_point, nnratio, check_orientation, _keyframes=True)

    def SearchByBoW(self, kps, desc, n_per_frame, fv, pairs, has_point=None, nnratio=0.7, check_orientation=True, _keyframes=False):
        """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (orb_matcher.cc:215-389) for a batch of
        (key frame, frame) pairs out of one pool of frames in the [frame][cap] layout of OrbExtractor.extract_batch and
        ORBVocabulary.transform_batch: kps [F, cap] keypoint records, desc [F, cap, 32], n_per_frame [F] (or None),
        fv = the dict transform_batch returned, pairs [n_pairs, 2] int32 (key frame, frame), has_point [F, cap] uint8
        or None.  Returns (n_matches [n_pairs], match [n_pairs, cap]): match[p, i] = key-frame feature whose map point
        frame feature i receives, -1 = none.  numpy in -> numpy out; CUDA torch tensors in -> CUDA tensors out, the call
        only enqueues on torch's current stream."""
        if type(desc).__module__.startswith("torch"):
            import torch
            F, cap = desc.shape[0], desc.shape[1]
            dev = desc.device
            assert desc.is_cuda and desc.is_contiguous() and kps.is_cuda and kps.is_contiguous() and kps.element_size() * kps[0, 0].numel() == 28
            pairs = pairs.to(device=dev, dtype=torch.int32)
            kf, ff = pairs[:, 0].contiguous(), pairs[:, 1].contiguous()
            match = torch.empty((len(kf), cap), dtype=torch.int32, device=dev)
            nm = torch.empty(len(kf), dtype=torch.int32, device=dev)
            mem, stream = A.MEM_DEVICE, A.torch_stream(dev)
            # the C ABI takes plain pointers: a tensor of another width, layout or device would be reinterpreted silently
            for key in ("fv_nodes", "fv_begin", "fv_n", "fv_feats", "fv_total"):
                t = fv[key]
                assert t.is_cuda and t.device == dev and t.is_contiguous() and t.element_size() == 4, key
            if n_per_frame is not None:
                assert n_per_frame.device == dev and n_per_frame.dtype == torch.int32 and n_per_frame.is_contiguous()
            if has_point is not None:
                assert has_point.device == dev and has_point.element_size() == 1 and has_point.is_contiguous() and tuple(has_point.shape) == (F, cap)
        else:
            kps = np.ascontiguousarray(kps, A.KP_DTYPE)
            F, cap = kps.shape
            fv = {k: np.ascontiguousarray(v) for k, v in fv.items()}
            for key in ("fv_nodes", "fv_begin", "fv_n", "fv_feats", "fv_total"):
                assert fv[key].dtype.itemsize == 4, key
            desc = np.ascontiguousarray(desc, np.uint8).reshape(F, cap, 32)
            n_per_frame = None if n_per_frame is None else np.ascontiguousarray(n_per_frame, np.int32)
            has_point = None if has_point is None else np.ascontiguousarray(has_point, np.uint8).reshape(F, cap)
            pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
            kf, ff = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
            match = np.empty((len(kf), cap), np.int32)
            nm = np.empty(len(kf), np.int32)
            mem, stream = A.MEM_HOST, None
        for key in ("fv_nodes", "fv_begin", "fv_feats"):
            assert tuple(fv[key].shape) == (F, cap), key
        fn = self._lib.orbm_search_by_bow_kf if _keyframes else self._lib.orbm_search_by_bow
        self._check(fn(self._m, A.ptr(kps), A.ptr(desc), cap, F, A.ptr(n_per_frame), A.ptr(fv["fv_nodes"]), A.ptr(fv["fv_begin"]),
                       A.ptr(fv["fv_n"]), A.ptr(fv["fv_feats"]), A.ptr(fv["fv_total"]), A.ptr(has_point), A.ptr(kf), A.ptr(ff), len(kf),
                       nnratio, int(check_orientation), A.ptr(match), A.ptr(nm), mem, stream))
        return nm, match

    def SearchForTriangulation(self, kps, desc, n_per_frame, fv, pairs, has_point, u_right, pair_f12, pair_ep, scale_factors,
                               level_sigma2, only_stereo=False, coarse=False, check_orientation=None):
        """ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040) for a batch of key-frame pairs of one frame pool
        (layout of SearchByBoW); has_point / u_right [F, cap], pair_f12 [n_pairs, 3, 3] (Pinhole::EpipolarConstrain's F12),
        pair_ep [n_pairs, 2] (epipole in image 2).  Returns (n_matches [n_pairs], match [n_pairs, cap]): match[p, i] =
        feature of key frame 2 matched to feature i of key frame 1, -1 = none."""
        kps = np.ascontiguousarray(kps, A.KP_DTYPE)
        F, cap = kps.shape
        desc = np.ascontiguousarray(desc, np.uint8).reshape(F, cap, 32)
        npf = None if n_per_frame is None else np.ascontiguousarray(n_per_frame, np.int32)
        hp = np.ascontiguousarray(has_point, np.uint8).reshape(F, cap)
        ur = np.ascontiguousarray(u_right, np.float32).reshape(F, cap)
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        p1, p2 = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
        f12 = np.ascontiguousarray(pair_f12, np.float32).reshape(len(pairs), 9)
        ep = np.ascontiguousarray(pair_ep, np.float32).reshape(len(pairs), 2)
        sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
        match = np.empty((len(pairs), cap), np.int32)
        nm = np.empty(len(pairs), np.int32)
        self._check(self._lib.orbm_search_for_triangulation(
            self._m, A.ptr(kps), A.ptr(desc), cap, F, A.ptr(npf), A.ptr(fv["fv_nodes"]), A.ptr(fv["fv_begin"]), A.ptr(fv["fv_n"]),
            A.ptr(fv["fv_feats"]), A.ptr(fv["fv_total"]), A.ptr(hp), A.ptr(ur), A.ptr(p1), A.ptr(p2), len(pairs), A.ptr(f12), A.ptr(ep),
            A.ptr(sf), A.ptr(s2), len(sf), int(only_stereo), int(coarse),
            int(self.check_ori if check_orientation is None else check_orientation), A.ptr(match), A.ptr(nm), A.MEM_HOST, None))
        return nm, match

    # ---- projection window
    def window_search(self, kps, desc, geom, queries, qdesc, skip=None, kp_u_right=None, q_u_right=None, q_max_err=None):
        """Best / second-best frame keypoint per projected map point (orb_matcher.cc:66-113, 451-479, 1567-1608).
        kp_u_right (Frame::mvuRight) with q_u_right / q_max_err per query adds the stereo gate of :89-92 / :1586-1590."""
        kps = np.ascontiguousarray(kps, A.KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8)
        queries = np.ascontiguousarray(queries, A.WQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8)
        sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
        ur = qr = qe = None
        if kp_u_right is not None:
            ur = np.ascontiguousarray(kp_u_right, np.float32)
            qr = np.ascontiguousarray(q_u_right, np.float32)
            qe = np.ascontiguousarray(q_max_err, np.float32)
            assert len(ur) == len(kps) and len(qr) == len(qe) == len(queries)
        out = np.empty(len(queries), A.WR_DTYPE)
        g = A.GridGeom(*geom)
        self._check(self._lib.orbm_window_search_stereo(
            self._m, kps.ctypes.data, desc.ctypes.data, len(kps), C.byref(g), queries.ctypes.data, qdesc.ctypes.data,
            len(queries), A.ptr(sk), A.ptr(ur), A.ptr(qr), A.ptr(qe), out.ctypes.data, A.MEM_HOST, None))
        return out


def synth_descriptors(first, n, seed, device=0, out=None):
    """[n,32] uint8 CUDA tensor of SURVEY.md 8(d) config-5 descriptors."""
    import torch
    dev = torch.device("cuda", device)
    if out is None:
        out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    rc = A.lib().orbm_synth_descriptors(device, out.data_ptr(), first, n, seed, A.torch_stream(dev))
    if rc:
        raise A.OrbxError(rc, "orbm_synth_descriptors")
    return out


def popc_peak(mode=0, device=0):
    v = C.c_double()
    rc = A.lib().orbm_popc_peak(device, mode, C.byref(v))
    if rc:
        raise A.OrbxError(rc, "orbm_popc_peak")
    return v.value
