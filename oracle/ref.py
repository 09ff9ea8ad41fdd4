"""ctypes binding of oracle/_ref/liborb_ref.so -- TEST INFRASTRUCTURE.

The library is the reference's own src/cam/orb_feature/orb_extractor.cc compiled unmodified on
the mini-cv shim (oracle/minicv), plus DescriptorDistance; _ref/libbow_ref.so is the reference's vendored DBoW2 and
_ref/libframe_ref.so the reference's own lines of Frame::ComputeStereoMatches, the frame grid,
MapPoint::ComputeDistinctiveDescriptors and ORBmatcher::SearchByProjection (oracle/ref_frame_shim.cc).  It is built in the development
container (`make -C oracle ref`, needs /root/reference) and travels to the GPU box as a binary.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import KP_DTYPE, _p, _u8img

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liborb_ref.so")
_LIB = None


def available(try_build=True):
    if os.path.exists(_SO):
        return True
    if try_build and os.path.isdir("/root/reference/src"):
        try:
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
        except Exception:
            return False
        return os.path.exists(_SO)
    return False


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/liborb_ref.so not built (needs /root/reference)")
        L = C.CDLL(_SO)
        vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.ref_create.restype = vp
        L.ref_create.argtypes = [i, f, i, i, i]
        L.ref_destroy.argtypes = [vp]
        L.ref_extract.argtypes = [vp, vp, i, i, sz, i, i, vp, vp, i, C.POINTER(i)]
        L.ref_level.argtypes = [vp, i, vp, sz, C.POINTER(i), C.POINTER(i)]
        L.ref_pyramid.argtypes = [vp, vp, i, i, sz]
        L.ref_octree.argtypes = [vp, vp, i, i, i, i, i, i, i, vp, i]
        L.ref_tables.argtypes = [vp, vp, vp, vp, vp]
        L.ref_hamming.argtypes = [vp, vp]
        L.ref_knn2.argtypes = [vp, i, vp, C.c_longlong, vp, vp, i]
        _LIB = L
    return _LIB


class Extractor:
    def __init__(self, num_feats=1000, scale_factor=1.2, num_levs=8, ini_th_fast=20, min_th_fast=7):
        self.h = lib().ref_create(num_feats, scale_factor, num_levs, ini_th_fast, min_th_fast)
        self.num_feats, self.num_levs = num_feats, num_levs

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def __call__(self, img, lapping=(0, 0)):
        if img is None or img.size == 0:
            n = C.c_int()
            rc = lib().ref_extract(self.h, None, 0, 0, 0, 0, 0, None, None, 0, C.byref(n))
            return rc, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        img = _u8img(img)
        cap = self.num_feats * 2 + 64 * self.num_levs
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int()
        rc = lib().ref_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                               int(lapping[0]), int(lapping[1]), _p(kps), _p(desc), cap, C.byref(n))
        assert n.value <= cap
        return rc, kps[:n.value].copy(), desc[:n.value].copy()

    def compute_pyramid(self, img):
        img = _u8img(img)
        lib().ref_pyramid(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0])
        return [self.level(l) for l in range(self.num_levs)]

    def level(self, lev, with_border=False):
        w, h = C.c_int(), C.c_int()
        lib().ref_level(self.h, lev, None, 0, C.byref(w), C.byref(h))
        out = np.empty((h.value + 38, w.value + 38), np.uint8)
        lib().ref_level(self.h, lev, _p(out), out.strides[0], C.byref(w), C.byref(h))
        return out if with_border else out[19:-19, 19:-19].copy()

    def tables(self):
        L = self.num_levs
        a, b, c, d = (np.empty(L, np.float32) for _ in range(4))
        lib().ref_tables(self.h, _p(a), _p(b), _p(c), _p(d))
        return dict(scale=a, inv_scale=b, sigma2=c, inv_sigma2=d)

    def octree(self, xyr, w, h, quota, lev=0):
        xyr = np.ascontiguousarray(xyr, np.int32).reshape(-1, 3)
        cap = quota + 8 + len(xyr)
        out = np.empty((cap, 3), np.int32)
        n = lib().ref_octree(self.h, _p(xyr), len(xyr), 16, w - 16, 16, h - 16, quota, lev, _p(out), cap)
        return out[:n].copy()


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().ref_hamming(_p(a), _p(b))


def knn2(q, d, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    d = np.ascontiguousarray(d, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), 2), np.int64)
    dist = np.empty((len(q), 2), np.int32)
    lib().ref_knn2(_p(q), len(q), _p(d), len(d), _p(idx), _p(dist), nthreads)
    return idx, dist


# ---------------------------------------------------------------- the reference's DBoW2 (libbow_ref.so)
_BOW_SO = os.path.join(_HERE, "_ref", "libbow_ref.so")
_BOW = None


def bow_available():
    return available() and os.path.exists(_BOW_SO)


def _bow():
    global _BOW
    if _BOW is None:
        if not bow_available():
            raise RuntimeError("oracle/_ref/libbow_ref.so not built (needs /root/reference)")
        L = C.CDLL(_BOW_SO)
        vp, i = C.c_void_p, C.c_int
        L.refbow_load_text.restype = vp
        L.refbow_load_text.argtypes = [C.c_char_p]
        L.refbow_destroy.argtypes = [vp]
        L.refbow_words.argtypes = [vp]
        L.refbow_transform.argtypes = [vp, vp, i, i, vp, vp, C.POINTER(i), vp, vp, C.POINTER(i), vp, C.POINTER(i)]
        L.refbow_words_of.argtypes = [vp, vp, i, vp]
        _BOW = L
    return _BOW


class Vocabulary:
    """The reference's ORBVocabulary (TemplatedVocabulary<FORB>) loaded with its own loadFromTextFile."""

    def __init__(self, path):
        self.h = _bow().refbow_load_text(os.fsencode(path))
        if not self.h:
            raise ValueError("loadFromTextFile failed")

    def __del__(self):
        if getattr(self, "h", None):
            _bow().refbow_destroy(self.h)
            self.h = None

    @property
    def n_words(self):
        return _bow().refbow_words(self.h)

    def words(self, desc):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        out = np.zeros(len(desc), np.uint32)
        _bow().refbow_words_of(self.h, _p(desc), len(desc), _p(out))
        return out

    def transform(self, desc, levelsup=4):
        from .oracle import unpack_bow
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        m = max(n, 1)
        ids, vals = np.zeros(m, np.uint32), np.zeros(m, np.float64)
        nodes, begin, feats = np.zeros(m, np.uint32), np.zeros(m, np.int32), np.zeros(m, np.uint32)
        nb, nf, tot = C.c_int(), C.c_int(), C.c_int()
        _bow().refbow_transform(self.h, _p(desc), n, levelsup, _p(ids), _p(vals), C.byref(nb), _p(nodes), _p(begin),
                                C.byref(nf), _p(feats), C.byref(tot))
        return unpack_bow(n, ids, vals, nb.value, nodes, begin, nf.value, feats, tot.value)


# ---------------------------------------------------------------- the reference's frame / map-point / matcher lines (libframe_ref.so)
_FRAME_SO = os.path.join(_HERE, "_ref", "libframe_ref.so")
_FRAME = None

TRACK_POINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("depth", "<f4"),
                              ("level", "<i4"), ("in_view", "<i4"), ("bad", "<i4")])


def frame_available():
    return available() and os.path.exists(_FRAME_SO)


def _frame():
    global _FRAME
    if _FRAME is None:
        if not frame_available():
            raise RuntimeError("oracle/_ref/libframe_ref.so not built (needs /root/reference)")
        L = C.CDLL(_FRAME_SO)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        # (tests/test_cpp_matcher.py binds this module to a library that exports only the matcher entry points)
        for name, at in [("reff_stereo_matches", [vp, vp, i, vp, i, vp, vp, i, vp, vp, vp, f, f, vp, vp]),
                         ("reff_distinctive", [vp, vp, i, vp, vp]),
                         ("reff_search_by_projection", [vp, vp, i, vp, f, f, f, f, vp, i, vp, vp, i, vp, f, f, i, f, vp]),
                         ("reff_features_in_area", [vp, i, f, f, f, f, f, f, f, i, i, vp, i]),
                         ("reff_search_by_bow_kf", [vp, vp, i, vp, vp, vp, i, vp, i, vp, vp, i, vp, vp, vp, i, vp, i, f, i, vp]),
                         ("reff_search_by_bow", [vp, vp, i, vp, vp, vp, i, vp, i, vp, vp, i, vp, vp, i, vp, i, f, i, vp])]:
            if hasattr(L, name):
                getattr(L, name).argtypes = at
        _FRAME = L
    return _FRAME


def stereo_matches(levels_left, levels_right, kl, dl, kr, dr, scale_factors, inv_scale_factors, bf, mb):
    """Frame::ComputeStereoMatches (frame.cc:828-986).  levels_*: per-level arrays WITH the 19-px border."""
    from .oracle import LevelView
    keep = []

    def views(levels):
        arr = (LevelView * len(levels))()
        for j, a in enumerate(levels):
            a = np.ascontiguousarray(a, np.uint8)
            keep.append(a)
            arr[j] = LevelView(a.ctypes.data + 19 * a.strides[0] + 19, a.shape[1] - 38, a.shape[0] - 38, a.strides[0])
        return arr

    vl, vr = views(levels_left), views(levels_right)
    kl, kr = np.ascontiguousarray(kl, KP_DTYPE), np.ascontiguousarray(kr, KP_DTYPE)
    dl, dr = np.ascontiguousarray(dl, np.uint8), np.ascontiguousarray(dr, np.uint8)
    sf, isf = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(inv_scale_factors, np.float32)
    ur, dp = np.empty(len(kl), np.float32), np.empty(len(kl), np.float32)
    _frame().reff_stereo_matches(C.cast(vl, C.c_void_p), C.cast(vr, C.c_void_p), len(levels_left), _p(kl), len(kl), _p(dl),
                                 _p(kr), len(kr), _p(dr), _p(sf), _p(isf), float(bf), float(mb), _p(ur), _p(dp))
    return ur, dp


def distinctive(desc, offsets):
    """MapPoint::ComputeDistinctiveDescriptors (mappoint.cc:365-433) per point: (chosen descriptor rows, has-descriptor flags)."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    n = len(offsets) - 1
    out = np.zeros((n, 32), np.uint8)
    chosen = np.zeros(n, np.int32)
    _frame().reff_distinctive(_p(desc), _p(offsets), n, _p(out), _p(chosen))
    return out, chosen


def search_by_projection(keys_un, desc, bounds, scale_factors, points, point_desc, pre_matched=None, u_right=None, th=3.0,
                         nnratio=0.8, far_points=False, th_far=50.0):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
    (orb_matcher.cc:42-206) on a frame with Nleft == -1.  Returns (nmatches, assigned[n])."""
    keys_un = np.ascontiguousarray(keys_un, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    points = np.ascontiguousarray(points, TRACK_POINT_DTYPE)
    point_desc = np.ascontiguousarray(point_desc, np.uint8)
    pm = None if pre_matched is None else np.ascontiguousarray(pre_matched, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    assigned = np.empty(len(keys_un), np.int32)
    nm = _frame().reff_search_by_projection(_p(keys_un), _p(desc), len(keys_un), None if ur is None else _p(ur),
                                            *[float(b) for b in bounds], _p(sf), len(sf), _p(points), _p(point_desc), len(points),
                                            None if pm is None else _p(pm), float(th), float(nnratio), int(far_points), float(th_far),
                                            _p(assigned))
    return nm, assigned


def features_in_area(keys_un, bounds, x, y, r, min_level=-1, max_level=-1):
    """Frame::GetFeaturesInArea (frame.cc:679-746) on the grid of AssignFeaturesToGrid (:438-465): indices in visiting order."""
    keys_un = np.ascontiguousarray(keys_un, KP_DTYPE)
    out = np.empty(len(keys_un) + 1, np.int32)
    n = _frame().reff_features_in_area(_p(keys_un), len(keys_un), *[float(b) for b in bounds], float(x), float(y), float(r),
                                       int(min_level), int(max_level), _p(out), len(out))
    return out[:n].copy()


def search_by_bow(kps_kf, desc_kf, has_point_kf, fv_kf, kps_f, desc_f, fv_f, nnratio=0.7, check_orientation=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (orb_matcher.cc:215-389 + ComputeThreeMaxima
    :1841-1873).  fv_* = (nodes u32, begin i32, feats u32).  Returns (nmatches, match_of_f[n_f])."""
    kps_kf, kps_f = np.ascontiguousarray(kps_kf, KP_DTYPE), np.ascontiguousarray(kps_f, KP_DTYPE)
    desc_kf, desc_f = np.ascontiguousarray(desc_kf, np.uint8), np.ascontiguousarray(desc_f, np.uint8)
    hp = None if has_point_kf is None else np.ascontiguousarray(has_point_kf, np.uint8)
    (nk, bk, fk), (nf, bf_, ff) = fv_kf, fv_f
    fk2 = fk if len(fk) else np.zeros(1, np.uint32)
    ff2 = ff if len(ff) else np.zeros(1, np.uint32)
    out = np.empty(max(len(kps_f), 1), np.int32)
    nm = _frame().reff_search_by_bow(_p(kps_kf), _p(desc_kf), len(kps_kf), None if hp is None else _p(hp), _p(nk), _p(bk), len(nk),
                                     _p(fk2), len(fk), _p(kps_f), _p(desc_f), len(kps_f), _p(nf), _p(bf_), len(nf), _p(ff2), len(ff),
                                     float(nnratio), int(check_orientation), _p(out))
    return nm, out[:len(kps_f)].copy()


def search_by_bow_kf(kps1, desc1, has_point1, fv1, kps2, desc2, has_point2, fv2, nnratio=0.7, check_orientation=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cc:697-815).  Returns (nmatches, match_of_1[n1])."""
    kps1, kps2 = np.ascontiguousarray(kps1, KP_DTYPE), np.ascontiguousarray(kps2, KP_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    hp1 = None if has_point1 is None else np.ascontiguousarray(has_point1, np.uint8)
    hp2 = None if has_point2 is None else np.ascontiguousarray(has_point2, np.uint8)
    (n1, b1, f1), (n2, b2, f2) = fv1, fv2
    f1p = f1 if len(f1) else np.zeros(1, np.uint32)
    f2p = f2 if len(f2) else np.zeros(1, np.uint32)
    out = np.empty(max(len(kps1), 1), np.int32)
    nm = _frame().reff_search_by_bow_kf(_p(kps1), _p(desc1), len(kps1), None if hp1 is None else _p(hp1), _p(n1), _p(b1), len(n1),
                                        _p(f1p), len(f1), _p(kps2), _p(desc2), len(kps2), None if hp2 is None else _p(hp2), _p(n2),
                                        _p(b2), len(n2), _p(f2p), len(f2), float(nnratio), int(check_orientation), _p(out))
    return nm, out[:len(kps1)].copy()


def search_by_projection_last(keys_un, desc, bounds, scale_factors, bf, mb, cam4, t_cw, t_lw, last_keys, last_has_point, last_outlier,
                              last_world, last_desc, pre_matched=None, u_right=None, th=7.0, mono=False, check_orientation=True):
    """ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (orb_matcher.cc:1518-1728) with
    translation-only poses and a pinhole camera (fx, fy, cx, cy).  Returns (nmatches, assigned[n])."""
    L = _frame()
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.reff_search_by_projection_last.argtypes = [vp, vp, i, vp, f, f, f, f, vp, i, f, f, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, f, i, i, vp]
    keys_un, last_keys = np.ascontiguousarray(keys_un, KP_DTYPE), np.ascontiguousarray(last_keys, KP_DTYPE)
    desc, last_desc = np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(last_desc, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    cam4, t_cw, t_lw = (np.ascontiguousarray(a, np.float32) for a in (cam4, t_cw, t_lw))
    hp, ol = np.ascontiguousarray(last_has_point, np.uint8), np.ascontiguousarray(last_outlier, np.uint8)
    world = np.ascontiguousarray(last_world, np.float32).reshape(-1, 3)
    pm = None if pre_matched is None else np.ascontiguousarray(pre_matched, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    assigned = np.empty(max(len(keys_un), 1), np.int32)
    nm = L.reff_search_by_projection_last(_p(keys_un), _p(desc), len(keys_un), None if ur is None else _p(ur),
                                          *[float(b) for b in bounds], _p(sf), len(sf), float(bf), float(mb), _p(cam4), _p(t_cw),
                                          _p(t_lw), _p(last_keys), len(last_keys), _p(hp), _p(ol), _p(world), _p(last_desc),
                                          None if pm is None else _p(pm), float(th), int(mono), int(check_orientation), _p(assigned))
    return nm, assigned[:len(keys_un)].copy()


def search_for_triangulation(kps1, desc1, has_point1, u_right1, fv1, kps2, desc2, has_point2, u_right2, fv2, f12, cam4, c2,
                             scale_factors, level_sigma2, only_stereo=False, coarse=False, check_orientation=True):
    """ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040) with Pinhole::EpipolarConstrain's own line test
    (pinhole_model.cc:121-134) on a given F12; c2 = T2w * Cw fixes the epipole.  Returns (nmatches, match_of_1[n1], epipole)."""
    L = _frame()
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.reff_search_for_triangulation.argtypes = [vp, vp, i, vp, vp, vp, vp, i, vp, i, vp, vp, i, vp, vp, vp, vp, i, vp, i,
                                                vp, vp, vp, vp, vp, i, f, i, i, i, vp, vp]
    kps1, kps2 = np.ascontiguousarray(kps1, KP_DTYPE), np.ascontiguousarray(kps2, KP_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    hp1, hp2 = np.ascontiguousarray(has_point1, np.uint8), np.ascontiguousarray(has_point2, np.uint8)
    ur1, ur2 = np.ascontiguousarray(u_right1, np.float32), np.ascontiguousarray(u_right2, np.float32)
    f12, cam4, c2 = (np.ascontiguousarray(a, np.float32).ravel() for a in (f12, cam4, c2))
    sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
    (n1, b1, f1), (n2, b2, f2) = fv1, fv2
    f1p = f1 if len(f1) else np.zeros(1, np.uint32)
    f2p = f2 if len(f2) else np.zeros(1, np.uint32)
    out = np.empty(max(len(kps1), 1), np.int32)
    ep = np.zeros(2, np.float32)
    nm = L.reff_search_for_triangulation(_p(kps1), _p(desc1), len(kps1), _p(hp1), _p(ur1), _p(n1), _p(b1), len(n1), _p(f1p), len(f1),
                                         _p(kps2), _p(desc2), len(kps2), _p(hp2), _p(ur2), _p(n2), _p(b2), len(n2), _p(f2p), len(f2),
                                         _p(f12), _p(cam4), _p(c2), _p(sf), _p(s2), len(sf), 0.6, int(only_stereo), int(coarse),
                                         int(check_orientation), _p(ep), _p(out))
    return nm, out[:len(kps1)].copy(), ep


def fuse_search(keys_un, desc, bounds, inv_level_sigma2, pt_u, pt_v, pt_ur, pt_radius, pt_level, pt_desc, u_right=None):
    """The candidate loop of ORBmatcher::Fuse (orb_matcher.cc:1145-1190) after KeyFrame::GetFeaturesInArea, per projected
    map point: (best_idx, best_dist)."""
    L = _frame()
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.reff_fuse_search.argtypes = [vp, vp, i, vp, f, f, f, f, vp, i, vp, vp, vp, vp, vp, vp, i, vp, vp]
    keys_un = np.ascontiguousarray(keys_un, KP_DTYPE)
    desc, pt_desc = np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(pt_desc, np.uint8)
    inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
    pu, pv, pr, prad = (np.ascontiguousarray(a, np.float32) for a in (pt_u, pt_v, pt_ur, pt_radius))
    plev = np.ascontiguousarray(pt_level, np.int32)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    bi, bd = np.empty(len(pu), np.int32), np.empty(len(pu), np.int32)
    L.reff_fuse_search(_p(keys_un), _p(desc), len(keys_un), None if ur is None else _p(ur), *[float(b) for b in bounds], _p(inv), len(inv),
                       _p(pu), _p(pv), _p(pr), _p(prad), _p(plev), _p(pt_desc), len(pu), _p(bi), _p(bd))
    return bi, bd
