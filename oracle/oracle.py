"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm import
this module.  The product package (orb_slam_fusion_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

TRIG_LIBM, TRIG_CR = 0, 1


class Params(C.Structure):
    _fields_ = [("num_feats", C.c_int), ("scale_factor", C.c_float), ("num_levs", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int)]


class GridGeom(C.Structure):
    _fields_ = [("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float),
                ("inv_h", C.c_float), ("cols", C.c_int), ("rows", C.c_int)]


WQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("min_level", "<i4"),
                     ("max_level", "<i4")])
WR_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("best_level", "<i4"),
                     ("best_dist2", "<i4"), ("best_level2", "<i4")])


def build(force=False):
    so = os.path.join(_HERE, "liborb_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cvprim.c", "orb_oracle.c", "cvprim.h", "orb_oracle.h", "bow_oracle.c", "bow_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i, f, sz, u64, i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_uint64, C.c_int64
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(Params)]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_trig.argtypes = [vp, i]
        L.orc_tables.argtypes = [vp] + [vp] * 6
        L.orc_compute_pyramid.argtypes = [vp, vp, i, i, sz]
        L.orc_level.restype = vp
        L.orc_level.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(sz)]
        L.orc_blurred_level.restype = vp
        L.orc_blurred_level.argtypes = [vp, i, C.POINTER(sz)]
        L.orc_extract.argtypes = [vp, vp, i, i, sz, i, i, vp, vp, i, C.POINTER(i), C.POINTER(i)]
        L.orc_candidates.argtypes = [vp, i, vp, i]
        L.orc_selected.argtypes = [vp, i, vp, i]
        L.orc_fast_grid.argtypes = [vp, i, i, sz, i, i, vp, i]
        L.orc_octree.argtypes = [vp, i, i, i, i, i, i, vp, i]
        L.orc_ic_angle.restype = f
        L.orc_ic_angle.argtypes = [vp, sz, i, i]
        L.orc_rbrief.argtypes = [vp, sz, i, i, f, i, vp]
        L.orc_hamming.argtypes = [vp, vp]
        L.orc_knn2.argtypes = [vp, i, vp, i64, vp, vp, i]
        L.orc_ratio_accept.argtypes = [i, i, i, C.c_double]
        L.orc_stereo_rowband.argtypes = [vp, vp, i, vp, vp, i, vp, i, f, f, vp, vp]
        L.orc_distinctive.argtypes = [vp, vp, i, vp, vp]
        L.orc_stereo_refine.argtypes = [vp, vp, i, vp, i, vp, vp, vp, vp, vp, i, f, f, f, vp, vp, vp]
        L.orc_window_search.argtypes = [vp, vp, i, C.POINTER(GridGeom), vp, vp, i, vp, vp]
        L.orc_window_search_stereo.argtypes = [vp, vp, i, C.POINTER(GridGeom), vp, vp, i, vp, vp, vp, vp, vp]
        L.orc_window_search_fuse.argtypes = [vp, vp, i, C.POINTER(GridGeom), vp, vp, i, vp, vp, vp, vp]
        L.orc_search_by_projection.argtypes = [vp, vp, i, C.POINTER(GridGeom), vp, vp, i, vp, vp, vp, vp, i, f, vp]
        L.orc_search_by_projection_last.argtypes = [vp, vp, i, C.POINTER(GridGeom), vp, vp, vp, i, vp, vp, vp, vp, i, i, vp]
        L.orc_search_by_bow.argtypes = [vp, vp, vp, vp, vp, i, vp, i, vp, vp, i, vp, vp, i, vp, i, f, i, vp]
        L.orc_search_by_bow_kf.argtypes = [vp, vp, vp, i, vp, vp, i, vp, i, vp, vp, vp, i, vp, vp, i, vp, i, f, i, vp]
        L.orc_search_for_triangulation.argtypes = [vp, vp, vp, vp, i, vp, vp, i, vp, i, vp, vp, vp, vp, i, vp, vp, i, vp, i,
                                                   vp, vp, vp, vp, i, i, i, vp]
        L.orc_splitmix64.restype = u64
        L.orc_splitmix64.argtypes = [u64]
        L.orc_synth_blocks_v1.argtypes = [vp, i, i, sz, u64, u64, i, u64]
        L.orc_synth_uniform_v1.argtypes = [vp, i, i, sz, u64, u64]
        L.orc_synth_descriptors.argtypes = [vp, i64, i64, u64]
        for name, rt, at in [
            ("cvp_resize_linear_u8", None, [vp, i, i, sz, vp, i, i, sz]),
            ("cvp_border_reflect101_u8", None, [vp, i, i, sz, vp, sz, i]),
            ("cvp_fast9_nms_u8", i, [vp, i, i, sz, i, vp, i]),
            ("cvp_gauss7x7_u8", None, [vp, i, i, sz, vp, sz]),
            ("cvp_fast_atan2", f, [f, f]),
        ]:
            fn = getattr(L, name)
            fn.restype = rt
            fn.argtypes = at
        L.orc_vocab_create.restype = vp
        L.orc_vocab_create.argtypes = [i, i, i, i, i, vp, vp, vp, vp]
        L.orc_vocab_load_text.restype = vp
        L.orc_vocab_load_text.argtypes = [C.c_char_p]
        L.orc_vocab_destroy.argtypes = [vp]
        L.orc_vocab_nodes.argtypes = [vp]
        L.orc_vocab_words.argtypes = [vp]
        L.orc_vocab_arrays.argtypes = [vp] + [C.POINTER(i)] * 4 + [vp] * 4
        L.orc_bow_features.argtypes = [vp, vp, i, i, vp, vp, vp]
        L.orc_bow_transform.argtypes = [vp, vp, i, i, vp, vp, C.POINTER(i), vp, vp, C.POINTER(i), vp, C.POINTER(i)]
        L.orc_synth_vocab.argtypes = [i, i, u64, vp, vp, vp, vp]
        L.orc_vocab_save_text.argtypes = [C.c_char_p, i, i, i, i, i, vp, vp, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8img(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ---------------------------------------------------------------- primitives
def resize_linear(img, dw, dh):
    img = _u8img(img)
    out = np.empty((dh, dw), np.uint8)
    lib().cvp_resize_linear_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw)
    return out


def border_reflect101(img, b):
    img = _u8img(img)
    h, w = img.shape
    out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
    lib().cvp_border_reflect101_u8(_p(img), w, h, img.strides[0], _p(out), w + 2 * b, b)
    return out


def fast9_nms(img, threshold):
    img = _u8img(img)
    h, w = img.shape
    cap = max(16, w * h // 2)
    out = np.empty((cap, 3), np.int32)
    n = lib().cvp_fast9_nms_u8(_p(img), w, h, img.strides[0], threshold, _p(out), cap)
    return out[:n].copy()


def gauss7x7(img):
    img = _u8img(img)
    h, w = img.shape
    out = np.empty((h, w), np.uint8)
    lib().cvp_gauss7x7_u8(_p(img), w, h, img.strides[0], _p(out), w)
    return out


def fast_atan2(y, x):
    return lib().cvp_fast_atan2(float(y), float(x))


# ---------------------------------------------------------------- stages
def fast_grid(level, ini_th=20, min_th=7):
    level = _u8img(level)
    h, w = level.shape
    cap = max(64, w * h // 4)
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast_grid(_p(level), w, h, level.strides[0], ini_th, min_th, _p(out), cap)
    return out[:n].copy()


def octree(xyr, w, h, quota):
    """DistributeOctTree on candidates (x, y, response) relative to (16,16) of a w x h level.
    Returns indices into xyr in the reference's output order."""
    xyr = np.ascontiguousarray(xyr, np.int32).reshape(-1, 3)
    cap = quota + 8 + len(xyr)
    out = np.empty(cap, np.int32)
    n = lib().orc_octree(_p(xyr), len(xyr), 16, w - 16, 16, h - 16, quota, _p(out), cap)
    if n < 0:
        raise ValueError("octree: degenerate geometry")
    return out[:n].copy()


def ic_angle(level, cx, cy):
    level = _u8img(level)
    return lib().orc_ic_angle(_p(level), level.strides[0], int(cx), int(cy))


def rbrief(blurred, cx, cy, angle_deg, trig=TRIG_LIBM):
    blurred = _u8img(blurred)
    d = np.empty(32, np.uint8)
    lib().orc_rbrief(_p(blurred), blurred.strides[0], int(cx), int(cy), float(angle_deg), trig, _p(d))
    return d


class Extractor:
    """Oracle counterpart of ORB_SLAM_FUSION::OrbExtractor (orb_extractor.h:44-104)."""

    def __init__(self, num_feats=1000, scale_factor=1.2, num_levs=8, ini_th_fast=20, min_th_fast=7,
                 trig=TRIG_LIBM):
        self.params = Params(num_feats, scale_factor, num_levs, ini_th_fast, min_th_fast)
        self.h = lib().orc_create(C.byref(self.params))
        if not self.h:
            raise ValueError("bad parameters")
        self.num_levs = num_levs
        lib().orc_set_trig(self.h, trig)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def tables(self):
        L = self.num_levs
        sc, isc, s2, is2 = (np.empty(L, np.float32) for _ in range(4))
        quota = np.empty(L, np.int32)
        umax = np.empty(16, np.int32)
        lib().orc_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(quota), _p(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, quota=quota, umax=umax)

    def compute_pyramid(self, img):
        img = _u8img(img)
        rc = lib().orc_compute_pyramid(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0])
        if rc:
            raise ValueError("pyramid rc=%d" % rc)
        return [self.level(l) for l in range(self.num_levs)]

    def level(self, lev, with_border=False):
        w, h, st = C.c_int(), C.c_int(), C.c_size_t()
        p = lib().orc_level(self.h, lev, C.byref(w), C.byref(h), C.byref(st))
        b = 19
        buf = (C.c_uint8 * (st.value * (h.value + 2 * b))).from_address(p - b * st.value - b)
        a = np.frombuffer(buf, np.uint8).reshape(h.value + 2 * b, st.value).copy()
        return a if with_border else a[b:b + h.value, b:b + w.value].copy()

    def blurred(self, lev):
        st = C.c_size_t()
        w, h = C.c_int(), C.c_int()
        lib().orc_level(self.h, lev, C.byref(w), C.byref(h), None)
        p = lib().orc_blurred_level(self.h, lev, C.byref(st))
        buf = (C.c_uint8 * (w.value * h.value)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(h.value, w.value).copy()

    def __call__(self, img, lapping=(0, 0)):
        """Returns (n_mono, kps[KP_DTYPE], desc[N,32])."""
        if img is None or img.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        img = _u8img(img)
        cap = self.params.num_feats * 2 + 64 * self.num_levs
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n, nm = C.c_int(), C.c_int()
        rc = lib().orc_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                               int(lapping[0]), int(lapping[1]), _p(kps), _p(desc), cap,
                               C.byref(n), C.byref(nm))
        if rc:
            raise RuntimeError("orc_extract rc=%d" % rc)
        return nm.value, kps[:n.value].copy(), desc[:n.value].copy()

    def candidates(self, lev):
        n = lib().orc_candidates(self.h, lev, None, 0)
        out = np.empty((max(n, 1), 3), np.int32)
        lib().orc_candidates(self.h, lev, _p(out), n)
        return out[:n]

    def selected(self, lev):
        n = lib().orc_selected(self.h, lev, None, 0)
        out = np.empty(max(n, 1), KP_DTYPE)
        lib().orc_selected(self.h, lev, _p(out), n)
        return out[:n]


# ---------------------------------------------------------------- matching
def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming(_p(a), _p(b))


def knn2(q, d, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    d = np.ascontiguousarray(d, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), 2), np.int64)
    dist = np.empty((len(q), 2), np.int32)
    lib().orc_knn2(_p(q), len(q), _p(d), len(d), _p(idx), _p(dist), nthreads)
    return idx, dist


def ratio_accept(idx, dist, ratio=0.7):
    return np.array([bool(lib().orc_ratio_accept(int(d[0]), int(d[1]), int(i[1] >= 0), ratio))
                     for i, d in zip(idx, dist)], bool)


def stereo_rowband(kl, dl, kr, dr, scale_factors, n_rows, min_d, max_d):
    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8)
    dr = np.ascontiguousarray(dr, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    bi = np.empty(len(kl), np.int32)
    bd = np.empty(len(kl), np.int32)
    lib().orc_stereo_rowband(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(sf), int(n_rows),
                             float(min_d), float(max_d), _p(bi), _p(bd))
    return bi, bd


def distinctive(desc, offsets):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    n = len(offsets) - 1
    bi = np.empty(n, np.int32)
    bm = np.empty(n, np.int32)
    lib().orc_distinctive(_p(desc), _p(offsets), n, _p(bi), _p(bm))
    return bi, bm


class LevelView(C.Structure):
    _fields_ = [("px", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("stride", C.c_size_t)]


def stereo_refine(levels_left, levels_right, kl, kr, best_idx, best_dist, scale_factors, inv_scale_factors,
                  th_orb_dist, min_d, max_d, bf):
    """levels_*: lists of level images WITH their 19-px border (Extractor.level(l, with_border=True)).
    Returns (u_right, depth, sad)."""
    nlev = len(levels_left)
    keep = []

    def views(levels):
        arr = (LevelView * nlev)()
        for l, im in enumerate(levels):
            im = np.ascontiguousarray(im, np.uint8)
            keep.append(im)
            arr[l] = LevelView(im.ctypes.data + 19 * im.strides[0] + 19, im.shape[1] - 38, im.shape[0] - 38, im.strides[0])
        return arr

    vl, vr = views(levels_left), views(levels_right)
    kl = np.ascontiguousarray(kl, KP_DTYPE)
    kr = np.ascontiguousarray(kr, KP_DTYPE)
    bi = np.ascontiguousarray(best_idx, np.int32)
    bd = np.ascontiguousarray(best_dist, np.int32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    isf = np.ascontiguousarray(inv_scale_factors, np.float32)
    ur = np.empty(len(kl), np.float32)
    dp = np.empty(len(kl), np.float32)
    sad = np.empty(len(kl), np.int32)
    lib().orc_stereo_refine(C.cast(vl, C.c_void_p), C.cast(vr, C.c_void_p), nlev, _p(kl), len(kl), _p(kr), _p(bi), _p(bd),
                            _p(sf), _p(isf), int(th_orb_dist), float(min_d), float(max_d), float(bf), _p(ur), _p(dp), _p(sad))
    return ur, dp, sad


def window_search(kps, desc, geom, queries, qdesc, skip=None, kp_u_right=None, q_u_right=None, q_max_err=None):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    out = np.empty(len(queries), WR_DTYPE)
    g = GridGeom(*geom)
    if kp_u_right is None:
        lib().orc_window_search(_p(kps), _p(desc), len(kps), C.byref(g), _p(queries), _p(qdesc),
                                len(queries), None if sk is None else _p(sk), _p(out))
    else:
        ur = np.ascontiguousarray(kp_u_right, np.float32)
        qr = np.ascontiguousarray(q_u_right, np.float32)
        qe = np.ascontiguousarray(q_max_err, np.float32)
        lib().orc_window_search_stereo(_p(kps), _p(desc), len(kps), C.byref(g), _p(queries), _p(qdesc), len(queries),
                                       None if sk is None else _p(sk), _p(ur), _p(qr), _p(qe), _p(out))
    return out


def window_search_fuse(kps, desc, geom, queries, qdesc, inv_level_sigma2, kp_u_right=None, q_u_right=None):
    """The search of ORBmatcher::Fuse (orb_matcher.cc:1130-1187)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
    ur = qr = None
    if kp_u_right is not None:
        ur, qr = np.ascontiguousarray(kp_u_right, np.float32), np.ascontiguousarray(q_u_right, np.float32)
    out = np.empty(len(queries), WR_DTYPE)
    g = GridGeom(*geom)
    lib().orc_window_search_fuse(_p(kps), _p(desc), len(kps), C.byref(g), _p(queries), _p(qdesc), len(queries),
                                 None if ur is None else _p(ur), None if qr is None else _p(qr), _p(inv), _p(out))
    return out


def search_by_projection(kps, desc, geom, queries, qdesc, skip=None, kp_u_right=None, q_u_right=None, q_max_err=None,
                         th_high=100, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, ...) (orb_matcher.cc:42-134) over prepared windows:
    (nmatches, assigned[n])."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    ur = qr = qe = None
    if kp_u_right is not None:
        ur, qr, qe = (np.ascontiguousarray(a, np.float32) for a in (kp_u_right, q_u_right, q_max_err))
    out = np.empty(max(len(kps), 1), np.int32)
    g = GridGeom(*geom)
    nm = lib().orc_search_by_projection(_p(kps), _p(desc), len(kps), C.byref(g), _p(queries), _p(qdesc), len(queries),
                                        None if sk is None else _p(sk), None if ur is None else _p(ur),
                                        None if qr is None else _p(qr), None if qe is None else _p(qe), int(th_high),
                                        float(nnratio), _p(out))
    return nm, out[:len(kps)].copy()


def search_by_projection_last(kps, desc, geom, queries, qdesc, q_angle, skip=None, kp_u_right=None, q_u_right=None,
                              q_max_err=None, th_high=100, check_orientation=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (orb_matcher.cc:1518-1728) after the projection:
    (nmatches, assigned[n])."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    qa = np.ascontiguousarray(q_angle, np.float32)
    sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    ur = qr = qe = None
    if kp_u_right is not None:
        ur, qr, qe = (np.ascontiguousarray(a, np.float32) for a in (kp_u_right, q_u_right, q_max_err))
    out = np.empty(max(len(kps), 1), np.int32)
    g = GridGeom(*geom)
    nm = lib().orc_search_by_projection_last(_p(kps), _p(desc), len(kps), C.byref(g), _p(queries), _p(qdesc), _p(qa), len(queries),
                                             None if sk is None else _p(sk), None if ur is None else _p(ur),
                                             None if qr is None else _p(qr), None if qe is None else _p(qe), int(th_high),
                                             int(check_orientation), _p(out))
    return nm, out[:len(kps)].copy()


def pack_feature_vector(nodes, groups):
    """(node ids, [feature index arrays per node]) -> (nodes u32, begin i32, feats u32): the flat FeatureVector layout."""
    nodes = np.ascontiguousarray(nodes, np.uint32)
    begin = np.zeros(max(len(nodes), 1), np.int32)
    if len(nodes):
        begin[:len(nodes)] = np.concatenate([[0], np.cumsum([len(g) for g in groups])[:-1]])
    feats = np.ascontiguousarray(np.concatenate([np.asarray(g, np.uint32) for g in groups]) if len(groups) else np.zeros(0, np.uint32), np.uint32)
    return nodes, begin, feats


def search_by_bow(kps_kf, desc_kf, has_point_kf, fv_kf, kps_f, desc_f, fv_f, nnratio=0.7, check_orientation=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (orb_matcher.cc:215-389, Nleft == -1).  fv_* = (nodes, begin, feats).
    Returns (nmatches, match_of_f[n_f])."""
    kps_kf, kps_f = np.ascontiguousarray(kps_kf, KP_DTYPE), np.ascontiguousarray(kps_f, KP_DTYPE)
    desc_kf, desc_f = np.ascontiguousarray(desc_kf, np.uint8), np.ascontiguousarray(desc_f, np.uint8)
    hp = None if has_point_kf is None else np.ascontiguousarray(has_point_kf, np.uint8)
    nk, bk, fk = fv_kf
    nf, bf_, ff = fv_f
    fk = fk if len(fk) else np.zeros(1, np.uint32)
    ff = ff if len(ff) else np.zeros(1, np.uint32)
    out = np.empty(max(len(kps_f), 1), np.int32)
    nm = lib().orc_search_by_bow(_p(kps_kf), _p(desc_kf), None if hp is None else _p(hp), _p(nk), _p(bk), len(nk), _p(fk),
                                 len(fv_kf[2]), _p(kps_f), _p(desc_f), len(kps_f), _p(nf), _p(bf_), len(nf), _p(ff),
                                 len(fv_f[2]), float(nnratio), int(check_orientation), _p(out))
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
    nm = lib().orc_search_by_bow_kf(_p(kps1), _p(desc1), None if hp1 is None else _p(hp1), len(kps1), _p(n1), _p(b1), len(n1), _p(f1p),
                                    len(f1), _p(kps2), _p(desc2), None if hp2 is None else _p(hp2), len(kps2), _p(n2), _p(b2), len(n2),
                                    _p(f2p), len(f2), float(nnratio), int(check_orientation), _p(out))
    return nm, out[:len(kps1)].copy()


def search_for_triangulation(kps1, desc1, has_point1, u_right1, fv1, kps2, desc2, has_point2, u_right2, fv2, f12, ep,
                             scale_factors, level_sigma2, only_stereo=False, coarse=False, check_orientation=True):
    """ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040), one pinhole camera per key frame: (nmatches, match_of_1[n1])."""
    kps1, kps2 = np.ascontiguousarray(kps1, KP_DTYPE), np.ascontiguousarray(kps2, KP_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    hp1, hp2 = np.ascontiguousarray(has_point1, np.uint8), np.ascontiguousarray(has_point2, np.uint8)
    ur1, ur2 = np.ascontiguousarray(u_right1, np.float32), np.ascontiguousarray(u_right2, np.float32)
    f12, ep = np.ascontiguousarray(f12, np.float32).reshape(9), np.ascontiguousarray(ep, np.float32)
    sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
    (n1, b1, f1), (n2, b2, f2) = fv1, fv2
    f1p = f1 if len(f1) else np.zeros(1, np.uint32)
    f2p = f2 if len(f2) else np.zeros(1, np.uint32)
    out = np.empty(max(len(kps1), 1), np.int32)
    nm = lib().orc_search_for_triangulation(_p(kps1), _p(desc1), _p(hp1), _p(ur1), len(kps1), _p(n1), _p(b1), len(n1), _p(f1p), len(f1),
                                            _p(kps2), _p(desc2), _p(hp2), _p(ur2), len(kps2), _p(n2), _p(b2), len(n2), _p(f2p), len(f2),
                                            _p(f12), _p(ep), _p(sf), _p(s2), int(only_stereo), int(coarse), int(check_orientation), _p(out))
    return nm, out[:len(kps1)].copy()


# ---------------------------------------------------------------- synthetic inputs
def splitmix64(x):
    return lib().orc_splitmix64(x & 0xFFFFFFFFFFFFFFFF)


def blocks_v1(w, h, seed=1, frame=0, shift_x=0, noise_seed=None):
    img = np.empty((h, w), np.uint8)
    lib().orc_synth_blocks_v1(_p(img), w, h, w, seed, frame, shift_x,
                              seed if noise_seed is None else noise_seed)
    return img


def uniform_v1(w, h, seed=1, frame=0):
    img = np.empty((h, w), np.uint8)
    lib().orc_synth_uniform_v1(_p(img), w, h, w, seed, frame)
    return img


def synth_descriptors(first, n, seed):
    out = np.empty((n, 32), np.uint8)
    lib().orc_synth_descriptors(_p(out), first, n, seed)
    return out


# ---------------------------------------------------------------- bag of words (bow_oracle.c)
def synth_vocab(k, L, seed=7):
    """(parent, is_leaf, desc, weight) of the deterministic synthetic vocabulary tree, node-id indexed."""
    n = lib().orc_synth_vocab(k, L, seed, None, None, None, None)
    parent = np.zeros(n, np.int32)
    leaf = np.zeros(n, np.uint8)
    desc = np.zeros((n, 32), np.uint8)
    weight = np.zeros(n, np.float64)
    lib().orc_synth_vocab(k, L, seed, _p(parent), _p(leaf), _p(desc), _p(weight))
    return parent, leaf, desc, weight


def save_vocab_text(path, k, L, parent, leaf, desc, weight, scoring=0, weighting=0):
    rc = lib().orc_vocab_save_text(os.fsencode(path), k, L, scoring, weighting, len(parent), _p(parent), _p(leaf),
                                   _p(desc), _p(weight))
    assert rc == 0


def unpack_bow(n, ids, vals, nb, nodes, begin, nf, feats, total):
    """Common result form: (word ids, values, node ids, [feature index arrays per node])."""
    ends = list(begin[1:nf]) + [total]
    return (ids[:nb].copy(), vals[:nb].copy(), nodes[:nf].copy(),
            [feats[begin[j]:ends[j]].copy() for j in range(nf)])


class Vocabulary:
    """CPU restatement of DBoW2's TemplatedVocabulary<FORB> (load + transform)."""

    def __init__(self, k=None, L=None, parent=None, leaf=None, desc=None, weight=None, scoring=0, weighting=0,
                 path=None):
        if path is not None:
            self.h = lib().orc_vocab_load_text(os.fsencode(path))
        else:
            self.h = lib().orc_vocab_create(k, L, scoring, weighting, len(parent), _p(parent), _p(leaf), _p(desc),
                                            _p(weight))
        if not self.h:
            raise ValueError("vocabulary could not be created / loaded")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_vocab_destroy(self.h)
            self.h = None

    @property
    def n_nodes(self):
        return lib().orc_vocab_nodes(self.h)

    @property
    def n_words(self):
        return lib().orc_vocab_words(self.h)

    def arrays(self):
        n = self.n_nodes
        k, L, sc, we = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        parent, leaf = np.zeros(n, np.int32), np.zeros(n, np.uint8)
        desc, weight = np.zeros((n, 32), np.uint8), np.zeros(n, np.float64)
        lib().orc_vocab_arrays(self.h, C.byref(k), C.byref(L), C.byref(sc), C.byref(we), _p(parent), _p(leaf), _p(desc),
                               _p(weight))
        return k.value, L.value, sc.value, we.value, parent, leaf, desc, weight

    def features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid, w, nid = np.zeros(n, np.uint32), np.zeros(n, np.float64), np.zeros(n, np.uint32)
        lib().orc_bow_features(self.h, _p(desc), n, levelsup, _p(wid), _p(w), _p(nid))
        return wid, w, nid

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        m = max(n, 1)
        ids, vals = np.zeros(m, np.uint32), np.zeros(m, np.float64)
        nodes, begin, feats = np.zeros(m, np.uint32), np.zeros(m, np.int32), np.zeros(m, np.uint32)
        nb, nf, tot = C.c_int(), C.c_int(), C.c_int()
        lib().orc_bow_transform(self.h, _p(desc), n, levelsup, _p(ids), _p(vals), C.byref(nb), _p(nodes), _p(begin),
                                C.byref(nf), _p(feats), C.byref(tot))
        return unpack_bow(n, ids, vals, nb.value, nodes, begin, nf.value, feats, tot.value)
