"""The ORBmatcher drop-in CLASS (orb_slam_fusion_b200/cpp/include/cam/orb_feature/orb_matcher.h, cpp/src/orb_matcher.cc):
the reference's exact class interface (orb_matcher.h:36-129) whose methods gather Frame / KeyFrame / MapPoint fields,
run the search on the GPU through the C ABI and scatter pointers back.

CPU: the class compiles against the reference-shaped stand-in headers and binds only to the C ABI.
GPU: tests/cpp/libmatcher_facade.so exposes the class behind the same extern "C" entry points as
oracle/_ref/libframe_ref.so (the reference's OWN method bodies, spliced); the same Python wrappers (oracle/ref.py) drive
both with identical arrays, and what the two ORBmatcher classes did to the frames must be identical -- and equal to the
oracle's restatement."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

from test_oracle_vs_ref_frame import (BF_LAST, CAM4, MB_LAST, H, W, _bow_pair, _frame_and_points, _greedy_with_oracle,
                                      last_frame_case, last_frame_windows, triangulation_case)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "cpp", "libmatcher_facade.so")


def build():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp"), "libmatcher_facade.so"], stdout=subprocess.DEVNULL)
    return SO


def test_matcher_class_compiles_against_the_reference_shaped_headers_and_binds_only_to_the_c_abi():
    so = build()
    syms = subprocess.run(["nm", "-D", "-C", so], capture_output=True, text=True).stdout
    undefined = {l.split()[-1] for l in syms.splitlines() if " U " in l}
    used = {s for s in undefined if s.startswith(("orbm_", "orbx_", "orbv_"))}
    assert {"orbm_search_by_projection", "orbm_search_by_projection_last", "orbm_search_by_bow", "orbm_search_by_bow_kf",
            "orbm_search_for_triangulation", "orbm_window_search", "orbm_window_search_fuse"} <= used, used
    # every method of the reference's class (orb_matcher.h:36-129) is defined by the drop-in
    defined = "\n".join(l for l in syms.splitlines() if " T " in l)
    for m in ["ORBmatcher::ORBmatcher(float, bool)", "ORBmatcher::DescriptorDistance(cv::Mat const&, cv::Mat const&)",
              "ORBmatcher::SearchByProjection(ORB_SLAM_FUSION::Frame&, std::vector<ORB_SLAM_FUSION::MapPoint*",
              "ORBmatcher::SearchByProjection(ORB_SLAM_FUSION::Frame&, ORB_SLAM_FUSION::Frame const&, float, bool)",
              "ORBmatcher::SearchByProjection(ORB_SLAM_FUSION::Frame&, ORB_SLAM_FUSION::KeyFrame*, std::set<",
              "ORBmatcher::SearchByProjection(ORB_SLAM_FUSION::KeyFrame*, Sophus::Sim3<float>&",
              "ORBmatcher::SearchByBoW(ORB_SLAM_FUSION::KeyFrame*, ORB_SLAM_FUSION::Frame&",
              "ORBmatcher::SearchByBoW(ORB_SLAM_FUSION::KeyFrame*, ORB_SLAM_FUSION::KeyFrame*",
              "ORBmatcher::SearchForInitialization(", "ORBmatcher::SearchForTriangulation(", "ORBmatcher::SearchBySim3(",
              "ORBmatcher::Fuse(ORB_SLAM_FUSION::KeyFrame*, std::vector<", "ORBmatcher::Fuse(ORB_SLAM_FUSION::KeyFrame*, Sophus::Sim3<float>&",
              "ORBmatcher::RadiusByViewingCos(", "ORBmatcher::ComputeThreeMaxima("]:
        assert m in defined, m


@pytest.fixture(scope="module")
def F():
    """oracle/ref.py bound to the facade library instead of the reference's lines."""
    build()
    spec = importlib.util.spec_from_file_location("oracle.ref_facade", os.path.join(ROOT, "oracle", "ref.py"),
                                                  submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "oracle"
    spec.loader.exec_module(mod)
    mod._FRAME_SO = SO
    mod.frame_available = lambda: True
    return mod


@pytest.fixture(scope="module")
def R():
    from oracle import ref
    return ref


@pytest.mark.gpu
def test_static_descriptor_distance(oracle):
    build()
    L = C.CDLL(SO)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    out = np.empty(500, np.int32)
    L.facade_descriptor_distance(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), 500, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, np.array([oracle.hamming(x, y) for x, y in zip(a, b)], np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th,nnratio,stereo,far", [(1, 3.0, 0.8, False, False), (2, 1.0, 0.8, False, True),
                                                        (3, 5.0, 0.9, True, False), (4, 15.0, 0.6, True, True)])
def test_class_search_by_projection(oracle, F, R, seed, th, nnratio, stereo, far):
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints) (orb_matcher.cc:42-134)."""
    kps, desc, pts, qdesc, src, rng = _frame_and_points(oracle, seed, nq=900)
    bounds = (0.0, float(W), 0.0, float(H))
    sf = oracle.Extractor(1000).tables()["scale"]
    pre = (rng.random(len(kps)) < 0.2).astype(np.uint8)
    u_right = None
    if stereo:
        u_right = np.where(rng.random(len(kps)) < 0.6, kps["x"] - rng.uniform(2, 60, len(kps)), -1.0).astype(np.float32)
        pts["proj_xr"] = pts["proj_x"] - rng.uniform(2, 60, len(pts)).astype(np.float32)
        ok = u_right[src] > 0
        pts["proj_xr"][ok] = u_right[src][ok] + rng.normal(0, 2, ok.sum()).astype(np.float32)
    nm, got = F.search_by_projection(kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, 40.0)
    onm, owant = _greedy_with_oracle(oracle, kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, 40.0)
    assert nm == onm and np.array_equal(got, owant) and nm > 150
    if R.frame_available():
        rnm, rwant = R.search_by_projection(kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, 40.0)
        assert nm == rnm and np.array_equal(got, rwant)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th,t_lw_z,mono,stereo,ori", [(1, 7.0, 0.0, False, True, True), (2, 15.0, 0.5, False, True, True),
                                                            (3, 15.0, -0.6, False, False, True), (4, 7.0, 0.5, True, False, False)])
def test_class_search_by_projection_last_frame(oracle, F, R, seed, th, t_lw_z, mono, stereo, ori):
    """ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (orb_matcher.cc:1518-1728), the
    projection of :1539-1566 included (host geometry inside the class)."""
    kps, desc, last, ldesc, has_point, outlier, world, t_cw, src, rng = last_frame_case(oracle, seed, n_last=900)
    bounds = (0.0, float(W), 0.0, float(H))
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    sf = oracle.Extractor(1000).tables()["scale"]
    t_lw = np.array([0.0, 0.0, t_lw_z], np.float32)
    pre = (rng.random(len(kps)) < 0.15).astype(np.uint8)
    u_right = np.where(rng.random(len(kps)) < 0.6, kps["x"] - rng.uniform(2, 40, len(kps)), -1.0).astype(np.float32) if stereo else None
    args = (kps, desc, bounds, sf, BF_LAST, MB_LAST, CAM4, t_cw, t_lw, last, has_point, outlier, world, ldesc, pre, u_right, th, mono, ori)
    nm, got = F.search_by_projection_last(*args)
    keep, q, qur = last_frame_windows(oracle, sf, last, has_point, outlier, world, t_cw, t_lw, th, mono, bounds)
    onm, ogot = oracle.search_by_projection_last(kps, desc, geom, q, ldesc[keep], last["angle"][keep], pre, u_right,
                                                 qur if stereo else None, q["r"] if stereo else None, 100, ori)
    assert nm == onm and np.array_equal(got, np.where(ogot >= 0, keep[np.maximum(ogot, 0)], -1)) and nm > 150
    if R.frame_available():
        rnm, rwant = R.search_by_projection_last(*args)
        assert nm == rnm and np.array_equal(got, rwant)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,shift,ratio,ori,levelsup", [(1, 3, 0.7, True, 2), (2, 8, 0.75, True, 3), (3, 0, 0.9, False, 2)])
def test_class_search_by_bow_both_forms(oracle, F, R, seed, shift, ratio, ori, levelsup):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (orb_matcher.cc:215-389) and (KeyFrame*, KeyFrame*, ...) (:697-815)."""
    ka, da, fva, kb, db, fvb = _bow_pair(oracle, seed, shift, levelsup=levelsup)
    rng = np.random.default_rng(seed)
    hp1 = (rng.random(len(ka)) < 0.8).astype(np.uint8)
    hp2 = (rng.random(len(kb)) < 0.8).astype(np.uint8)
    nm, got = F.search_by_bow(ka, da, hp1, fva, kb, db, fvb, ratio, ori)
    onm, owant = oracle.search_by_bow(ka, da, hp1, fva, kb, db, fvb, ratio, ori)
    assert nm == onm and np.array_equal(got, owant) and nm > 30
    nm2, got2 = F.search_by_bow_kf(ka, da, hp1, fva, kb, db, hp2, fvb, ratio, ori)
    onm2, owant2 = oracle.search_by_bow_kf(ka, da, hp1, fva, kb, db, hp2, fvb, ratio, ori)
    assert nm2 == onm2 and np.array_equal(got2, owant2) and nm2 > 20
    if R.frame_available():
        rnm, rwant = R.search_by_bow(ka, da, hp1, fva, kb, db, fvb, ratio, ori)
        assert nm == rnm and np.array_equal(got, rwant)
        rnm2, rwant2 = R.search_by_bow_kf(ka, da, hp1, fva, kb, db, hp2, fvb, ratio, ori)
        assert nm2 == rnm2 and np.array_equal(got2, rwant2)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,shift,only_stereo,coarse,ori,c2", [(1, 12, False, False, True, (0.5, 0.01, 0.05)),
                                                                  (3, 9, True, False, True, (0.02, -0.01, 1.0)),
                                                                  (4, 12, False, True, False, (-0.3, 0.2, 1.0))])
def test_class_search_for_triangulation(oracle, F, R, seed, shift, only_stereo, coarse, ori, c2):
    """ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040)."""
    ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, sf, s2 = triangulation_case(oracle, seed, shift)
    args = (ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, CAM4, c2, sf, s2, only_stereo, coarse, ori)
    nm, got, ep = F.search_for_triangulation(*args)
    onm, owant = oracle.search_for_triangulation(ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, ep, sf, s2, only_stereo, coarse, ori)
    assert nm == onm and np.array_equal(got, owant) and nm > (10 if only_stereo else 25)
    if R.frame_available():
        rnm, rwant, rep = R.search_for_triangulation(*args)
        assert nm == rnm and np.array_equal(got, rwant) and np.array_equal(ep, rep)


# ---------------------------------------------------------------- the remaining methods: oracle/matcher_harness.inc on both sides
class View(C.Structure):
    _fields_ = [("keys_un", C.c_void_p), ("desc", C.c_void_p), ("n", C.c_int), ("u_right", C.c_void_p),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("scale", C.c_void_p), ("inv_sigma2", C.c_void_p), ("n_levels", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("mb", C.c_float),
                ("tx", C.c_float), ("ty", C.c_float), ("tz", C.c_float)]


MP_DTYPE = np.dtype([("wx", "<f4"), ("wy", "<f4"), ("wz", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                     ("min_dist", "<f4"), ("max_dist", "<f4"), ("level", "<i4"), ("bad", "<i4"), ("n_obs", "<i4")])


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_view(oracle, kps, desc, t, u_right=None, keep=None):
    tab = oracle.Extractor(1000).tables()
    keep = keep if keep is not None else []
    sc, inv = np.ascontiguousarray(tab["scale"], np.float32), np.ascontiguousarray(tab["inv_sigma2"], np.float32)
    keep += [kps, desc, sc, inv, u_right]
    return View(_vp(kps), _vp(desc), len(kps), _vp(u_right), 0.0, float(W), 0.0, float(H), _vp(sc), _vp(inv), 8,
                float(CAM4[0]), float(CAM4[1]), float(CAM4[2]), float(CAM4[3]), float(BF_LAST), float(MB_LAST),
                float(t[0]), float(t[1]), float(t[2]))


def points_near_keypoints(kps, desc, src, cam_t, rng, sigma=3.0, flips=14):
    """Map points whose projection through the pinhole camera at translation cam_t (x_c = x_w + cam_t) lands near keypoints
    kps[src]: world position, a normal towards the camera (some turned away), distance range (some too tight), predicted level."""
    n = len(src)
    z = rng.uniform(2.0, 12.0, n).astype(np.float32)
    u = kps["x"][src] + rng.normal(0, sigma, n).astype(np.float32)
    v = kps["y"][src] + rng.normal(0, sigma, n).astype(np.float32)
    xc = np.stack([(u - CAM4[2]) / CAM4[0] * z, (v - CAM4[3]) / CAM4[1] * z, z], 1).astype(np.float32)
    xw = (xc - np.asarray(cam_t, np.float32)).astype(np.float32)
    po = xw + np.asarray(cam_t, np.float32)                      # p3Dw - Ow with Ow = -cam_t
    nrm = po / np.linalg.norm(po, axis=1, keepdims=True)
    nrm[rng.random(n) < 0.1] *= -1.0                              # viewing angle test fails
    mp = np.zeros(n, MP_DTYPE)
    mp["wx"], mp["wy"], mp["wz"] = xw[:, 0], xw[:, 1], xw[:, 2]
    mp["nx"], mp["ny"], mp["nz"] = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    mp["min_dist"], mp["max_dist"] = 0.1, 100.0
    tight = rng.random(n) < 0.05
    mp["max_dist"][tight] = 1.0                                   # outside the scale invariance region
    mp["level"] = np.clip(kps["octave"][src] + rng.integers(-1, 2, n), 0, 7)
    mp["bad"] = rng.random(n) < 0.05
    mp["n_obs"] = rng.integers(1, 6, n)
    d = desc[src].copy()
    fl = rng.integers(0, 256, (n, flips))
    for j in range(flips):
        d[np.arange(n), fl[:, j] // 8] ^= (1 << (fl[:, j] % 8)).astype(np.uint8)
    return mp, np.ascontiguousarray(d)


@pytest.fixture(scope="module")
def libs(R):
    """(facade library, reference library or None) with the shared harness entry points."""
    build()
    fac = C.CDLL(SO)
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libframe_ref.so")) if R.frame_available() else None
    if ref is not None and not hasattr(ref, "reffh_fuse"):
        ref = None
    return fac, ref


def _frame_kps(oracle, seed):
    img = oracle.blocks_v1(W, H, seed, 0)
    _, kps, desc = oracle.Extractor(1000)(img)
    return np.ascontiguousarray(kps), np.ascontiguousarray(desc)


def both(libs, name, make_args, outs):
    """Run entry point `name` in the facade and (when it travelled) the reference library on fresh copies of the in/out arrays;
    return [(ret, outs...) per library]."""
    res = []
    for L in libs:
        if L is None:
            continue
        args, out_arrays = make_args()
        fn = getattr(L, name)
        fn.restype = C.c_int
        ret = fn(*args)
        res.append((ret,) + tuple(a.copy() for a in out_arrays))
    return res


def same(res):
    if len(res) == 2:
        assert res[0][0] == res[1][0], (res[0][0], res[1][0])
        for a, b in zip(res[0][1:], res[1][1:]):
            assert np.array_equal(a, b)
    return res[0]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th,orb_dist,ori", [(1, 10.0, 100, True), (2, 3.0, 64, True), (3, 10.0, 100, False)])
def test_class_search_by_projection_relocalisation(oracle, libs, seed, th, orb_dist, ori):
    """ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist) (orb_matcher.cc:1730-1840)."""
    kps, desc = _frame_kps(oracle, seed)
    rng = np.random.default_rng(seed)
    t = np.array([0.05, -0.02, 0.1], np.float32)
    n_kf = 900
    src = rng.integers(0, len(kps), n_kf)
    mp, mdesc = points_near_keypoints(kps, desc, src, t, rng)
    kf_keys = np.ascontiguousarray(kps[src])                      # the key frame's own keypoints (their angles feed the histogram)
    kf_keys["angle"] = (kf_keys["angle"] + rng.normal(0, 8, n_kf)).astype(np.float32) % np.float32(360)
    kf_point = np.where(rng.random(n_kf) < 0.9, np.arange(n_kf), -1).astype(np.int32)
    found = (rng.random(n_kf) < 0.1).astype(np.uint8)
    taken = (rng.random(len(kps)) < 0.15).astype(np.uint8)
    keep = []
    view = make_view(oracle, kps, desc, t, keep=keep)

    def make_args():
        assigned = np.full(len(kps), -7, np.int32)
        return (C.byref(view), _vp(kf_keys), n_kf, _vp(kf_point), _vp(mp), _vp(mdesc), n_kf, _vp(found), _vp(taken), C.c_float(th),
                orb_dist, int(ori), _vp(assigned)), [assigned]
    nm, assigned = same(both(libs, "reffh_sbp_reloc", make_args, None))
    assert nm > 150 and (assigned >= 0).sum() == nm and (assigned[taken != 0] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th,ratio,with_kfs", [(1, 8, 1.0, False), (2, 4, 1.5, True), (3, 8, 0.8, True)])
def test_class_search_by_projection_sim3(oracle, libs, seed, th, ratio, with_kfs):
    """ORBmatcher::SearchByProjection(KeyFrame*, Sim3, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th, ratioHamming)
    (orb_matcher.cc:391-488, 490-596)."""
    kps, desc = _frame_kps(oracle, seed + 10)
    rng = np.random.default_rng(seed + 10)
    s, st = np.float32(1.25), np.array([0.1, 0.05, -0.2], np.float32)
    cam_t = (st / s).astype(np.float32)                           # Tcw = (R, t / s)
    n_pts = 900
    src = rng.integers(0, len(kps), n_pts)
    mp, mdesc = points_near_keypoints(kps, desc, src, cam_t, rng)
    matched0 = np.where(rng.random(len(kps)) < 0.1, rng.integers(0, n_pts, len(kps)), -1).astype(np.int32)
    keep = []
    view = make_view(oracle, kps, desc, cam_t, keep=keep)

    def make_args():
        matched, matched_kf = matched0.copy(), np.full(len(kps), -7, np.int32)
        return (C.byref(view), C.c_float(s), _vp(st), _vp(mp), _vp(mdesc), n_pts, th, C.c_float(ratio), int(with_kfs), _vp(matched),
                _vp(matched_kf)), [matched, matched_kf]
    nm, matched, matched_kf = same(both(libs, "reffh_sbp_sim3", make_args, None))
    new = (matched >= 0) & (matched0 < 0)
    assert nm > 100 and new.sum() == nm and np.array_equal(matched[matched0 >= 0], matched0[matched0 >= 0])
    if with_kfs:
        assert np.array_equal(matched_kf[new], matched[new])      # point p came with "its key frame" number p


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th", [(1, 7.5), (2, 3.0)])
def test_class_search_by_sim3(oracle, libs, seed, th):
    """ORBmatcher::SearchBySim3 (orb_matcher.cc:1320-1516): both projection loops + the mutual check."""
    kps, desc = _frame_kps(oracle, seed + 20)
    rng = np.random.default_rng(seed + 20)
    t1 = np.array([0.0, 0.0, 0.0], np.float32)
    n = len(kps)
    idx = np.arange(n)
    mp1, d1 = points_near_keypoints(kps, desc, idx, t1, rng, sigma=1.5, flips=8)
    mp2, d2 = points_near_keypoints(kps, desc, idx, t1, rng, sigma=1.5, flips=8)
    point1 = np.where(rng.random(n) < 0.8, idx, -1).astype(np.int32)
    point2 = np.where(rng.random(n) < 0.8, idx, -1).astype(np.int32)
    m0 = np.where(rng.random(n) < 0.05, idx, -1).astype(np.int32)  # a few pairs are matched already
    s, st = np.float32(1.0), np.array([0.001, -0.001, 0.002], np.float32)
    keep = []
    v1, v2 = make_view(oracle, kps, desc, t1, keep=keep), make_view(oracle, kps, desc, t1, keep=keep)

    def make_args():
        m12 = m0.copy()
        return (C.byref(v1), C.byref(v2), _vp(point1), _vp(mp1), _vp(d1), n, _vp(point2), _vp(mp2), _vp(d2), n, C.c_float(s), _vp(st),
                C.c_float(th), _vp(m12)), [m12]
    nf, m12 = same(both(libs, "reffh_search_by_sim3", make_args, None))
    assert nf > 100 and (m12 >= 0).sum() >= nf


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th,stereo", [(1, 3.0, True), (2, 3.0, False), (3, 5.0, True)])
def test_class_fuse(oracle, libs, seed, th, stereo):
    """ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th, bRight) (orb_matcher.cc:1042-1212): search + Replace /
    AddObservation / AddMapPoint bookkeeping, in the reference's order (the log of calls must be identical)."""
    kps, desc = _frame_kps(oracle, seed + 30)
    rng = np.random.default_rng(seed + 30)
    t = np.array([0.02, 0.01, 0.0], np.float32)
    n_pts = 900
    src = rng.integers(0, len(kps), n_pts)
    mp, mdesc = points_near_keypoints(kps, desc, src, t, rng, sigma=2.0, flips=10)
    u_right = np.where(rng.random(len(kps)) < 0.5, kps["x"] - rng.uniform(1, 40, len(kps)), -1.0).astype(np.float32) if stereo else None
    kf_point0 = np.where(rng.random(len(kps)) < 0.4, rng.permutation(n_pts)[:len(kps)] if n_pts >= len(kps) else rng.integers(0, n_pts, len(kps)),
                         -1).astype(np.int32)
    null = (rng.random(n_pts) < 0.03).astype(np.uint8)
    keep = []
    view = make_view(oracle, kps, desc, t, u_right, keep=keep)

    def make_args():
        kfp, bad, log, nlog = kf_point0.copy(), np.zeros(n_pts, np.uint8), np.zeros((4 * n_pts, 3), np.int32), C.c_int(0)
        keep.append(nlog)
        return (C.byref(view), _vp(kfp), _vp(mp), _vp(mdesc), n_pts, _vp(null), C.c_float(th), _vp(bad), _vp(log), len(log),
                C.byref(nlog)), [kfp, bad, log]
    nf, kfp, bad, log = same(both(libs, "reffh_fuse", make_args, None))
    ops = log[(log != 0).any(1)]
    assert nf > 80 and (ops[:, 0] == 2).sum() > 10 and (ops[:, 0] == 1).sum() > 10    # both Replace and AddObservation happen
    assert (kfp != kf_point0).sum() == (ops[:, 0] == 3).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,th", [(1, 4.0), (2, 8.0)])
def test_class_fuse_sim3(oracle, libs, seed, th):
    """ORBmatcher::Fuse(KeyFrame*, Sim3, vpPoints, th, vpReplacePoint) (orb_matcher.cc:1214-1318)."""
    kps, desc = _frame_kps(oracle, seed + 40)
    rng = np.random.default_rng(seed + 40)
    s, st = np.float32(0.9), np.array([-0.05, 0.02, 0.1], np.float32)
    cam_t = (st / s).astype(np.float32)
    n_pts = 800
    src = rng.integers(0, len(kps), n_pts)
    mp, mdesc = points_near_keypoints(kps, desc, src, cam_t, rng, sigma=2.0, flips=10)
    kf_point0 = np.where(rng.random(len(kps)) < 0.4, rng.integers(0, n_pts, len(kps)), -1).astype(np.int32)
    keep = []
    view = make_view(oracle, kps, desc, cam_t, keep=keep)

    def make_args():
        kfp, rep, log, nlog = kf_point0.copy(), np.full(n_pts, -7, np.int32), np.zeros((4 * n_pts, 3), np.int32), C.c_int(0)
        keep.append(nlog)
        return (C.byref(view), C.c_float(s), _vp(st), _vp(kfp), _vp(mp), _vp(mdesc), n_pts, C.c_float(th), _vp(rep), _vp(log), len(log),
                C.byref(nlog)), [kfp, rep, log]
    nf, kfp, rep, log = same(both(libs, "reffh_fuse_sim3", make_args, None))
    assert nf > 60 and (rep >= 0).sum() > 10 and (kfp != kf_point0).sum() > 10


@pytest.mark.gpu
@pytest.mark.parametrize("seed,shift,ratio,ori", [(1, 3, 0.9, True), (2, 10, 0.9, True), (3, 5, 0.7, False)])
def test_class_search_for_initialization(oracle, libs, seed, shift, ratio, ori):
    """ORBmatcher::SearchForInitialization (orb_matcher.cc:597-695; host code in the drop-in, Frame::GetFeaturesInArea is the frame's)."""
    a = oracle.blocks_v1(W, H, seed, 0)
    b = oracle.blocks_v1(W, H, seed, 0, shift_x=shift, noise_seed=seed + 7)
    ex = oracle.Extractor(1000)
    _, ka, da = ex(a)
    _, kb, db = ex(b)
    ka, da, kb, db = (np.ascontiguousarray(x) for x in (ka, da, kb, db))
    t = np.zeros(3, np.float32)
    keep = []
    v1, v2 = make_view(oracle, ka, da, t, keep=keep), make_view(oracle, kb, db, t, keep=keep)
    prev0 = np.ascontiguousarray(np.stack([ka["x"], ka["y"]], 1).astype(np.float32))

    def make_args():
        prev, m12 = prev0.copy(), np.full(len(ka), -7, np.int32)
        return (C.byref(v1), C.byref(v2), _vp(prev), 100, C.c_float(ratio), int(ori), _vp(m12)), [prev, m12]
    nm, prev, m12 = same(both(libs, "reffh_search_for_initialization", make_args, None))
    assert nm > 50 and (m12 >= 0).sum() == nm
    assert (m12[ka["octave"] > 0] == -1).all()                   # only level-0 keypoints take part
