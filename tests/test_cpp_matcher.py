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
