"""CPU: the matcher-side oracle rows against the reference's OWN LINES (oracle/_ref/libframe_ref.so =
frame.cc:828-986, :438-465, :679-759, mappoint.cc:365-433, orb_matcher.cc:35-213 spliced by line range into the
member declarations of oracle/ref_frame_shim.cc and compiled on the mini-cv shim).  This pins
orc_stereo_rowband + orc_stereo_refine (Frame::ComputeStereoMatches), orc_distinctive
(MapPoint::ComputeDistinctiveDescriptors) and orc_window_search(_stereo) (Frame::GetFeaturesInArea + the inner loop of
ORBmatcher::SearchByProjection) -- the oracles the GPU tests of tests/test_gpu_match.py compare against."""
import numpy as np
import pytest

from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.frame_available(), reason="oracle/_ref not built and /root/reference absent")

W, H = 752, 480
BF, MB = np.float32(47.90639384423901), np.float32(0.11)   # settings/EuRoC.yaml: Stereo.b * fx, mb = bf / fx


def _stereo_pair(oracle, frame, shift, nf=1200):
    left = oracle.blocks_v1(W, H, 1, frame)
    right = oracle.blocks_v1(W, H, 1, frame, shift_x=shift, noise_seed=2)
    el, er = oracle.Extractor(nf), oracle.Extractor(nf)
    _, kl, dl = el(left)
    _, kr, dr = er(right)
    ll = [el.level(l, with_border=True) for l in range(8)]
    lr = [er.level(l, with_border=True) for l in range(8)]
    t = el.tables()
    return kl, dl, kr, dr, ll, lr, t["scale"], t["inv_scale"]


@pytest.mark.parametrize("frame,shift", [(3, 12), (5, 30), (6, 0), (7, 47)])
def test_compute_stereo_matches_identical(oracle, frame, shift):
    """BASELINE config 2 through the reference's Frame::ComputeStereoMatches vs row band + SAD refinement + median cut."""
    kl, dl, kr, dr, ll, lr, sf, isf = _stereo_pair(oracle, frame, shift)
    want_ur, want_dp = R.stereo_matches(ll, lr, kl, dl, kr, dr, sf, isf, BF, MB)
    max_d = float(BF / MB)                                     # frame.cc:853-856: minZ = mb, maxD = bf / minZ
    bi, bd = oracle.stereo_rowband(kl, dl, kr, dr, sf, H, 0.0, max_d)
    ur, dp, sad = oracle.stereo_refine(ll, lr, kl, kr, bi, bd, sf, isf, 75, 0.0, max_d, BF)
    assert ur.tobytes() == want_ur.tobytes() and dp.tobytes() == want_dp.tobytes()
    assert (want_ur >= 0).sum() > 300


def test_compute_stereo_matches_small_baseline(oracle):
    """A tight disparity range (maxD = 20 px) rejects most candidates in the row band and in the disparity gate."""
    kl, dl, kr, dr, ll, lr, sf, isf = _stereo_pair(oracle, 4, 12, nf=800)
    bf, mb = np.float32(20.0), np.float32(1.0)
    want_ur, want_dp = R.stereo_matches(ll, lr, kl, dl, kr, dr, sf, isf, bf, mb)
    bi, bd = oracle.stereo_rowband(kl, dl, kr, dr, sf, H, 0.0, float(bf / mb))
    ur, dp, _ = oracle.stereo_refine(ll, lr, kl, kr, bi, bd, sf, isf, 75, 0.0, float(bf / mb), bf)
    assert ur.tobytes() == want_ur.tobytes() and dp.tobytes() == want_dp.tobytes()
    assert 50 < (want_ur >= 0).sum() < len(kl)


def test_distinctive_descriptors_identical(oracle):
    rng = np.random.default_rng(11)
    sizes = np.concatenate([[0, 1, 2, 3], rng.integers(2, 40, 200), [150]])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = np.empty((offsets[-1], 32), np.uint8)
    for p, (o, n) in enumerate(zip(offsets[:-1], sizes)):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        desc[o:o + n] = base
        flips = rng.integers(0, 256, (n, 20))
        for j in range(20):
            desc[o + np.arange(n), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
        if n > 4 and p % 3 == 0:
            desc[o + 1] = desc[o]                              # duplicates: equal medians, the first row wins
    bi, _ = oracle.distinctive(desc, offsets)
    want, chosen = R.distinctive(desc, offsets)
    assert np.array_equal(chosen != 0, bi >= 0)
    for p in range(len(sizes)):
        if bi[p] >= 0:
            assert np.array_equal(desc[offsets[p] + bi[p]], want[p]), p


def _frame_and_points(oracle, seed, nq=600):
    img = oracle.blocks_v1(W, H, seed, 0)
    _, kps, desc = oracle.Extractor(1000)(img)
    rng = np.random.default_rng(seed)
    src = rng.integers(0, len(kps), nq)
    pts = np.zeros(nq, R.TRACK_POINT_DTYPE)
    pts["proj_x"] = kps["x"][src] + rng.normal(0, 4, nq).astype(np.float32)
    pts["proj_y"] = kps["y"][src] + rng.normal(0, 4, nq).astype(np.float32)
    pts["view_cos"] = rng.choice([0.9, 0.9985, 1.0], nq).astype(np.float32)
    pts["depth"] = rng.uniform(1, 80, nq).astype(np.float32)
    pts["level"] = np.clip(kps["octave"][src] + rng.integers(-1, 2, nq), 0, 7)
    pts["in_view"] = rng.random(nq) < 0.9
    pts["bad"] = rng.random(nq) < 0.05
    pts["proj_x"][:10] = -300.0                                 # windows that miss the grid
    pts["proj_y"][10:20] = 3000.0
    qdesc = desc[src].copy()
    flips = rng.integers(0, 256, (nq, 14))
    for j in range(14):
        qdesc[np.arange(nq), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
    qdesc[::9] = desc[src[::9]]                                 # distance-0 ties
    return kps, desc, pts, qdesc, src, rng


def projection_windows(oracle, sf, pts, th, far, th_far):
    """orb_matcher.cc:50-70: the map points that pass the entry tests and their windows (the host side of the call)."""
    keep = [i for i, p in enumerate(pts) if p["in_view"] and not (far and p["depth"] > th_far) and not p["bad"]]
    q = np.zeros(len(keep), oracle.WQ_DTYPE)
    for j, i in enumerate(keep):
        p = pts[i]
        r = np.float32(2.5) if p["view_cos"] > 0.998 else np.float32(4.0)       # RadiusByViewingCos :208-213
        if th != 1.0:
            r = np.float32(r * np.float32(th))
        q[j] = (p["proj_x"], p["proj_y"], np.float32(r * sf[p["level"]]), p["level"] - 1, p["level"])
    return np.array(keep, np.int64), q


def _greedy_with_oracle(oracle, kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, th_far):
    """The whole SearchByProjection through the oracle: windows on the host side, search + greedy claim in orc_search_by_projection."""
    min_x, max_x, min_y, max_y = bounds
    geom = (min_x, min_y, np.float32(64) / np.float32(max_x - min_x), np.float32(48) / np.float32(max_y - min_y), 64, 48)
    keep, q = projection_windows(oracle, sf, pts, th, far, th_far)
    if u_right is None:
        nm, assigned = oracle.search_by_projection(kps, desc, geom, q, qdesc[keep], pre, None, None, None, 100, nnratio)
    else:
        nm, assigned = oracle.search_by_projection(kps, desc, geom, q, qdesc[keep], pre, u_right, pts["proj_xr"][keep], q["r"], 100, nnratio)
    return nm, np.where(assigned >= 0, keep[np.maximum(assigned, 0)], -1).astype(np.int32)


@pytest.mark.parametrize("seed,th,nnratio,stereo,far", [(1, 3.0, 0.8, False, False), (2, 1.0, 0.8, False, True),
                                                        (3, 5.0, 0.9, True, False), (4, 15.0, 0.6, True, True)])
def test_search_by_projection_identical(oracle, seed, th, nnratio, stereo, far):
    kps, desc, pts, qdesc, src, rng = _frame_and_points(oracle, seed)
    bounds = (0.0, float(W), 0.0, float(H))
    sf = oracle.Extractor(1000).tables()["scale"]
    pre = (rng.random(len(kps)) < 0.2).astype(np.uint8)
    u_right = None
    if stereo:
        u_right = np.where(rng.random(len(kps)) < 0.6, kps["x"] - rng.uniform(2, 60, len(kps)), -1.0).astype(np.float32)
        pts["proj_xr"] = pts["proj_x"] - rng.uniform(2, 60, len(pts)).astype(np.float32)
        ok = u_right[src] > 0
        pts["proj_xr"][ok] = u_right[src][ok] + rng.normal(0, 2, ok.sum()).astype(np.float32)
    want_nm, want = R.search_by_projection(kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, 40.0)
    nm, got = _greedy_with_oracle(oracle, kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, 40.0)
    assert nm == want_nm and np.array_equal(got, want)
    assert want_nm > 100


def test_features_in_area_order(oracle):
    """The visiting order of the grid lookup (cells column-major, keypoints in index order inside a cell) is what
    resolves distance ties in every window search: the oracle's best index on all-equal descriptors must be the
    first index the reference's GetFeaturesInArea returns."""
    img = oracle.blocks_v1(W, H, 5, 0)
    _, kps, _ = oracle.Extractor(1000)(img)
    bounds = (0.0, float(W), 0.0, float(H))
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    same = np.zeros((len(kps), 32), np.uint8)
    rng = np.random.default_rng(9)
    for _ in range(200):
        x, y = rng.uniform(-20, W + 20), rng.uniform(-20, H + 20)
        r = float(rng.choice([2.0, 9.0, 30.0, 90.0]))
        lo, hi = int(rng.integers(-1, 6)), int(rng.integers(-1, 8))
        idx = R.features_in_area(kps, bounds, x, y, r, lo, hi)
        q = np.zeros(1, oracle.WQ_DTYPE)
        q["u"], q["v"], q["r"], q["min_level"], q["max_level"] = x, y, r, lo, hi
        res = oracle.window_search(kps, same, geom, q, same[:1])[0]
        assert res["best_idx"] == (idx[0] if len(idx) else -1)


def _bow_pair(oracle, seed, shift, k=6, L=4, levelsup=2, nf=800):
    """Two views of one scene (a key frame and a frame shifted by `shift` px with other noise) and their FeatureVectors."""
    vocab = oracle.Vocabulary(k, L, *oracle.synth_vocab(k, L, seed=seed + 40))
    a = oracle.blocks_v1(640, 400, seed, 0)
    b = oracle.blocks_v1(640, 400, seed, 0, shift_x=shift, noise_seed=seed + 7)
    ex = oracle.Extractor(nf)
    _, ka, da = ex(a)
    _, kb, db = ex(b)
    fva = oracle.pack_feature_vector(*vocab.transform(da, levelsup)[2:])
    fvb = oracle.pack_feature_vector(*vocab.transform(db, levelsup)[2:])
    return ka, da, fva, kb, db, fvb


@pytest.mark.parametrize("seed,shift,ratio,ori,levelsup", [(1, 3, 0.7, True, 2), (2, 8, 0.75, True, 3), (3, 0, 0.9, False, 2),
                                                           (4, 5, 0.6, True, 1), (5, 2, 0.7, True, 4)])
def test_search_by_bow_identical(oracle, seed, shift, ratio, ori, levelsup):
    ka, da, fva, kb, db, fvb = _bow_pair(oracle, seed, shift, levelsup=levelsup)
    rng = np.random.default_rng(seed)
    has_point = (rng.random(len(ka)) < 0.8).astype(np.uint8)
    want_nm, want = R.search_by_bow(ka, da, has_point, fva, kb, db, fvb, ratio, ori)
    nm, got = oracle.search_by_bow(ka, da, has_point, fva, kb, db, fvb, ratio, ori)
    assert nm == want_nm and np.array_equal(got, want)
    assert want_nm > 30
    # all key-frame features hold points; identical frames -> every match is the feature itself
    nm2, got2 = oracle.search_by_bow(ka, da, None, fva, ka, da, fva, ratio, ori)
    want_nm2, want2 = R.search_by_bow(ka, da, None, fva, ka, da, fva, ratio, ori)
    assert nm2 == want_nm2 and np.array_equal(got2, want2)
    hit = got2 >= 0
    assert hit.sum() > 100 and (da[got2[hit]] == da[np.nonzero(hit)[0]]).all()


@pytest.mark.parametrize("seed,shift,ratio,ori,levelsup", [(1, 3, 0.9, True, 2), (2, 8, 0.75, True, 3), (3, 0, 0.9, False, 2),
                                                           (4, 5, 0.6, True, 1), (5, 2, 0.8, True, 4)])
def test_search_by_bow_keyframes_identical(oracle, seed, shift, ratio, ori, levelsup):
    """The loop-closing form SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cc:697-815)."""
    ka, da, fva, kb, db, fvb = _bow_pair(oracle, seed, shift, levelsup=levelsup)
    rng = np.random.default_rng(seed + 100)
    hp1 = (rng.random(len(ka)) < 0.8).astype(np.uint8)
    hp2 = (rng.random(len(kb)) < 0.8).astype(np.uint8)
    want_nm, want = R.search_by_bow_kf(ka, da, hp1, fva, kb, db, hp2, fvb, ratio, ori)
    nm, got = oracle.search_by_bow_kf(ka, da, hp1, fva, kb, db, hp2, fvb, ratio, ori)
    assert nm == want_nm and np.array_equal(got, want)
    assert want_nm > 20
    want_nm2, want2 = R.search_by_bow_kf(ka, da, None, fva, ka, da, None, fva, ratio, ori)
    nm2, got2 = oracle.search_by_bow_kf(ka, da, None, fva, ka, da, None, fva, ratio, ori)
    assert nm2 == want_nm2 and np.array_equal(got2, want2) and nm2 > 100


CAM4 = np.array([458.654, 457.296, 367.215, 248.375], np.float32)          # settings/EuRoC.yaml Camera1 fx fy cx cy
BF_LAST, MB_LAST = np.float32(47.90639384423901), np.float32(0.11)


def last_frame_case(oracle, seed, n_last=700):
    """A current frame and a `last frame` whose map points project near the current keypoints (translation-only poses)."""
    img = oracle.blocks_v1(W, H, seed, 0)
    _, kps, desc = oracle.Extractor(1000)(img)
    rng = np.random.default_rng(seed + 50)
    src = rng.integers(0, len(kps), n_last)
    t_cw = np.array([0.10, -0.05, 0.20], np.float32)
    z = rng.uniform(2.0, 30.0, n_last).astype(np.float32)
    u = (kps["x"][src] + rng.normal(0, 3, n_last)).astype(np.float32)
    v = (kps["y"][src] + rng.normal(0, 3, n_last)).astype(np.float32)
    xc = ((u - CAM4[2]) * z / CAM4[0]).astype(np.float32)
    yc = ((v - CAM4[3]) * z / CAM4[1]).astype(np.float32)
    world = np.stack([xc - t_cw[0], yc - t_cw[1], z - t_cw[2]], 1).astype(np.float32)
    world[:15, 2] = -5.0                                                       # behind the camera
    world[15:30, 0] += 500.0                                                   # projects outside the image
    last = np.zeros(n_last, oracle.KP_DTYPE)
    last["octave"] = np.clip(kps["octave"][src] + rng.integers(-1, 2, n_last), 0, 7)
    last["angle"] = np.where(rng.random(n_last) < 0.8, kps["angle"][src] + rng.normal(0, 4, n_last), rng.uniform(0, 360, n_last)) % 360
    last["angle"] = last["angle"].astype(np.float32)
    ldesc = desc[src].copy()
    flips = rng.integers(0, 256, (n_last, 14))
    for j in range(14):
        ldesc[np.arange(n_last), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
    has_point = (rng.random(n_last) < 0.85).astype(np.uint8)
    outlier = (rng.random(n_last) < 0.05).astype(np.uint8)
    return kps, desc, last, ldesc, has_point, outlier, world, t_cw, src, rng


def last_frame_windows(oracle, sf, last, has_point, outlier, world, t_cw, t_lw, th, mono, bounds):
    """orb_matcher.cc:1529-1576 and :1587 in float32, the host side of the call: (kept last-frame indices, windows, u_right, angles)."""
    twc = -t_cw
    tlc_z = np.float32(twc[2] + t_lw[2])
    forward, backward = bool(tlc_z > MB_LAST and not mono), bool(-tlc_z > MB_LAST and not mono)
    keep, q, qur = [], [], []
    for i in range(len(last)):
        if not has_point[i] or outlier[i]:
            continue
        xc, yc, zc = (np.float32(world[i, k] + t_cw[k]) for k in range(3))
        invzc = np.float32(1.0 / np.float64(zc))
        if invzc < 0:
            continue
        u = np.float32(np.float32(np.float32(CAM4[0] * xc) / zc) + CAM4[2])
        v = np.float32(np.float32(np.float32(CAM4[1] * yc) / zc) + CAM4[3])
        if u < bounds[0] or u > bounds[1] or v < bounds[2] or v > bounds[3]:
            continue
        o = int(last["octave"][i])
        radius = np.float32(np.float32(th) * sf[o])
        lo, hi = (o, -1) if forward else ((0, o) if backward else (o - 1, o + 1))
        keep.append(i)
        q.append((u, v, radius, lo, hi))
        qur.append(np.float32(u - np.float32(BF_LAST * invzc)))
    return np.array(keep, np.int64), np.array(q, oracle.WQ_DTYPE), np.array(qur, np.float32)


@pytest.mark.parametrize("seed,th,t_lw_z,mono,stereo,ori", [(1, 7.0, 0.0, False, True, True), (2, 15.0, 0.5, False, True, True),
                                                            (3, 15.0, -0.6, False, False, True), (4, 7.0, 0.5, True, False, False)])
def test_search_by_projection_last_frame_identical(oracle, seed, th, t_lw_z, mono, stereo, ori):
    """SearchByProjection(CurrentFrame, LastFrame, th, bMono) (orb_matcher.cc:1518-1728): forward / backward / neutral level
    windows, stereo gate, greedy claim, rotation histogram."""
    kps, desc, last, ldesc, has_point, outlier, world, t_cw, src, rng = last_frame_case(oracle, seed)
    bounds = (0.0, float(W), 0.0, float(H))
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    sf = oracle.Extractor(1000).tables()["scale"]
    t_lw = np.array([0.0, 0.0, t_lw_z], np.float32)
    pre = (rng.random(len(kps)) < 0.15).astype(np.uint8)
    u_right = np.where(rng.random(len(kps)) < 0.6, kps["x"] - rng.uniform(2, 40, len(kps)), -1.0).astype(np.float32) if stereo else None
    want_nm, want = R.search_by_projection_last(kps, desc, bounds, sf, BF_LAST, MB_LAST, CAM4, t_cw, t_lw, last, has_point, outlier,
                                                world, ldesc, pre, u_right, th, mono, ori)
    keep, q, qur = last_frame_windows(oracle, sf, last, has_point, outlier, world, t_cw, t_lw, th, mono, bounds)
    nm, got = oracle.search_by_projection_last(kps, desc, geom, q, ldesc[keep], last["angle"][keep], pre, u_right,
                                               qur if stereo else None, q["r"] if stereo else None, 100, ori)
    assert nm == want_nm and np.array_equal(np.where(got >= 0, keep[np.maximum(got, 0)], -1), want)
    assert want_nm > 100


def triangulation_case(oracle, seed, shift, levelsup=3):
    ka, da, fva, kb, db, fvb = _bow_pair(oracle, seed, shift, levelsup=levelsup)
    rng = np.random.default_rng(seed + 300)
    hp1 = (rng.random(len(ka)) < 0.4).astype(np.uint8)          # features that already have a map point are skipped
    hp2 = (rng.random(len(kb)) < 0.4).astype(np.uint8)
    ur1 = np.where(rng.random(len(ka)) < 0.5, ka["x"] - rng.uniform(1, 40, len(ka)), -1.0).astype(np.float32)
    ur2 = np.where(rng.random(len(kb)) < 0.5, kb["x"] - rng.uniform(1, 40, len(kb)), -1.0).astype(np.float32)
    tx = np.float32(0.11)
    f12 = np.zeros((3, 3), np.float32)                          # pure x translation, equal pinhole cameras: horizontal lines
    f12[1, 2], f12[2, 1] = -tx / CAM4[1], tx / CAM4[1]
    f12 += rng.normal(0, 2e-9, (3, 3)).astype(np.float32)       # ... slightly tilted
    t = oracle.Extractor(800).tables()
    return ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, t["scale"], t["sigma2"]


@pytest.mark.parametrize("seed,shift,only_stereo,coarse,ori,c2", [(1, 12, False, False, True, (0.5, 0.01, 0.05)),
                                                                  (2, 6, False, False, True, (0.02, -0.01, 1.0)),
                                                                  (3, 9, True, False, True, (0.02, -0.01, 1.0)),
                                                                  (4, 12, False, True, False, (-0.3, 0.2, 1.0))])
def test_search_for_triangulation_identical(oracle, seed, shift, only_stereo, coarse, ori, c2):
    """SearchForTriangulation (orb_matcher.cc:817-1040) + the line test of Pinhole::EpipolarConstrain (pinhole_model.cc:121-134)."""
    ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, sf, s2 = triangulation_case(oracle, seed, shift)
    want_nm, want, ep = R.search_for_triangulation(ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, CAM4, c2, sf, s2,
                                                   only_stereo, coarse, ori)
    nm, got = oracle.search_for_triangulation(ka, da, hp1, ur1, fva, kb, db, hp2, ur2, fvb, f12, ep, sf, s2, only_stereo, coarse, ori)
    assert nm == want_nm and np.array_equal(got, want)
    assert want_nm > (10 if only_stereo else 25)
    assert (got[hp1 != 0] == -1).all() and (hp2[got[got >= 0]] == 0).all()


def fuse_case(oracle, seed, nq=700):
    img = oracle.blocks_v1(W, H, seed, 0)
    _, kps, desc = oracle.Extractor(1000)(img)
    rng = np.random.default_rng(seed + 900)
    src = rng.integers(0, len(kps), nq)
    t = oracle.Extractor(1000).tables()
    lev = np.clip(kps["octave"][src] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32)
    q = np.zeros(nq, oracle.WQ_DTYPE)
    q["u"] = kps["x"][src] + rng.normal(0, 2.5, nq).astype(np.float32)     # chi-square gate: some pass, some do not
    q["v"] = kps["y"][src] + rng.normal(0, 2.5, nq).astype(np.float32)
    q["r"] = (np.float32(3.0) * t["scale"][lev]).astype(np.float32)
    q["min_level"], q["max_level"] = lev - 1, lev
    u_right = np.where(rng.random(len(kps)) < 0.5, kps["x"] - rng.uniform(1, 40, len(kps)), -1.0).astype(np.float32)
    qur = np.where(u_right[src] >= 0, u_right[src] + rng.normal(0, 2.0, nq), q["u"] - 20).astype(np.float32)
    qdesc = desc[src].copy()
    flips = rng.integers(0, 256, (nq, 14))
    for j in range(14):
        qdesc[np.arange(nq), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
    return kps, desc, q, qdesc, qur, lev, u_right, t["inv_sigma2"]


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False), (3, True)])
def test_fuse_search_identical(oracle, seed, stereo):
    """The candidate loop of ORBmatcher::Fuse (orb_matcher.cc:1145-1190): level gate, chi-square reprojection gate, nearest."""
    kps, desc, q, qdesc, qur, lev, u_right, inv_s2 = fuse_case(oracle, seed)
    bounds = (0.0, float(W), 0.0, float(H))
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    wi, wd = R.fuse_search(kps, desc, bounds, inv_s2, q["u"], q["v"], qur, q["r"], lev, qdesc, u_right if stereo else None)
    got = oracle.window_search_fuse(kps, desc, geom, q, qdesc, inv_s2, u_right if stereo else None, qur if stereo else None)
    assert np.array_equal(got["best_idx"], wi) and np.array_equal(got["best_dist"], wd)
    assert (wi >= 0).mean() > 0.3 and (wi < 0).mean() > 0.05
