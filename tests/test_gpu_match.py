"""GPU parity of the Hamming matching kernels through the C ABI against the golden BFMatcher vectors
and the CPU oracle: distances, 2-NN with ties, sharded merge, ratio test, stereo row band, windows."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import orb_slam_fusion_b200 as P
    return P


@pytest.fixture(scope="module")
def m(P):
    return P.ORBmatcher(0.7, True)


def test_descriptor_distance(P, m, oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    b[:100] = a[:100]
    b[100:200] = ~a[100:200]
    want = np.array([oracle.hamming(x, y) for x, y in zip(a, b)], np.int32)
    assert np.array_equal(m.DescriptorDistance(a, b), want)
    assert want[:100].max() == 0 and want[100:200].min() == 256
    assert m.DescriptorDistance(a[7], b[7]) == want[7]
    # 4-byte aligned rows, as cv::Mat::ptr<int32_t> (orb_matcher.cc:1879-1880)
    buf = np.zeros(5000 * 32 + 4, np.uint8)
    buf[4:] = a.reshape(-1)
    assert np.array_equal(m.DescriptorDistance(buf[4:].reshape(-1, 32), b), want)
    assert P.ORBmatcher.TH_LOW == 50 and P.ORBmatcher.TH_HIGH == 100 and P.ORBmatcher.HISTO_LENGTH == 30


def test_knn2_golden_bfmatcher(m, oracle):
    g = load_golden("knn2_200x5000")
    idx, dist = m.knn2(g["q"], g["db"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    assert np.array_equal(m.ratio_test(idx, dist, 0.7), g["accept"])


def test_knn2_ties_and_small_databases(m, oracle):
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (1100, 32), dtype=np.uint8)       # > one 1024-query block
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    db = base[rng.integers(0, 40, 3000)].copy()                 # many exact duplicates -> ties
    flips = rng.integers(0, 256, 3000)
    db[np.arange(3000), flips % 32] ^= (1 << (flips // 32)).astype(np.uint8)
    q[:40] = base
    for nd in (3000, 257, 256, 2, 1, 0):
        idx, dist = m.knn2(q, db[:nd])
        ri, rd = oracle.knn2(q, db[:nd])
        assert np.array_equal(idx, ri) and np.array_equal(dist, rd), nd
        acc = m.ratio_test(idx, dist, 0.7)
        assert np.array_equal(acc, oracle.ratio_accept(ri, rd, 0.7))
    idx, dist = m.knn2(q[:0], db)
    assert idx.shape == (0, 2)


def test_sharded_top2_merge_equals_whole(m, oracle):
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    db = rng.integers(0, 4, (20000, 32), dtype=np.uint8)        # low-entropy rows: plenty of distance ties
    whole_i, whole_d = m.knn2(q, db)
    ri, rd = oracle.knn2(q, db)
    assert np.array_equal(whole_i, ri) and np.array_equal(whole_d, rd)
    for G in (2, 3, 8):
        bounds = [len(db) * s // G for s in range(G + 1)]
        parts = [m.knn2(q, db[bounds[s]:bounds[s + 1]], index_base=bounds[s]) for s in range(G)]
        pi = np.stack([p[0] for p in parts])
        pd = np.stack([p[1] for p in parts])
        mi, md = m.top2_merge(pi[::-1].copy(), pd[::-1].copy())   # shard order must not matter
        assert np.array_equal(mi, whole_i) and np.array_equal(md, whole_d), G
    # a shard with a single row contributes one neighbour and a missing entry
    parts = [m.knn2(q, db[:1]), m.knn2(q, db[1:], index_base=1)]
    mi, md = m.top2_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    assert np.array_equal(mi, whole_i) and np.array_equal(md, whole_d)


def test_config5_reduced_with_planted_matches(P, m, oracle):
    """BASELINE config 5 at 1/50 size: device-generated database, planted matches and decoys."""
    import torch
    nq, nd = 1000, 200_000
    db = P.synth_descriptors(0, nd, seed=101)
    q = P.synth_descriptors(0, nq, seed=202)
    assert np.array_equal(db[:64].cpu().numpy(), oracle.synth_descriptors(0, 64, 101))
    dbh, qh = db.cpu().numpy().copy(), q.cpu().numpy()
    rng = np.random.default_rng(1)
    rows = rng.choice(nd, nq, replace=False)
    for i in range(nq):
        d = qh[i].copy()
        bits = rng.choice(256, 5 + int(rng.integers(0, 36)), replace=False)
        np.bitwise_xor.at(d, bits // 8, (1 << (bits % 8)).astype(np.uint8))
        dbh[rows[i]] = d
        if i % 10 == 0:
            e = d.copy()
            bits = rng.choice(256, int(rng.integers(0, 11)), replace=False)
            np.bitwise_xor.at(e, bits // 8, (1 << (bits % 8)).astype(np.uint8))
            dbh[(rows[i] + 1) % nd] = e
    db = torch.from_numpy(dbh).cuda()
    idx, dist = m.knn2(q, db)
    acc = m.ratio_test(idx, dist, 0.7)
    torch.cuda.synchronize()
    ri, rd = oracle.knn2(qh, dbh, nthreads=8)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd)
    racc = oracle.ratio_accept(ri, rd, 0.7)
    assert np.array_equal(acc.cpu().numpy().astype(bool), racc)
    assert 0.85 < racc.mean() < 0.95     # the decoys (every 10th query) fail the ratio test


def test_full_size_database_properties(P, m):
    """Config 5 at full size (1000 x 10M): the planted rows are found (known answer by construction)
    and an 8-way sharded search merged with top2_merge equals the single search."""
    import torch
    nq, nd = 1000, 10_000_000
    db = P.synth_descriptors(0, nd, seed=7)
    q = P.synth_descriptors(0, nq, seed=8)
    rows = torch.arange(nq, device="cuda", dtype=torch.int64) * 9973 + 12345
    planted = q.clone()
    planted[:, 0] ^= 0xFF          # 8 flipped bits: distance 8, random rows sit near 128
    db[rows] = planted
    idx, dist = m.knn2(q, db)
    torch.cuda.synchronize()
    assert torch.equal(idx[:, 0], rows) and (dist[:, 0] == 8).all()
    assert (dist[:, 1] > 60).all() and (dist[:, 1] < 128).all()
    assert m.ratio_test(idx, dist, 0.7).all()
    G = 8
    parts = [m.knn2(q, db[nd * s // G: nd * (s + 1) // G], index_base=nd * s // G) for s in range(G)]
    mi, md = m.top2_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    torch.cuda.synchronize()
    assert torch.equal(mi, idx) and torch.equal(md, dist)


def test_stereo_rowband_config2(P, m, oracle):
    """BASELINE config 2: EuRoC-shaped stereo pair, 1200 features per side, left-right search."""
    w, h = 752, 480
    left = oracle.blocks_v1(w, h, 1, 3)
    right = oracle.blocks_v1(w, h, 1, 3, shift_x=12, noise_seed=2)
    exl, exr = P.OrbExtractor(1200, 1.2, 8, 20, 7), P.OrbExtractor(1200, 1.2, 8, 20, 7)
    _, kl, dl = exl(left)
    _, kr, dr = exr(right)
    sf = exl.GetScaleFactors()
    for (min_d, max_d) in [(0.0, 458.654), (0.0, 20.0), (5.0, 60.0)]:
        bi, bd = m.stereo_rowband(kl, dl, kr, dr, sf, h, min_d, max_d)
        ri, rd = oracle.stereo_rowband(kl, dl, kr, dr, sf, h, min_d, max_d)
        assert np.array_equal(bi, ri) and np.array_equal(bd, rd)
    assert (ri >= 0).sum() > 300
    # degenerate sides
    bi, bd = m.stereo_rowband(kl, dl, kr[:0], dr[:0], sf, h, 0.0, 458.654)
    assert (bi == -1).all() and (bd == 100).all()


def test_stereo_refine_and_full_stereo_matches(P, m, oracle):
    """the SAD / parabola / median part of Frame::ComputeStereoMatches
    (frame.cc:903-985) on the two extractors' device pyramids, BASELINE config 2 end to end."""
    w, h = 752, 480
    exl, exr = P.OrbExtractor(1200, 1.2, 8, 20, 7), P.OrbExtractor(1200, 1.2, 8, 20, 7)
    rl, rr = oracle.Extractor(1200, 1.2, 8, 20, 7), oracle.Extractor(1200, 1.2, 8, 20, 7)
    bf, mb = np.float32(47.90639384423901), np.float32(0.11)     # EuRoC.yaml: fx * baseline, ThDepth-free minZ
    for frame, shift in [(3, 12), (5, 30), (6, 0)]:
        left = oracle.blocks_v1(w, h, 1, frame)
        right = oracle.blocks_v1(w, h, 1, frame, shift_x=shift, noise_seed=2)
        _, kl, dl = exl(left)
        _, kr, dr = exr(right)
        rl.compute_pyramid(left)
        rr.compute_pyramid(right)
        ll = [rl.level(l, with_border=True) for l in range(8)]
        lr = [rr.level(l, with_border=True) for l in range(8)]
        sf, isf = exl.GetScaleFactors(), exl.GetInverseScaleFactors()
        max_d = float(bf / mb)
        bi, bd = m.stereo_rowband(kl, dl, kr, dr, sf, h, 0.0, max_d)
        ur, dp, sad = m.stereo_refine(exl, exr, kl, kr, bi, bd, 0.0, max_d, bf)
        wur, wdp, wsad = oracle.stereo_refine(ll, lr, kl, kr, bi, bd, sf, isf, 75, 0.0, max_d, bf)
        assert np.array_equal(sad, wsad) and ur.tobytes() == wur.tobytes() and dp.tobytes() == wdp.tobytes()
        ur2, dp2 = m.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, bf, mb)
        assert ur2.tobytes() == wur.tobytes() and dp2.tobytes() == wdp.tobytes()
        # the one-call form on the device-resident results of the two extractors, images in flight together
        exl.extract_begin(left)
        exr.extract_begin(right)
        kl2, dl2 = np.empty(len(kl) + 8, P.KP_DTYPE), np.empty((len(kl) + 8, 32), np.uint8)
        kr2, dr2 = np.empty(len(kr) + 8, P.KP_DTYPE), np.empty((len(kr) + 8, 32), np.uint8)
        _, nl2 = exl.extract_end(kl2, dl2)
        _, nr2 = exr.extract_end(kr2, dr2)
        assert kl2[:nl2].tobytes() == kl.tobytes() and np.array_equal(dr2[:nr2], dr)
        ur3, dp3 = m.stereo_matches_last(exl, exr, bf, mb)
        assert ur3.tobytes() == wur.tobytes() and dp3.tobytes() == wdp.tobytes()
        ok = wur >= 0
        assert ok.sum() > 400
        assert abs(np.median(kl["x"][ok] - wur[ok]) - shift) < 0.6 or shift == 0
        assert (wsad[~ok & (wsad >= 0)]).size > 0        # the median cut dropped some accepted matches
    # nothing to refine
    ur, dp, sad = m.stereo_refine(exl, exr, kl, kr, np.full(len(kl), -1, np.int32), np.full(len(kl), 100, np.int32), 0.0, 400.0, bf)
    assert (ur == -1).all() and (dp == -1).all() and (sad == -1).all()


def test_distinctive_descriptors(m, oracle):
    """MapPoint::ComputeDistinctiveDescriptors batched over map points."""
    rng = np.random.default_rng(11)
    sizes = np.concatenate([[0, 1, 2, 3], rng.integers(2, 40, 300), [150, 257]])
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = np.empty((offsets[-1], 32), np.uint8)
    for p, (o, n) in enumerate(zip(offsets[:-1], sizes)):
        base = rng.integers(0, 256, 32, dtype=np.uint8)          # observations of one point: noisy copies
        desc[o:o + n] = base
        flips = rng.integers(0, 256, (n, 20))
        for j in range(20):
            desc[o + np.arange(n), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
        if n > 4 and p % 3 == 0:
            desc[o + 1] = desc[o]                                  # duplicates: equal medians, first row wins
    bi, bm = m.ComputeDistinctiveDescriptors(desc, offsets)
    wi, wm = oracle.distinctive(desc, offsets)
    assert np.array_equal(bi, wi) and np.array_equal(bm, wm)
    assert bi[0] == -1 and bi[1] == 0 and bm[1] == 0


def test_window_search(P, m, oracle):
    w, h = 752, 480
    img = oracle.blocks_v1(w, h, 1, 0)
    ex = P.OrbExtractor(1000, 1.2, 8, 20, 7)
    _, kps, desc = ex(img)
    geom = (0.0, 0.0, 64.0 / w, 48.0 / h, 64, 48)       # frame.cc:199-204, frame.h:40-41
    rng = np.random.default_rng(2)
    nq = 1500
    src = rng.integers(0, len(kps), nq)
    q = np.zeros(nq, P.WQ_DTYPE)
    q["u"] = kps["x"][src] + rng.normal(0, 6, nq).astype(np.float32)
    q["v"] = kps["y"][src] + rng.normal(0, 6, nq).astype(np.float32)
    q["r"] = rng.choice([3.0, 7.0, 15.0, 40.0, 120.0], nq).astype(np.float32)
    lv = kps["octave"][src]
    q["min_level"] = np.where(rng.random(nq) < 0.5, lv - 1, -1)
    q["max_level"] = np.where(rng.random(nq) < 0.5, lv + 1, -1)
    q["u"][:20] = -500.0                                  # windows that miss the image
    q["v"][20:40] = 5000.0
    qdesc = desc[src].copy()
    flips = rng.integers(0, 256, (nq, 12))
    for j in range(12):
        qdesc[np.arange(nq), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
    qdesc[::7] = desc[src[::7]]                           # exact duplicates -> distance-0 ties
    for skip in (None, (rng.random(len(kps)) < 0.3).astype(np.uint8)):
        got = m.window_search(kps, desc, geom, q, qdesc, skip)
        want = oracle.window_search(kps, desc, geom, q, qdesc, skip)
        assert got.tobytes() == want.tobytes()
    assert (want["best_idx"] >= 0).mean() > 0.5
    # duplicated keypoints make (distance) ties that only the visiting order resolves
    k2 = np.concatenate([kps, kps[:200]])
    d2 = np.concatenate([desc, desc[:200]])
    got = m.window_search(k2, d2, geom, q, qdesc)
    want = oracle.window_search(k2, d2, geom, q, qdesc)
    assert got.tobytes() == want.tobytes()
    # stereo gate (orb_matcher.cc:89-92, 1586-1590): keypoints with a right coordinate must agree with the
    # projection in the right image; ~half the keypoints are monocular (mvuRight = -1)
    ur = np.where(rng.random(len(kps)) < 0.5, kps["x"] - rng.uniform(2, 60, len(kps)), -1.0).astype(np.float32)
    qur = (q["u"] - rng.uniform(2, 60, nq)).astype(np.float32)
    qerr = (q["r"] * np.float32(1.2) ** np.maximum(lv, 0)).astype(np.float32)
    qur[::3] = ur[src[::3]] + rng.normal(0, 1, len(qur[::3])).astype(np.float32)     # consistent stereo projections
    for skip in (None, (rng.random(len(kps)) < 0.3).astype(np.uint8)):
        got = m.window_search(kps, desc, geom, q, qdesc, skip, ur, qur, qerr)
        want = oracle.window_search(kps, desc, geom, q, qdesc, skip, ur, qur, qerr)
        assert got.tobytes() == want.tobytes()
    plain = oracle.window_search(kps, desc, geom, q, qdesc, skip)
    assert (want["best_idx"] != plain["best_idx"]).mean() > 0.05 and (want["best_idx"] >= 0).mean() > 0.3


def test_popc_peak_microbenchmark(P):
    peak = P.popc_peak(0)
    assert 1e12 < peak < 2e13      # 148 SMs x 16 popc/clk x ~1.9 GHz = 4.5e12


def test_search_by_bow(P, m, oracle):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (orb_matcher.cc:215-389) batched over pairs of a frame pool:
    extraction -> bag-of-words transform -> node-restricted greedy matching -> rotation histogram, vs the CPU oracle
    (which tests/test_oracle_vs_ref_frame.py pins on the reference's own lines)."""
    w, h, nfeat, k, L = 640, 400, 800, 6, 4
    vp_ = oracle.synth_vocab(k, L, seed=41)
    voc, vo = P.ORBVocabulary(k, L, *vp_), oracle.Vocabulary(k, L, *vp_)
    ex = P.OrbExtractor(nfeat, 1.2, 8, 20, 7)
    views = [(1, 0, 1), (1, 3, 8), (1, 9, 9), (2, 0, 2), (2, 4, 5), (3, 0, 3)]      # (scene, shift, noise seed)
    frames = [oracle.blocks_v1(w, h, s, 0, shift_x=sh, noise_seed=ns) for s, sh, ns in views]
    ext = [ex(f) for f in frames]
    cap = max(len(e[1]) for e in ext) + 5
    F = len(frames)
    kps = np.zeros((F, cap), P.KP_DTYPE)
    desc = np.zeros((F, cap, 32), np.uint8)
    npf = np.zeros(F, np.int32)
    for f, (_, kk, dd) in enumerate(ext):
        kps[f, :len(kk)], desc[f, :len(kk)], npf[f] = kk, dd, len(kk)
    rng = np.random.default_rng(5)
    has_point = (rng.random((F, cap)) < 0.85).astype(np.uint8)
    pairs = np.array([(0, 1), (1, 0), (0, 2), (3, 4), (4, 3), (0, 3), (5, 5), (2, 1)], np.int32)
    for levelsup, ratio, ori in [(2, 0.7, True), (3, 0.75, True), (1, 0.9, False), (4, 0.6, True)]:
        fv = voc.transform_batch(desc, npf, levelsup)
        nm, match = m.SearchByBoW(kps, desc, npf, fv, pairs, has_point, ratio, ori)
        for p, (a, b) in enumerate(pairs):
            fva = oracle.pack_feature_vector(*vo.transform(desc[a, :npf[a]], levelsup)[2:])
            fvb = oracle.pack_feature_vector(*vo.transform(desc[b, :npf[b]], levelsup)[2:])
            wnm, want = oracle.search_by_bow(kps[a, :npf[a]], desc[a, :npf[a]], has_point[a, :npf[a]], fva,
                                             kps[b, :npf[b]], desc[b, :npf[b]], fvb, ratio, ori)
            assert nm[p] == wnm and np.array_equal(match[p, :npf[b]], want), (levelsup, p)
            assert (match[p, npf[b]:] == -1).all()
        assert nm[0] > 40 and nm[6] > 100 and nm[5] < nm[0]      # same scene matches, different scenes mostly do not
        # the loop-closing form SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cc:697-815)
        nm2, match2 = m.SearchByBoWKeyFrames(kps, desc, npf, fv, pairs, has_point, ratio, ori)
        for p, (a, b) in enumerate(pairs):
            fva = oracle.pack_feature_vector(*vo.transform(desc[a, :npf[a]], levelsup)[2:])
            fvb = oracle.pack_feature_vector(*vo.transform(desc[b, :npf[b]], levelsup)[2:])
            wnm, want = oracle.search_by_bow_kf(kps[a, :npf[a]], desc[a, :npf[a]], has_point[a, :npf[a]], fva,
                                                kps[b, :npf[b]], desc[b, :npf[b]], has_point[b, :npf[b]], fvb, ratio, ori)
            assert nm2[p] == wnm and np.array_equal(match2[p, :npf[a]], want), (levelsup, p)
            assert (match2[p, npf[a]:] == -1).all()
        assert nm2[0] > 30 and nm2[6] > 100
    # no map points at all / an empty pair list
    nm, match = m.SearchByBoW(kps, desc, npf, fv, pairs[:2], np.zeros((F, cap), np.uint8))
    assert (nm == 0).all() and (match == -1).all()
    nm, match = m.SearchByBoW(kps, desc, npf, fv, pairs[:0])
    assert len(nm) == 0
    # device-resident pool: CUDA tensors in, CUDA tensors out
    import torch
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a.view(np.uint8) if a.dtype.fields else a).to(dev)   # noqa: E731
    tk = torch.from_numpy(kps.view(np.float32).reshape(F, cap, 7).copy()).to(dev)
    td = t(desc)
    fvd = voc.transform_batch(td, t(npf), 2)
    nm_d, match_d = m.SearchByBoW(tk, td, t(npf), fvd, torch.from_numpy(pairs), t(has_point), 0.7, True)
    fvh = voc.transform_batch(desc, npf, 2)
    nm_h, match_h = m.SearchByBoW(kps, desc, npf, fvh, pairs, has_point, 0.7, True)
    torch.cuda.synchronize()
    assert np.array_equal(nm_d.cpu().numpy(), nm_h) and np.array_equal(match_d.cpu().numpy(), match_h)


@pytest.mark.parametrize("seed,th,nnratio,stereo,far", [(1, 3.0, 0.8, False, False), (2, 1.0, 0.8, False, True),
                                                        (3, 5.0, 0.9, True, False), (4, 15.0, 0.6, True, True)])
def test_search_by_projection_whole_function(P, m, oracle, seed, th, nnratio, stereo, far):
    """ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints) (orb_matcher.cc:42-134),
    greedy claim included, against the oracle and -- where oracle/_ref travelled -- the reference's own lines."""
    from oracle import ref as R
    from test_oracle_vs_ref_frame import _frame_and_points, projection_windows, W, H
    kps, desc, pts, qdesc, src, rng = _frame_and_points(oracle, seed, nq=900)     # 900 points on ~1000 keypoints: many conflicts
    bounds = (0.0, float(W), 0.0, float(H))
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    sf = oracle.Extractor(1000).tables()["scale"]
    pre = (rng.random(len(kps)) < 0.2).astype(np.uint8)
    u_right = qur = qerr = None
    keep, q = projection_windows(oracle, sf, pts, th, far, 40.0)
    if stereo:
        u_right = np.where(rng.random(len(kps)) < 0.6, kps["x"] - rng.uniform(2, 60, len(kps)), -1.0).astype(np.float32)
        pts["proj_xr"] = pts["proj_x"] - rng.uniform(2, 60, len(pts)).astype(np.float32)
        ok = u_right[src] > 0
        pts["proj_xr"][ok] = u_right[src][ok] + rng.normal(0, 2, ok.sum()).astype(np.float32)
        qur, qerr = pts["proj_xr"][keep], q["r"]
    nm, got = m.SearchByProjection(kps, desc, geom, q, qdesc[keep], pre, u_right, qur, qerr, 100, nnratio)
    wnm, want = oracle.search_by_projection(kps, desc, geom, q, qdesc[keep], pre, u_right, qur, qerr, 100, nnratio)
    assert nm == wnm and np.array_equal(got, want) and wnm > 150
    m.set_option(P.OPT_CLAIM_SEQUENTIAL, 1)            # the fallback kernel of frames too large for the shared-memory tables
    try:
        nm_s, got_s = m.SearchByProjection(kps, desc, geom, q, qdesc[keep], pre, u_right, qur, qerr, 100, nnratio)
    finally:
        m.set_option(P.OPT_CLAIM_SEQUENTIAL, 0)
    assert nm_s == wnm and np.array_equal(got_s, want)
    # the greedy claim matters: without it (every window against the initial state) other keypoints would be assigned
    free = oracle.window_search(kps, desc, geom, q, qdesc[keep], pre, u_right, qur, qerr)
    assert len(set(free["best_idx"][free["best_idx"] >= 0])) < (free["best_idx"] >= 0).sum()
    if R.frame_available():
        rnm, rwant = R.search_by_projection(kps, desc, bounds, sf, pts, qdesc, pre, u_right, th, nnratio, far, 40.0)
        assert nm == rnm and np.array_equal(np.where(got >= 0, keep[np.maximum(got, 0)], -1), rwant)
    # no windows / no keypoints
    nm0, got0 = m.SearchByProjection(kps, desc, geom, q[:0], qdesc[:0], pre)
    assert nm0 == 0 and (got0 == -1).all()


def test_search_by_bow_random_feature_vectors_and_errors(P, m, oracle):
    """FeatureVectors that do not come from a vocabulary: random node ids (many on one side only), ragged groups,
    features missing from every node (stopped words), duplicate descriptors (ties); plus the argument checks."""
    rng = np.random.default_rng(21)
    F, cap = 6, 300
    kps = np.zeros((F, cap), P.KP_DTYPE)
    kps["angle"] = rng.uniform(0, 360, (F, cap)).astype(np.float32)
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)                 # 40 "landmarks": near-duplicates across frames
    desc = base[rng.integers(0, 40, (F, cap))].copy()
    flips = rng.integers(0, 256, (F, cap, 6))
    for j in range(6):
        np.bitwise_xor.at(desc, (np.arange(F)[:, None], np.arange(cap)[None, :], flips[:, :, j] // 8),
                          (1 << (flips[:, :, j] % 8)).astype(np.uint8))
    npf = rng.integers(cap // 2, cap + 1, F).astype(np.int32)
    npf[0] = cap
    fv = {k: np.zeros((F, cap), np.uint32 if k != "fv_begin" else np.int32) for k in ("fv_nodes", "fv_begin", "fv_feats")}
    fv["fv_n"], fv["fv_total"] = np.zeros(F, np.int32), np.zeros(F, np.int32)
    packed = []
    for f in range(F):
        feats = rng.permutation(npf[f])[: int(npf[f] * 0.9)]             # 10 % of the features are in no node
        nodes = np.sort(rng.choice(60, size=int(rng.integers(5, 40)), replace=False)).astype(np.uint32)
        cuts = np.sort(rng.integers(0, len(feats) + 1, len(nodes) - 1))
        groups = [np.sort(g).astype(np.uint32) for g in np.split(feats, cuts)]
        keep = [i for i, g in enumerate(groups) if len(g)]                # std::map holds no empty vectors
        nodes, groups = nodes[keep], [groups[i] for i in keep]
        nd, bg, ft = oracle.pack_feature_vector(nodes, groups)
        packed.append((nd, bg, ft))
        fv["fv_nodes"][f, :len(nd)], fv["fv_begin"][f, :len(nd)], fv["fv_feats"][f, :len(ft)] = nd, bg[:len(nd)], ft
        fv["fv_n"][f], fv["fv_total"][f] = len(nd), len(ft)
    has_point = (rng.random((F, cap)) < 0.7).astype(np.uint8)
    pairs = np.array([(a, b) for a in range(F) for b in range(F)], np.int32)
    for ratio, ori in [(0.9, True), (0.6, False)]:
        nm, match = m.SearchByBoW(kps, desc, npf, fv, pairs, has_point, ratio, ori)
        nm2, match2 = m.SearchByBoWKeyFrames(kps, desc, npf, fv, pairs, has_point, ratio, ori)
        for p, (a, b) in enumerate(pairs):
            wnm, want = oracle.search_by_bow(kps[a, :npf[a]], desc[a, :npf[a]], has_point[a, :npf[a]], packed[a],
                                             kps[b, :npf[b]], desc[b, :npf[b]], packed[b], ratio, ori)
            assert nm[p] == wnm and np.array_equal(match[p, :npf[b]], want), ("frame form", p)
            wnm, want = oracle.search_by_bow_kf(kps[a, :npf[a]], desc[a, :npf[a]], has_point[a, :npf[a]], packed[a],
                                                kps[b, :npf[b]], desc[b, :npf[b]], has_point[b, :npf[b]], packed[b], ratio, ori)
            assert nm2[p] == wnm and np.array_equal(match2[p, :npf[a]], want), ("key-frame form", p)
        assert nm.sum() > 200
    with pytest.raises(P.OrbxError):                                      # a pair outside the pool
        m.SearchByBoW(kps, desc, npf, fv, np.array([[0, F]], np.int32))
    big = 2049
    with pytest.raises(P.OrbxError):                                      # more features per frame than a CTA handles
        m.SearchByBoW(np.zeros((1, big), P.KP_DTYPE), np.zeros((1, big, 32), np.uint8), None,
                      {"fv_nodes": np.zeros((1, big), np.uint32), "fv_begin": np.zeros((1, big), np.int32),
                       "fv_feats": np.zeros((1, big), np.uint32), "fv_n": np.zeros(1, np.int32), "fv_total": np.zeros(1, np.int32)},
                      np.array([[0, 0]], np.int32))


@pytest.mark.parametrize("seed,th,t_lw_z,mono,stereo,ori", [(1, 7.0, 0.0, False, True, True), (2, 15.0, 0.5, False, True, True),
                                                            (3, 15.0, -0.6, False, False, True), (4, 7.0, 0.5, True, False, False)])
def test_search_by_projection_last_frame(P, m, oracle, seed, th, t_lw_z, mono, stereo, ori):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (orb_matcher.cc:1518-1728; TrackWithMotionModel)
    after the projection: windows + greedy claim + rotation histogram on the device vs the oracle and the reference's lines."""
    from oracle import ref as R
    from test_oracle_vs_ref_frame import last_frame_case, last_frame_windows, W, H, BF_LAST, MB_LAST, CAM4
    kps, desc, last, ldesc, has_point, outlier, world, t_cw, src, rng = last_frame_case(oracle, seed, n_last=900)
    bounds = (0.0, float(W), 0.0, float(H))
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    sf = oracle.Extractor(1000).tables()["scale"]
    t_lw = np.array([0.0, 0.0, t_lw_z], np.float32)
    pre = (rng.random(len(kps)) < 0.15).astype(np.uint8)
    u_right = np.where(rng.random(len(kps)) < 0.6, kps["x"] - rng.uniform(2, 40, len(kps)), -1.0).astype(np.float32) if stereo else None
    keep, q, qur = last_frame_windows(oracle, sf, last, has_point, outlier, world, t_cw, t_lw, th, mono, bounds)
    args = (kps, desc, geom, q, ldesc[keep], last["angle"][keep], pre, u_right, qur if stereo else None, q["r"] if stereo else None, 100, ori)
    nm, got = m.SearchByProjectionLast(*args)
    wnm, want = oracle.search_by_projection_last(*args)
    assert nm == wnm and np.array_equal(got, want) and wnm > 150
    if R.frame_available():
        rnm, rwant = R.search_by_projection_last(kps, desc, bounds, sf, BF_LAST, MB_LAST, CAM4, t_cw, t_lw, last, has_point, outlier,
                                                 world, ldesc, pre, u_right, th, mono, ori)
        assert nm == rnm and np.array_equal(np.where(got >= 0, keep[np.maximum(got, 0)], -1), rwant)
    nm0, got0 = m.SearchByProjectionLast(kps, desc, geom, q[:0], ldesc[:0], last["angle"][:0], pre)
    assert nm0 == 0 and (got0 == -1).all()


def test_search_for_triangulation(P, m, oracle):
    """ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040; LocalMapping::CreateNewMapPoints) batched over pairs of a
    key-frame pool: bag-of-words guided matching of features without map points, epipole and epipolar-line gates, greedy
    claim, rotation histogram -- vs the oracle (pinned on the reference's lines by tests/test_oracle_vs_ref_frame.py)."""
    from test_oracle_vs_ref_frame import CAM4
    w, h, nfeat, k, L = 640, 400, 800, 6, 4
    vp_ = oracle.synth_vocab(k, L, seed=43)
    voc, vo = P.ORBVocabulary(k, L, *vp_), oracle.Vocabulary(k, L, *vp_)
    ex = P.OrbExtractor(nfeat, 1.2, 8, 20, 7)
    views = [(1, 0, 1), (1, 12, 8), (1, 6, 9), (2, 0, 2), (2, 9, 5)]
    ext = [ex(oracle.blocks_v1(w, h, s, 0, shift_x=sh, noise_seed=ns)) for s, sh, ns in views]
    cap = max(len(e[1]) for e in ext) + 3
    F = len(views)
    kps, desc, npf = np.zeros((F, cap), P.KP_DTYPE), np.zeros((F, cap, 32), np.uint8), np.zeros(F, np.int32)
    for f, (_, kk, dd) in enumerate(ext):
        kps[f, :len(kk)], desc[f, :len(kk)], npf[f] = kk, dd, len(kk)
    rng = np.random.default_rng(8)
    has_point = (rng.random((F, cap)) < 0.4).astype(np.uint8)
    u_right = np.where(rng.random((F, cap)) < 0.5, kps["x"] - rng.uniform(1, 40, (F, cap)), -1.0).astype(np.float32)
    pairs = np.array([(0, 1), (1, 0), (0, 2), (2, 1), (3, 4), (4, 3), (0, 3), (1, 1)], np.int32)
    f12 = np.zeros((len(pairs), 3, 3), np.float32)
    f12[:, 1, 2], f12[:, 2, 1] = -np.float32(0.11) / CAM4[1], np.float32(0.11) / CAM4[1]
    f12 += rng.normal(0, 2e-9, f12.shape).astype(np.float32)
    ep = np.stack([rng.uniform(-2000, 2000, len(pairs)), rng.uniform(-500, 900, len(pairs))], 1).astype(np.float32)
    ep[2] = (330.0, 210.0)                                                 # an epipole inside the image
    t = ex.GetScaleFactors(), ex.GetScaleSigmaSquares()
    for levelsup, only_stereo, coarse, ori in [(3, False, False, True), (2, True, False, True), (3, False, True, False)]:
        fv = voc.transform_batch(desc, npf, levelsup)
        nm, match = m.SearchForTriangulation(kps, desc, npf, fv, pairs, has_point, u_right, f12, ep, t[0], t[1], only_stereo, coarse, ori)
        for p, (a, b) in enumerate(pairs):
            fva = oracle.pack_feature_vector(*vo.transform(desc[a, :npf[a]], levelsup)[2:])
            fvb = oracle.pack_feature_vector(*vo.transform(desc[b, :npf[b]], levelsup)[2:])
            wnm, want = oracle.search_for_triangulation(kps[a, :npf[a]], desc[a, :npf[a]], has_point[a, :npf[a]], u_right[a, :npf[a]], fva,
                                                        kps[b, :npf[b]], desc[b, :npf[b]], has_point[b, :npf[b]], u_right[b, :npf[b]], fvb,
                                                        f12[p], ep[p], t[0], t[1], only_stereo, coarse, ori)
            assert nm[p] == wnm and np.array_equal(match[p, :npf[a]], want), (levelsup, p)
            assert (match[p, npf[a]:] == -1).all()
        assert nm[0] > (8 if only_stereo else 25) and nm[6] < nm[0]


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False)])
def test_window_search_fuse(P, m, oracle, seed, stereo):
    """The search of ORBmatcher::Fuse (orb_matcher.cc:1130-1187) on the device vs the oracle and the reference's own loop."""
    from oracle import ref as R
    from test_oracle_vs_ref_frame import fuse_case, W, H
    kps, desc, q, qdesc, qur, lev, u_right, inv_s2 = fuse_case(oracle, seed, nq=1200)
    geom = (0.0, 0.0, np.float32(64) / np.float32(W), np.float32(48) / np.float32(H), 64, 48)
    got = m.window_search_fuse(kps, desc, geom, q, qdesc, inv_s2, u_right if stereo else None, qur if stereo else None)
    want = oracle.window_search_fuse(kps, desc, geom, q, qdesc, inv_s2, u_right if stereo else None, qur if stereo else None)
    assert got.tobytes() == want.tobytes() and (want["best_idx"] >= 0).mean() > 0.3
    if R.frame_available():
        wi, wd = R.fuse_search(kps, desc, (0.0, float(W), 0.0, float(H)), inv_s2, q["u"], q["v"], qur, q["r"], lev, qdesc,
                               u_right if stereo else None)
        assert np.array_equal(got["best_idx"], wi) and np.array_equal(got["best_dist"], wd)


def _claim_case(P, scenario, KL):
    """Hand-built frames for the two orders of events the batched greedy claim must get right (the claim walks the map points
    32 at a time; KL = length of the per-window candidate lists: 8 for the ratio form, 4 for the last-frame form).  All query
    descriptors are zero, so the distance to a keypoint is the popcount of its descriptor; all keypoints sit on level 0.
    (a) q33's whole list was claimed by the previous batch and its next-best keypoint `w` is the best of q32, one lane earlier:
        q32 must get w, q33 the keypoint after it.
    (b) q33 (blocked by q32 in round one) loses its whole list and re-scans: its next-best `y` is the best of the LATER q34,
        which must not take y first."""
    pts, pops, names = [], [], {}

    def kp(name, x, y, pop):
        names[name] = len(pts)
        pts.append((x, y))
        pops.append(pop)

    queries = []  # (u, v, r)
    if scenario == "a":
        for k in range(KL):
            kp("c%d" % k, 100 + 10 * k, 100, 1 + k)
        xw = 100 + 10 * KL
        kp("w", xw, 100, 10)
        kp("x", xw + 10, 100, 40)
        kp("x2", xw + 20, 100, 60)
        for k in range(KL - 1):
            kp("e%d" % k, xw + 30 + 10 * k, 100, 20 + k)
        for k in range(KL):                                   # previous batch: every c_k is claimed by its own tiny window
            queries.append((100 + 10 * k, 100, 3))
        queries += [(700, 400, 3)] * (32 - KL)                # empty windows fill the batch
        queries.append(((xw + xw + 30 + 10 * (KL - 2)) / 2, 100, (30 + 10 * (KL - 2)) / 2 + 4))   # q32: w .. e_last
        queries.append(((100 + xw + 20) / 2, 100, (xw + 20 - 100) / 2 + 4))                        # q33: c_0 .. x2
        expect = {("c%d" % k): k for k in range(KL)}
        expect.update({"w": 32, "x": 33})
    else:
        kp("a1", 100, 100, 1)
        for k in range(1, KL):
            kp("p%d" % k, 100 + 10 * k, 100, 1 + k)
        xy = 100 + 10 * KL
        kp("y", xy, 100, 30)
        kp("y2", xy + 10, 100, 60)
        kp("z1", xy, 112, 50)
        for k in range(1, KL):
            queries.append((100 + 10 * k, 100, 3))
        queries += [(700, 400, 3)] * (32 - (KL - 1))
        queries.append((100, 100, 3))                                                   # q32 = A: claims a1
        queries.append(((100 + xy + 10) / 2, 100, (xy + 10 - 100) / 2 + 4))             # q33 = E: a1, p*, y, y2 (and z1)
        queries.append((xy, 106, 8))                                                    # q34 = L: y and z1
        expect = {("p%d" % k): k - 1 for k in range(1, KL)}
        expect.update({"a1": 32, "y": 33, "z1": 34})
    n = len(pts)
    kps = np.zeros(n, P.KP_DTYPE)
    kps["x"], kps["y"] = [p[0] for p in pts], [p[1] for p in pts]
    kps["size"], kps["angle"], kps["octave"], kps["class_id"] = 31, 0, 0, -1
    desc = np.zeros((n, 32), np.uint8)
    for i, pop in enumerate(pops):
        bits = np.arange(pop)
        np.bitwise_or.at(desc[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    q = np.zeros(len(queries), P.WQ_DTYPE)
    q["u"], q["v"], q["r"] = [a[0] for a in queries], [a[1] for a in queries], [a[2] for a in queries]
    q["min_level"], q["max_level"] = 0, 0
    want = np.full(n, -1, np.int32)
    for name, qi in expect.items():
        want[names[name]] = qi
    return kps, desc, q, np.zeros((len(queries), 32), np.uint8), want


@pytest.mark.parametrize("scenario", ["a", "b"])
def test_projection_claim_order_when_a_list_is_exhausted(P, m, oracle, scenario):
    """The two cases the batched claim got wrong before the barrier rule (a lane that may have to scan its window again blocks
    every later lane and decides alone): hand-built, checked against the oracle AND the outcome worked out by hand."""
    geom = (0.0, 0.0, np.float32(64) / np.float32(752), np.float32(48) / np.float32(480), 64, 48)
    # ratio form, lists of eight (orb_matcher.cc:42-134)
    kps, desc, q, qdesc, want = _claim_case(P, scenario, 8)
    pre = np.zeros(len(kps), np.uint8)
    wnm, owant = oracle.search_by_projection(kps, desc, geom, q, qdesc, pre, None, None, None, 100, 0.8)
    assert np.array_equal(owant, want), "the oracle disagrees with the hand-derived outcome"
    nm, got = m.SearchByProjection(kps, desc, geom, q, qdesc, pre, None, None, None, 100, 0.8)
    assert nm == wnm and np.array_equal(got, want)
    m.set_option(P.OPT_CLAIM_SEQUENTIAL, 1)
    try:
        nm_s, got_s = m.SearchByProjection(kps, desc, geom, q, qdesc, pre, None, None, None, 100, 0.8)
    finally:
        m.set_option(P.OPT_CLAIM_SEQUENTIAL, 0)
    assert nm_s == wnm and np.array_equal(got_s, want)
    # last-frame form, lists of four (orb_matcher.cc:1518-1728), no orientation check
    kps, desc, q, qdesc, want = _claim_case(P, scenario, 4)
    pre = np.zeros(len(kps), np.uint8)
    ang = np.zeros(len(q), np.float32)
    wnm, owant = oracle.search_by_projection_last(kps, desc, geom, q, qdesc, ang, pre, None, None, None, 100, False)
    assert np.array_equal(owant, want), "the oracle disagrees with the hand-derived outcome"
    nm, got = m.SearchByProjectionLast(kps, desc, geom, q, qdesc, ang, pre, None, None, None, 100, False)
    assert nm == wnm and np.array_equal(got, want)


@pytest.mark.parametrize("seed,n,nq,spread,rmax", [(1, 60, 200, 30.0, 40.0), (2, 400, 500, 120.0, 25.0), (3, 1500, 1000, 700.0, 18.0),
                                                   (4, 25, 96, 5.0, 1e6), (5, 300, 333, 60.0, 12.0)])
def test_projection_claim_crowded_windows(P, m, oracle, seed, n, nq, spread, rmax):
    """The batched claim lets map points with disjoint windows decide together; here the windows crowd into a small part of
    the image (spread = side of the square the keypoints and projections fall in), so most of a batch overlaps, lists run
    dry and windows are scanned again -- every order-of-events case at once.  Few distinct descriptor values (ties), mixed
    levels, level ranges of every kind, some huge / empty windows.  Both forms against the oracle and the sequential kernel."""
    rng = np.random.default_rng(seed)
    geom = (0.0, 0.0, np.float32(64) / np.float32(752), np.float32(48) / np.float32(480), 64, 48)
    kps = np.zeros(n, P.KP_DTYPE)
    kps["x"] = (100 + rng.random(n) * spread).astype(np.float32)
    kps["y"] = (80 + rng.random(n) * min(spread, 390.0)).astype(np.float32)
    kps["octave"] = rng.integers(0, 4, n)
    kps["angle"] = (rng.random(n) * 360).astype(np.float32)
    kps["size"], kps["class_id"] = 31, -1
    words = rng.integers(0, 256, (6, 32)).astype(np.uint8)
    desc = words[rng.integers(0, 6, n)].copy()
    desc[:, 0] ^= rng.integers(0, 4, n).astype(np.uint8)
    q = np.zeros(nq, P.WQ_DTYPE)
    q["u"] = (100 + rng.random(nq) * spread).astype(np.float32)
    q["v"] = (80 + rng.random(nq) * min(spread, 390.0)).astype(np.float32)
    q["r"] = np.minimum(2 + rng.random(nq) * 38, rmax).astype(np.float32) if rmax < 1e5 else np.float32(rmax)
    lo = rng.integers(-1, 3, nq)
    q["min_level"], q["max_level"] = lo, np.where(rng.random(nq) < 0.2, -1, lo + rng.integers(0, 3, nq))
    q["r"][rng.random(nq) < 0.05] = 0.0                                         # empty windows in between
    qdesc = words[rng.integers(0, 6, nq)].copy()
    qdesc[:, 1] ^= rng.integers(0, 4, nq).astype(np.uint8)
    pre = (rng.random(n) < 0.1).astype(np.uint8)
    for th_high, ratio in ((100, 0.8), (256, 1.0), (40, 0.6)):
        wnm, want = oracle.search_by_projection(kps, desc, geom, q, qdesc, pre, None, None, None, th_high, ratio)
        nm, got = m.SearchByProjection(kps, desc, geom, q, qdesc, pre, None, None, None, th_high, ratio)
        assert nm == wnm and np.array_equal(got, want), (th_high, ratio)
        m.set_option(P.OPT_CLAIM_SEQUENTIAL, 1)
        try:
            nm_s, got_s = m.SearchByProjection(kps, desc, geom, q, qdesc, pre, None, None, None, th_high, ratio)
        finally:
            m.set_option(P.OPT_CLAIM_SEQUENTIAL, 0)
        assert nm_s == wnm and np.array_equal(got_s, want)
    ang = (rng.random(nq) * 360).astype(np.float32)
    for ori in (False, True):
        wnm, want = oracle.search_by_projection_last(kps, desc, geom, q, qdesc, ang, pre, None, None, None, 100, ori)
        nm, got = m.SearchByProjectionLast(kps, desc, geom, q, qdesc, ang, pre, None, None, None, 100, ori)
        assert nm == wnm and np.array_equal(got, want), ori
