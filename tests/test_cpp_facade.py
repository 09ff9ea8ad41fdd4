"""The C++ drop-in classes (orb_slam_fusion_b200/cpp): they must compile against an OpenCV-shaped API
(CPU test: oracle/minicv stands in for the absent OpenCV headers) and, on the GPU, return exactly what
the reference's OrbExtractor returns to Frame (frame.cc:467-476, 834, 1154)."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "facade_test")


def build():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp")], stdout=subprocess.DEVNULL)
    return EXE


def test_facade_compiles_and_links_against_the_c_abi():
    exe = build()
    assert os.path.exists(exe)
    # the facade binds only to the exported C ABI
    syms = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    used = {l.split()[-1] for l in syms.splitlines() if " orbx_" in l or " orbm_" in l}
    assert {"orbx_create", "orbx_extract", "orbx_pyramid_level", "orbm_knn2", "orbm_stereo_rowband"} <= used
    assert {"orbv_load_text", "orbv_transform"} <= {l.split()[-1] for l in syms.splitlines() if " orbv_" in l}
    assert "orbm_search_by_bow" in used


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,nf,lap", [(752, 480, 1200, (0, 0)), (640, 480, 1000, (0, 1000)), (640, 480, 800, (150, 480))])
def test_facade_outputs_equal_oracle(oracle, tmp_path, w, h, nf, lap):
    exe = build()
    img = oracle.blocks_v1(w, h, 2, 4)
    raw, out = tmp_path / "in.raw", tmp_path / "out.bin"
    img.tofile(raw)
    vk, vL = 6, 4
    vparent, vleaf, vdesc, vweight = oracle.synth_vocab(vk, vL, seed=9)
    voc_path = tmp_path / "voc.txt"
    oracle.save_vocab_text(str(voc_path), vk, vL, vparent, vleaf, vdesc, vweight)
    subprocess.check_call([exe, str(raw), str(w), str(h), str(nf), str(lap[0]), str(lap[1]), str(out), str(voc_path)])
    buf = out.read_bytes()
    hdr = np.frombuffer(buf, np.int32, 6)
    mono, n, drows, dcols, empty_rc, levels = (int(v) for v in hdr)
    off = 24
    kps = np.frombuffer(buf, oracle.KP_DTYPE, n, off); off += 28 * n
    desc = np.frombuffer(buf, np.uint8, 32 * n, off).reshape(n, 32); off += 32 * n
    sf = np.frombuffer(buf, np.float32, levels, off); off += 4 * levels
    ref = oracle.Extractor(nf, 1.2, 8, 20, 7, trig=oracle.TRIG_CR)
    rn, rk, rd = ref(img, lap)
    assert (mono, n, drows, dcols, empty_rc, levels) == (rn, len(rk), len(rk), 32, -1, 8)
    assert kps.tobytes() == rk.tobytes() and np.array_equal(desc, rd)
    assert np.array_equal(sf, ref.tables()["scale"])
    for lev in range(8):
        lw, lh = (int(v) for v in np.frombuffer(buf, np.int32, 2, off)); off += 8
        want = ref.level(lev, with_border=True)
        assert (lh + 38, lw + 38) == want.shape
        got = np.frombuffer(buf, np.uint8, want.size, off).reshape(want.shape); off += want.size
        assert np.array_equal(got, want), lev
    nq = min(64, n)
    rec = np.frombuffer(buf, np.int32, 4 * nq, off).reshape(nq, 4); off += 16 * nq
    ri, rdist = oracle.knn2(rd[:nq], rd)
    assert np.array_equal(rec[:, 0], ri[:, 0]) and np.array_equal(rec[:, 2], ri[:, 1])
    assert np.array_equal(rec[:, 1], rdist[:, 0]) and np.array_equal(rec[:, 3], rdist[:, 1])
    bi = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    bd = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    wi, wd = oracle.stereo_rowband(rk, rd, rk, rd, sf, h, 0.0, 40.0)
    assert np.array_equal(bi, wi) and np.array_equal(bd, wd)
    # ORBmatcherGpu::SearchByProjection (orb_matcher.cc:42-134), the whole function
    nproj = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    assigned = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    from test_oracle_vs_ref_frame import projection_windows
    from oracle import ref as R
    pts = np.zeros(n, R.TRACK_POINT_DTYPE)
    pts["proj_x"], pts["proj_y"] = rk["x"] + np.float32(1), rk["y"] - np.float32(1)
    pts["view_cos"] = np.where(np.arange(n) % 2, 1.0, 0.9)
    pts["depth"], pts["level"], pts["in_view"] = 10.0, rk["octave"], np.arange(n) % 5 != 0
    keep, q = projection_windows(oracle, sf, pts, 3.0, False, 50.0)
    geom = (0.0, 0.0, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h), 64, 48)
    wnm, want = oracle.search_by_projection(rk, rd, geom, q, rd[keep], (np.arange(n) % 7 == 0).astype(np.uint8), None, None, None, 100, 0.8)
    assert nproj == wnm and np.array_equal(assigned, np.where(want >= 0, keep[np.maximum(want, 0)], -1)) and wnm > n // 2
    # ORBmatcherGpu::SearchByProjectionLastFrame (orb_matcher.cc:1518-1728)
    nlast = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    assigned_last = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    ql = np.zeros(n, oracle.WQ_DTYPE)
    ql["u"], ql["v"] = rk["x"] + np.float32(2), rk["y"] - np.float32(1)
    ql["r"] = np.float32(7.0) * sf[np.clip(rk["octave"], 0, 7)]
    ql["min_level"], ql["max_level"] = rk["octave"] - 1, rk["octave"] + 1
    wnm, want = oracle.search_by_projection_last(rk, rd, geom, ql, rd, rk["angle"], (np.arange(n) % 7 == 0).astype(np.uint8), None, None,
                                                 None, 100, True)
    assert nlast == wnm and np.array_equal(assigned_last, want) and wnm > n // 2
    # ORBVocabularyGpu::transform as Frame::ComputeBoW calls it
    words, nb, nfv = (int(v) for v in np.frombuffer(buf, np.int32, 3, off)); off += 12
    vo = oracle.Vocabulary(vk, vL, vparent, vleaf, vdesc, vweight)
    ids, vals, nodes, feats = vo.transform(rd, 4)
    assert (words, nb, nfv) == (vo.n_words, len(ids), len(nodes))
    bow = np.frombuffer(buf, np.dtype([("id", "<u4"), ("v", "<f8")]), nb, off); off += 12 * nb
    assert np.array_equal(bow["id"], ids) and bow["v"].tobytes() == np.ascontiguousarray(vals).tobytes()
    for j in range(nfv):
        node, cnt = (int(v) for v in np.frombuffer(buf, np.uint32, 2, off)); off += 8
        fl = np.frombuffer(buf, np.uint32, cnt, off); off += 4 * cnt
        assert node == nodes[j] and np.array_equal(fl, feats[j])
    assert int(np.frombuffer(buf, np.uint32, 1, off)[0]) == int(vo.features(rd[:1])[0][0]); off += 4
    # ORBmatcherGpu::SearchByBoW of the frame against itself (orb_matcher.cc:215-389)
    nmatch = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    match = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    has_point = np.ones(n, np.uint8)
    has_point[::3] = 0
    fv = oracle.pack_feature_vector(nodes, feats)
    wnm, want = oracle.search_by_bow(rk, rd, has_point, fv, rk, rd, fv, 0.7, True)
    assert nmatch == wnm and np.array_equal(match, want) and wnm > n // 3
    nmatch_kf = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    match_kf = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    wnm, want = oracle.search_by_bow_kf(rk, rd, has_point, fv, rk, rd, has_point, fv, 0.8, True)
    assert nmatch_kf == wnm and np.array_equal(match_kf, want) and wnm > n // 3
    # ORBmatcherGpu::SearchForTriangulation (orb_matcher.cc:817-1040)
    ntri = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    tri = np.frombuffer(buf, np.int32, n, off); off += 4 * n
    with_point = np.zeros(n, np.uint8)
    with_point[::3] = 1
    f12 = np.array([0, 0, 0, 0, 0, -2.4e-4, 0, 2.4e-4, 0], np.float32)
    wnm, want = oracle.search_for_triangulation(rk, rd, with_point, np.full(n, -1, np.float32), fv, rk, rd, with_point,
                                                np.full(n, -1, np.float32), fv, f12, np.array([-5000, 200], np.float32), sf,
                                                ref.tables()["sigma2"], False, False, True)
    assert ntri == wnm and np.array_equal(tri, want) and wnm > n // 3
    assert off == len(buf)
