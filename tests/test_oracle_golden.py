"""CPU: the C oracle against the committed golden vectors (made from real OpenCV 4.13 + the
reference's own octree/DescriptorDistance code by tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import EXTRACT_CASES, golden_image, load_golden


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", EXTRACT_CASES)
def test_extract_matches_golden(oracle, name):
    g = load_golden(name)
    img = golden_image(g)
    ex = oracle.Extractor(int(g["num_feats"]), 1.2, int(g["num_levs"]), 20, 7)
    n_mono, kps, desc = ex(img, tuple(int(v) for v in g["lap"]))
    assert n_mono == int(g["n_mono"])
    assert kps.tobytes() == g["kps"].tobytes()          # 28-byte records, bit for bit
    assert np.array_equal(desc, g["desc"])
    for lev in range(int(g["num_levs"])):
        assert sha(ex.level(lev)) == str(g["level_sha"][lev])
        assert sha(ex.candidates(lev)) == str(g["cand_sha"][lev])
        assert len(ex.selected(lev)) == int(g["n_sel"][lev])
        if str(g["blur_sha"][lev]):
            assert sha(ex.blurred(lev)) == str(g["blur_sha"][lev])


def test_stagewise_small_frame(oracle):
    g = load_golden("stages_320x240_300_l4")
    img = g["img"]
    ex = oracle.Extractor(300, 1.2, 4, 20, 7)
    n_mono, kps, desc = ex(img)
    assert kps.tobytes() == g["kps"].tobytes() and np.array_equal(desc, g["desc"])
    quota = ex.tables()["quota"]
    for lev in range(4):
        lvl = g["level%d" % lev]
        assert np.array_equal(ex.level(lev), lvl)
        cand = oracle.fast_grid(lvl)
        assert np.array_equal(cand, g["cand%d" % lev])
        sel = oracle.octree(cand, lvl.shape[1], lvl.shape[0], int(quota[lev]))
        gs = g["sel%d" % lev]
        assert np.array_equal(cand[sel, 0] + 16, gs["x"].astype(np.int32))
        assert np.array_equal(cand[sel, 1] + 16, gs["y"].astype(np.int32))
        assert np.array_equal(oracle.gauss7x7(lvl), g["blur%d" % lev])
        for k in gs[:50]:
            assert np.float32(oracle.ic_angle(lvl, int(k["x"]), int(k["y"]))) == k["angle"]


def test_knn2_matches_bfmatcher_golden(oracle):
    g = load_golden("knn2_200x5000")
    idx, dist = oracle.knn2(g["q"], g["db"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    assert np.array_equal(oracle.ratio_accept(idx, dist, 0.7), g["accept"])
    idx2, dist2 = oracle.knn2(g["q"], g["db"], nthreads=3)
    assert np.array_equal(idx, idx2) and np.array_equal(dist, dist2)


def test_hamming_is_popcount(oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    want = np.unpackbits(a ^ b, axis=1).sum(1)
    got = [oracle.hamming(x, y) for x, y in zip(a, b)]
    assert list(want) == got
    assert oracle.hamming(a[0], a[0]) == 0
    assert oracle.hamming(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_constructor_tables(oracle):
    t = oracle.Extractor(1000, 1.2, 8, 20, 7).tables()
    assert t["quota"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert np.float32(t["scale"][1]) == np.float32(1.2000000477)
    assert oracle.Extractor(2000).tables()["quota"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]


def test_empty_image_returns_minus_one(oracle):
    ex = oracle.Extractor(1000)
    assert ex(np.empty((0, 0), np.uint8))[0] == -1


def test_synth_generators_are_pinned(oracle):
    # first pixels of blocks-v1(752,480,1,0); full-image hashes are checked through golden_image()
    img = oracle.blocks_v1(752, 480, 1, 0)
    assert img[0, :8].tolist() == [43, 41, 41, 42, 43, 38, 42, 43]
    assert oracle.splitmix64(0) == 0xE220A8397B1DCDAF
    d = oracle.synth_descriptors(5, 2, 99)
    w = np.frombuffer(d.tobytes(), "<u8")
    assert int(w[0]) == oracle.splitmix64(99 ^ 20) and int(w[7]) == oracle.splitmix64(99 ^ 27)
