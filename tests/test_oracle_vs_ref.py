"""CPU: the oracle restatement against the reference's own orb_extractor.cc compiled unmodified
on the mini-cv shim (oracle/_ref/liborb_ref.so; built by `make -C oracle ref` where
/root/reference exists, shipped as a binary elsewhere)."""
import numpy as np
import pytest

from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")


@pytest.mark.parametrize("w,h,nf,kind,seed,frame", [
    (752, 480, 1000, "b", 1, 0), (752, 480, 1200, "b", 1, 5), (1241, 376, 2000, "b", 1, 1),
    (640, 480, 5000, "b", 2, 1), (400, 300, 500, "u", 4, 0), (320, 240, 1000, "b", 6, 2)])
def test_full_extract_identical(oracle, w, h, nf, kind, seed, frame):
    img = (oracle.blocks_v1 if kind == "b" else oracle.uniform_v1)(w, h, seed, frame)
    eo, er = oracle.Extractor(nf), R.Extractor(nf)
    for lap in ((0, 0), (0, 1000), (w // 3, w // 2)):
        nm, k, d = eo(img, lap)
        rc, k2, d2 = er(img, lap)
        assert nm == rc and k.tobytes() == k2.tobytes() and np.array_equal(d, d2)
    for lev in range(8):
        assert np.array_equal(eo.level(lev, True), er.level(lev, True))   # incl. the 19-px border
    t, t2 = eo.tables(), er.tables()
    for key in t2:
        assert np.array_equal(t[key], t2[key])


def test_octree_identical_on_random_point_sets(oracle):
    rng = np.random.default_rng(3)
    er = R.Extractor(1000)
    for trial in range(60):
        w, h = int(rng.integers(120, 900)), int(rng.integers(100, 500))
        if round((w - 32) / (h - 32)) < 1:
            continue
        n = int(rng.integers(1, 3000))
        quota = int(rng.integers(1, 500))
        flat = rng.choice((w - 38) * (h - 38), size=min(n, (w - 38) * (h - 38)), replace=False)
        flat.sort()
        xyr = np.stack([flat % (w - 38) + 3, flat // (w - 38) + 3,
                        rng.integers(7, 40 if trial % 2 else 200, len(flat))], 1).astype(np.int32)
        got = xyr[oracle.octree(xyr, w, h, quota)]
        want = er.octree(xyr, w, h, quota)
        assert np.array_equal(got, want), (trial, w, h, n, quota)


def test_empty_image(oracle):
    assert R.Extractor(1000)(np.empty((0, 0), np.uint8))[0] == -1


def test_descriptor_distance(oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8)
        b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert R.hamming(a, b) == oracle.hamming(a, b) == int(np.unpackbits(a ^ b).sum())
