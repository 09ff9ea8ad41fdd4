"""CPU: the product's shared kernel arithmetic (csrc/orbx_math.cuh) and the block-parallel
quadtree (csrc/octree_algo.inl, serial emulation) compiled with g++ and checked against the
oracle.  The CUDA build of the same sources is checked on the GPU by tests/test_gpu_*.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "host_emul.cc")
SO = os.path.join(HERE, "host_emul", "libhost_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "orb_slam_fusion_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC, os.path.join(CSRC, "orbx_math.cuh"), os.path.join(CSRC, "octree_algo.inl")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                               SRC, "-o", SO])
    L = C.CDLL(SO)
    L.emul_fast_atan2.restype = C.c_float
    L.emul_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.emul_fast9_score.argtypes = [C.c_void_p, C.c_long, C.c_int]
    L.emul_rbrief_offset.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.emul_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.emul_splitmix64.restype = C.c_uint64
    L.emul_splitmix64.argtypes = [C.c_uint64]
    return L


def grid_geom(w, h):
    width, height = np.float32(w - 32), np.float32(h - 32)
    ncols, nrows = int(width / np.float32(35)), int(height / np.float32(35))
    return int(np.ceil(width / np.float32(ncols))), int(np.ceil(height / np.float32(nrows))), ncols


def run_octree(emul, xyr, w, h, quota):
    xyr = np.ascontiguousarray(xyr, np.int32)
    wcell, hcell, ncols = grid_geom(w, h)
    out = np.empty((quota + 16 + 4 * 8, 3), np.int32)
    n = emul.emul_octree(xyr.ctypes.data, len(xyr), w, h, quota, wcell, hcell, ncols, out.ctypes.data, len(out))
    assert n >= 0
    return out[:n]


def test_fast_atan2(emul, oracle):
    rng = np.random.default_rng(2)
    for y, x in zip(rng.integers(-300000, 300000, 20000), rng.integers(-300000, 300000, 20000)):
        assert emul.emul_fast_atan2(float(y), float(x)) == oracle.fast_atan2(y, x)
    for y, x in [(0, 0), (0, 5), (5, 0), (-5, 0), (0, -5), (7, 7), (-7, 7)]:
        assert emul.emul_fast_atan2(float(y), float(x)) == oracle.fast_atan2(y, x)


def test_has_run9(emul):
    for m in range(1 << 16):
        bits = [(m >> k) & 1 for k in range(16)]
        want = any(all(bits[(k + j) % 16] for j in range(9)) for k in range(16))
        assert bool(emul.emul_has_run9(m)) == want


def test_fast9_score_equals_cv_fast_response(emul, oracle):
    for img in (oracle.blocks_v1(120, 90, 4, 0), oracle.uniform_v1(80, 60, 2, 0)):
        h, w = img.shape
        for th in (7, 20):
            score = np.zeros((h, w), np.int32)
            for y in range(3, h - 3):
                for x in range(3, w - 3):
                    score[y, x] = emul.emul_fast9_score(img.ctypes.data + y * w + x, w, th)
            # NMS exactly as cv::FAST, then compare with the oracle's (cv2-pinned) FAST
            got = []
            for y in range(3, h - 3):
                for x in range(3, w - 3):
                    s = score[y, x]
                    if s and s > max(score[y - 1, x - 1:x + 2].max(), score[y + 1, x - 1:x + 2].max(),
                                     score[y, x - 1], score[y, x + 1]):
                        got.append((x, y, s))
            want = oracle.fast9_nms(img, th)
            assert np.array_equal(np.array(got, np.int32).reshape(-1, 3), want)


def test_rbrief_offset(emul):
    rng = np.random.default_rng(1)
    F = np.float32
    for _ in range(2000):
        ang = F(rng.uniform(0, 2 * np.pi))
        a, b = F(np.cos(ang)), F(np.sin(ang))
        px, py = int(rng.integers(-13, 14)), int(rng.integers(-13, 14))
        row, col = C.c_int(), C.c_int()
        emul.emul_rbrief_offset(a, b, px, py, C.byref(row), C.byref(col))
        assert row.value == int(np.rint(F(F(px) * b) + F(F(py) * a)))
        assert col.value == int(np.rint(F(F(px) * a) - F(F(py) * b)))


def test_octree_on_real_candidates(emul, oracle):
    for (w, h, nf, kind, seed) in [(752, 480, 1000, "b", 1), (1241, 376, 2000, "b", 1), (1280, 720, 1000, "b", 2),
                                   (752, 480, 1000, "u", 1), (640, 480, 5000, "b", 3), (400, 300, 300, "b", 4)]:
        img = (oracle.blocks_v1 if kind == "b" else oracle.uniform_v1)(w, h, seed, 0)
        ex = oracle.Extractor(nf)
        ex(img)
        quota = ex.tables()["quota"]
        for lev in range(8):
            lvl = ex.level(lev)
            cand = ex.candidates(lev)
            sel = ex.selected(lev)
            got = run_octree(emul, cand, lvl.shape[1], lvl.shape[0], int(quota[lev]))
            assert len(got) == len(sel), (w, h, lev)
            assert np.array_equal(got[:, 0], sel["x"].astype(np.int32))
            assert np.array_equal(got[:, 1], sel["y"].astype(np.int32))
            assert np.array_equal(got[:, 2], sel["response"].astype(np.int32))
            # the candidate order must not matter (the GPU appends with atomics)
            perm = np.random.default_rng(lev).permutation(len(cand))
            assert np.array_equal(run_octree(emul, cand[perm], lvl.shape[1], lvl.shape[0], int(quota[lev])), got)


def test_octree_random_point_sets(emul, oracle):
    rng = np.random.default_rng(7)
    checked = 0
    for trial in range(400):
        w, h = int(rng.integers(120, 1400)), int(rng.integers(100, 800))
        if round((w - 32) / (h - 32)) < 1 or (w - 32) < 35 or (h - 32) < 35:
            continue
        dw, dh = w - 38, h - 38
        n = int(rng.integers(1, min(4000, dw * dh)))
        quota = int(rng.integers(1, 600))
        flat = np.sort(rng.choice(dw * dh, size=n, replace=False))
        x, y = flat % dw + 3, flat // dw + 3
        r = rng.integers(7, 30 if trial % 2 else 250, n)
        # order the points as the reference's grid loop would (cell-major, then row-major)
        wcell, hcell, ncols = grid_geom(w, h)
        key = (((y - 3) // hcell * ncols + (x - 3) // wcell) * hcell + (y - 3) % hcell) * wcell + (x - 3) % wcell
        o = np.argsort(key, kind="stable")
        xyr = np.stack([x[o], y[o], r[o]], 1).astype(np.int32)
        want = xyr[oracle.octree(xyr, w, h, quota)] + np.array([16, 16, 0], np.int32)
        got = run_octree(emul, xyr[rng.permutation(n)], w, h, quota)
        assert np.array_equal(got, want), (trial, w, h, n, quota)
        checked += 1
    assert checked > 300


def test_splitmix(emul, oracle):
    for x in (0, 1, 12345678901234567, 2**64 - 1):
        assert emul.emul_splitmix64(x) == oracle.splitmix64(x)
