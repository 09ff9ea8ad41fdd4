#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ (run in the development container).

The vectors are produced WITHOUT the C oracle's own primitives, so that they pin it:
  * image primitives come from the real OpenCV in this image (cv2 4.13.0: resize,
    copyMakeBorder, FastFeatureDetector, GaussianBlur, fastAtan2, BFMatcher);
  * DistributeOctTree and DescriptorDistance come from the reference's own code, compiled
    unmodified (oracle/_ref/liborb_ref.so: orb_extractor.cc:476-742, orb_matcher.cc:1877-1891);
  * the glue (grid loop :744-825, IC_Angle :76-100, rBRIEF steering :102-146, assembly
    :1030-1090) is restated here in numpy, float32 step by step, libm cosf/sinf via ctypes.
Inputs are the deterministic generators of SURVEY.md 8(d), re-implemented here in numpy.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import ctypes
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402  (reference code on the mini-cv shim: octree + hamming only)

cv2.setNumThreads(1)
HERE = os.path.dirname(os.path.abspath(__file__))
libm = ctypes.CDLL("libm.so.6")
libm.cosf.restype = libm.sinf.restype = ctypes.c_float
libm.cosf.argtypes = libm.sinf.argtypes = [ctypes.c_float]
F = np.float32
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


# ------------------------------------------------------------------ synthetic inputs
def sm64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def blocks_v1(w, h, seed=1, frame=0, shift_x=0, noise_seed=None):
    noise_seed = seed if noise_seed is None else noise_seed
    base = sm64(np.uint64((seed << 32) ^ frame))
    nbase = sm64(np.uint64((noise_seed << 32) ^ frame))
    xs = np.arange(w, dtype=np.int64)
    img = np.tile((40 + (160 * xs) // (w - 1)).astype(np.int32), (h, 1))
    for k in range((w * h) // 900):
        r = [int(sm64(base ^ np.uint64(5 * k + j))) for j in range(1, 6)]
        x0, y0 = r[0] % w - shift_x, r[1] % h
        rw, rh, v = 8 + r[2] % 82, 8 + r[3] % 82, r[4] % 256
        img[y0:min(y0 + rh, h), max(x0, 0):max(min(x0 + rw, w), 0)] = v
    yy, xx = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), indexing="ij")
    nz = (sm64(nbase ^ np.uint64(0xABCDEF) ^ (yy << np.uint64(20)) ^ xx) % np.uint64(7)).astype(np.int32) - 3
    return np.clip(img + nz, 0, 255).astype(np.uint8)


def uniform_v1(w, h, seed=1, frame=0):
    base = sm64(np.uint64((seed << 32) ^ frame))
    yy, xx = np.meshgrid(np.arange(h, dtype=np.uint64), np.arange(w, dtype=np.uint64), indexing="ij")
    return (sm64(base ^ (yy << np.uint64(20)) ^ xx) & np.uint64(255)).astype(np.uint8)


def synth_descriptors(first, n, seed):
    i = np.arange(4 * first, 4 * (first + n), dtype=np.uint64)
    return sm64(np.uint64(seed) ^ i).astype("<u8").view(np.uint8).reshape(n, 32)


# ------------------------------------------------------------------ pipeline on real OpenCV
PATTERN = np.array([int(v) for v in open(os.path.join(ROOT, "include", "orb_pattern31.inc")).read()
                    .split("*/")[1].replace("\n", "").split(",") if v.strip()], np.int32).reshape(512, 2)
UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def rint(v):
    return int(np.rint(F(v)))


class CvPipeline:
    """orb_extractor.cc restated on cv2 primitives (tier A of SURVEY.md section 7)."""

    def __init__(self, num_feats, scale_factor=1.2, num_levs=8, ini_th=20, min_th=7):
        self.nf, self.L, self.ini, self.min = num_feats, num_levs, ini_th, min_th
        sfd = float(F(scale_factor))
        self.scale = [F(1.0)]
        for _ in range(1, num_levs):
            self.scale.append(F(float(self.scale[-1]) * sfd))
        self.inv_scale = [F(1.0) / s for s in self.scale]
        factor = F(1.0 / sfd)
        per = F(F(num_feats) * (F(1) - factor)) / (F(1) - F(float(factor) ** num_levs))
        self.quota, tot = [], 0
        for _ in range(num_levs - 1):
            self.quota.append(rint(per))
            tot += self.quota[-1]
            per = F(per * factor)
        self.quota.append(max(num_feats - tot, 0))
        self.refx = R.Extractor(num_feats, scale_factor, num_levs, ini_th, min_th)

    def pyramid(self, img):
        h, w = img.shape
        lv = []
        for l in range(self.L):
            sz = (rint(F(w) * self.inv_scale[l]), rint(F(h) * self.inv_scale[l]))
            lv.append(img.copy() if l == 0 else cv2.resize(lv[-1], sz, interpolation=cv2.INTER_LINEAR))
        return lv

    def fast_grid(self, lvl):
        h, w = lvl.shape
        minb, maxbx, maxby = 16, w - 16, h - 16
        width, height = F(maxbx - minb), F(maxby - minb)
        ncols, nrows = int(width / F(35)), int(height / F(35))
        wcell, hcell = int(np.ceil(width / F(ncols))), int(np.ceil(height / F(nrows)))
        out = []
        for i in range(nrows):
            y0 = minb + i * hcell
            y1 = min(y0 + hcell + 6, maxby)
            if y0 >= maxby - 3:
                continue
            for j in range(ncols):
                x0 = minb + j * wcell
                x1 = min(x0 + wcell + 6, maxbx)
                if x0 >= maxbx - 3:
                    continue
                roi = lvl[y0:y1, x0:x1]  # a view, like rowRange().colRange()
                kp = []
                for th in (self.ini, self.min):
                    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                    kp = det.detect(roi)
                    if kp:
                        break
                out += [(int(p.pt[0]) + j * wcell, int(p.pt[1]) + i * hcell, int(p.response)) for p in kp]
        return np.array(out, np.int32).reshape(-1, 3)

    @staticmethod
    def ic_angle(lvl, cx, cy):
        m01 = m10 = 0
        for v in range(-15, 16):
            d = UMAX[abs(v)]
            row = lvl[cy + v, cx - d:cx + d + 1].astype(np.int64)
            m10 += int((np.arange(-d, d + 1) * row).sum())
            m01 += v * int(row.sum())
        return F(cv2.fastAtan2(float(F(m01)), float(F(m10))))

    @staticmethod
    def rbrief(blur, cx, cy, angle):
        ang = F(angle) * F(np.pi / F(180.0))  # factorPI = (float)(CV_PI/180.f)
        a, b = F(libm.cosf(float(ang))), F(libm.sinf(float(ang)))
        px, py = PATTERN[:, 0].astype(F), PATTERN[:, 1].astype(F)
        col = np.rint(px * a - py * b).astype(np.int32)  # float32 mul, float32 sub, half-even
        row = np.rint(px * b + py * a).astype(np.int32)
        v = blur[cy + row, cx + col].astype(np.int32)
        bits = (v[0::2] < v[1::2]).astype(np.uint8).reshape(32, 8)
        return (bits << np.arange(8, dtype=np.uint8)).sum(1).astype(np.uint8)

    def __call__(self, img, lapping=(0, 0), stages=None):
        lv = self.pyramid(img)
        per_level = []
        for l, lvl in enumerate(lv):
            cand = self.fast_grid(lvl)
            sel = self.refx.octree(cand, lvl.shape[1], lvl.shape[0], self.quota[l], l) if len(cand) else cand
            ks = np.zeros(len(sel), KP_DTYPE)
            ks["x"], ks["y"], ks["response"] = sel[:, 0] + 16, sel[:, 1] + 16, sel[:, 2]
            ks["size"], ks["octave"], ks["class_id"] = F(int(F(31) * self.scale[l])), l, -1
            for k in ks:
                k["angle"] = self.ic_angle(lvl, rint(k["x"]), rint(k["y"]))
            per_level.append((cand, ks))
            if stages is not None:
                stages.append(dict(level=lvl, cand=cand, sel=ks.copy()))
        n = sum(len(k) for _, k in per_level)
        kps, desc = np.zeros(n, KP_DTYPE), np.zeros((n, 32), np.uint8)
        mono, stereo = 0, n - 1
        for l, (_, ks) in enumerate(per_level):
            if not len(ks):
                continue
            blur = cv2.GaussianBlur(lv[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            if stages is not None:
                stages[l]["blur"] = blur
            for k in ks:
                d = self.rbrief(blur, rint(k["x"]), rint(k["y"]), k["angle"])
                k = k.copy()
                if l:
                    k["x"], k["y"] = F(k["x"]) * self.scale[l], F(k["y"]) * self.scale[l]
                if F(lapping[0]) <= k["x"] <= F(lapping[1]):
                    slot, stereo = stereo, stereo - 1
                else:
                    slot, mono = mono, mono + 1
                kps[slot], desc[slot] = k, d
        return mono, kps, desc


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CASES = [  # name, generator, w, h, seed, frame, num_feats, num_levs, lapping
    ("cfg1_752x480_1000", "blocks", 752, 480, 1, 0, 1000, 8, (0, 0)),
    ("cfg2_752x480_1200_f3", "blocks", 752, 480, 1, 3, 1200, 8, (0, 0)),
    ("kitti_1241x376_2000", "blocks", 1241, 376, 1, 0, 2000, 8, (0, 0)),
    ("hd_1280x720_1000", "blocks", 1280, 720, 1, 0, 1000, 8, (0, 0)),
    ("uniform_400x300_500", "uniform", 400, 300, 1, 0, 500, 8, (0, 0)),
    ("mono_lap_640x480_1000", "blocks", 640, 480, 2, 1, 1000, 8, (0, 1000)),
    ("fisheye_lap_640x480_800_l6", "blocks", 640, 480, 3, 2, 800, 6, (200, 420)),
]


def main():
    for name, gen, w, h, seed, frame, nf, nl, lap in CASES:
        img = blocks_v1(w, h, seed, frame) if gen == "blocks" else uniform_v1(w, h, seed, frame)
        stages = []
        mono, kps, desc = CvPipeline(nf, 1.2, nl)(img, lap, stages)
        out = dict(gen=gen, w=w, h=h, seed=seed, frame=frame, num_feats=nf, num_levs=nl,
                   lap=np.array(lap, np.int32), img_sha=sha(img), n_mono=mono, kps=kps, desc=desc,
                   level_sha=np.array([sha(s["level"]) for s in stages]),
                   blur_sha=np.array([sha(s["blur"]) if "blur" in s else "" for s in stages]),
                   cand_sha=np.array([sha(s["cand"]) for s in stages]),
                   n_cand=np.array([len(s["cand"]) for s in stages], np.int32),
                   n_sel=np.array([len(s["sel"]) for s in stages], np.int32))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "N=%d mono=%d cand=%s" % (len(kps), mono, out["n_cand"].tolist()))

    # a small frame with every stage stored in full (stage-wise parity of the CUDA kernels)
    img = blocks_v1(320, 240, 7, 0)
    stages = []
    mono, kps, desc = CvPipeline(300, 1.2, 4)(img, (0, 0), stages)
    d = dict(img=img, kps=kps, desc=desc, n_mono=mono)
    for l, s in enumerate(stages):
        d["level%d" % l], d["cand%d" % l], d["sel%d" % l] = s["level"], s["cand"], s["sel"]
        d["blur%d" % l] = s["blur"]
    np.savez_compressed(os.path.join(HERE, "stages_320x240_300_l4.npz"), **d)
    print("stages_320x240 N=%d" % len(kps))

    # matching: cv2.BFMatcher knnMatch(k=2) + ratio (frame.cc:1154-1162) and the reference's
    # DescriptorDistance on seeded descriptors with planted near-duplicates
    rng = np.random.default_rng(5)
    q = synth_descriptors(0, 200, 11)
    db = synth_descriptors(0, 5000, 12).copy()
    for i in range(200):
        row = q[i].copy()
        flips = rng.choice(256, int(rng.integers(0, 40)), replace=False)
        for b in flips:
            row[b // 8] ^= np.uint8(1 << (b % 8))
        db[int(rng.integers(0, 5000))] = row
        if i % 5 == 0:  # exact duplicates -> distance ties, lowest train index must win
            db[int(rng.integers(0, 5000))] = row
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
    idx = np.array([[mm[0].trainIdx, mm[1].trainIdx] for mm in m], np.int64)
    dist = np.array([[mm[0].distance, mm[1].distance] for mm in m], np.int32)
    acc = np.array([mm[0].distance < mm[1].distance * 0.7 for mm in m])
    hd = np.array([R.hamming(q[i], db[j]) for i, j in zip(range(200), idx[:, 0])], np.int32)
    assert (hd == dist[:, 0]).all()
    np.savez_compressed(os.path.join(HERE, "knn2_200x5000.npz"), q=q, db=db, idx=idx, dist=dist, accept=acc)
    print("knn2 accepted %d/200, ties %d" % (acc.sum(), int((dist[:, 0] == dist[:, 1]).sum())))


if __name__ == "__main__":
    main()
