"""CPU: oracle primitives against the real OpenCV of this image (cv2 4.13.0).  Skipped where
cv2 is not importable; the golden vectors carry the same evidence to such machines."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


@pytest.mark.parametrize("src,dst", [((752, 480), (627, 400)), ((627, 400), (522, 333)),
                                     ((1241, 376), (1034, 313)), ((252, 161), (210, 134)), ((64, 48), (97, 71))])
def test_resize_linear(oracle, src, dst):
    img = oracle.uniform_v1(src[0], src[1], 3, 1)
    assert np.array_equal(oracle.resize_linear(img, *dst), cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR))


def test_border_and_blur(oracle):
    for img in (oracle.blocks_v1(363, 231, 2, 0), oracle.uniform_v1(210, 134), oracle.uniform_v1(9, 8)):
        assert np.array_equal(oracle.border_reflect101(img, 19) if min(img.shape) > 19 else img,
                              cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101) if min(img.shape) > 19 else img)
        assert np.array_equal(oracle.gauss7x7(img),
                              cv2.GaussianBlur(img.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("th", [20, 7])
def test_fast_per_cell(oracle, th):
    img = oracle.blocks_v1(435, 278, 1, 4)
    noise = oracle.uniform_v1(96, 80, 5, 0)
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for roi in (img[16:60, 16:58], img[100:147, 200:244], img, noise, noise[:7, :9], noise[:6, :30]):
        want = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in det.detect(np.ascontiguousarray(roi))],
                        np.int32).reshape(-1, 3)
        assert np.array_equal(oracle.fast9_nms(np.ascontiguousarray(roi), th), want)


def test_fast_atan2(oracle):
    rng = np.random.default_rng(1)
    ys = rng.integers(-300000, 300000, 5000)
    xs = rng.integers(-300000, 300000, 5000)
    for y, x in list(zip(ys, xs)) + [(0, 0), (0, 5), (5, 0), (-5, 0), (0, -5), (7, 7), (-7, 7)]:
        assert np.float32(oracle.fast_atan2(y, x)) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_knn2_vs_bfmatcher(oracle):
    q = oracle.synth_descriptors(0, 64, 21)
    db = oracle.synth_descriptors(0, 700, 22).copy()
    db[100] = db[50] = q[3]                      # exact duplicates: tie on distance 0
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, db, k=2)
    idx, dist = oracle.knn2(q, db)
    assert [[a.trainIdx, b.trainIdx] for a, b in m] == idx.tolist()
    assert [[int(a.distance), int(b.distance)] for a, b in m] == dist.tolist()
    assert idx[3].tolist() == [50, 100]


def test_rbrief_matches_cv2_orb_on_float_blur(oracle):
    """Third-party known-answer check of the pattern / steering restatement:
    cv2.ORB.compute keeps the given angles and samples the same pattern with the same float32
    arithmetic, but on OpenCV's float-path blur (its in-place blur of a sub-matrix), which
    sepFilter2D reproduces; the reference's clone()+GaussianBlur takes the 8-bit fixed-point
    path instead (that one is pinned by test_border_and_blur)."""
    img = oracle.blocks_v1(400, 300, 9, 0)
    kps = []
    for x, y, r in oracle.fast_grid(img):
        x, y = int(x) + 16, int(y) + 16
        if 40 <= x < 360 and 40 <= y < 260:      # cv2.ORB drops keypoints within edgeThreshold
            kps.append(cv2.KeyPoint(float(x), float(y), 31.0, float(oracle.ic_angle(img, x, y)), float(r), 0, -1))
    kps = kps[:400]
    kps2, want = cv2.ORB_create(nlevels=1, edgeThreshold=31, patchSize=31).compute(img, kps)
    assert len(kps2) == len(kps) > 200
    k1 = cv2.getGaussianKernel(7, 2)
    blur_float = cv2.sepFilter2D(img, cv2.CV_8U, k1, k1, borderType=cv2.BORDER_REFLECT_101)
    for k, w in zip(kps2, want):
        assert np.array_equal(oracle.rbrief(blur_float, int(k.pt[0]), int(k.pt[1]), k.angle), w)
