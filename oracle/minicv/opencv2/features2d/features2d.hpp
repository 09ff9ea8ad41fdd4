// oracle/minicv -- TEST INFRASTRUCTURE, not product code.  See core/core.hpp.
#ifndef MINICV_FEATURES2D_HPP
#define MINICV_FEATURES2D_HPP
#include <algorithm>
#include <vector>
#include "../core/core.hpp"

namespace cv {

// cv::FAST(image, keypoints, threshold, nonmaxSuppression=true) -- TYPE_9_16
// (orb_extractor.cc:783,800).
inline void FAST(InputArray img_, std::vector<KeyPoint>& kps, int threshold, bool nms = true) {
  assert(nms);
  Mat img = img_.getMat();
  kps.clear();
  if (img.cols < 7 || img.rows < 7) return;
  std::vector<int> xyr((size_t)img.cols * img.rows * 3 / 2 + 3);
  const int n = cvp_fast9_nms_u8(img.data, img.cols, img.rows, img.step, threshold, xyr.data(),
                                 (int)(xyr.size() / 3));
  kps.reserve(n);
  for (int i = 0; i < n; i++)
    kps.push_back(KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2]));
}

// cv::DMatch as returned by BFMatcher::knnMatch (frame.cc:1154-1162); used by the C++ facade tests.
struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

// Only referenced by the reference's dead ComputeKeyPointsOld (orb_extractor.cc:977,992).
struct KeyPointsFilter {
  static void retainBest(std::vector<KeyPoint>& kps, int n) {
    if (n >= 0 && (int)kps.size() > n) {
      std::stable_sort(kps.begin(), kps.end(),
                       [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
      kps.resize(n);
    }
  }
};

}  // namespace cv
#endif
