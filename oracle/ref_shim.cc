// oracle/ref_shim.cc -- TEST INFRASTRUCTURE, not product code.
//
// C entry points over the reference's own OrbExtractor (compiled unmodified from
// /root/reference/src/cam/orb_feature/orb_extractor.cc on the mini-cv shim) and its
// DescriptorDistance (orb_matcher.cc:1877-1891, spliced at build time into
// _ref/descriptor_distance.inc).  Built only by `make -C oracle ref`; used to validate the
// oracle restatement and as the "reference" CPU baseline of bench.py.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "cam/orb_feature/orb_extractor.h"

namespace ORB_SLAM_FUSION {
class ORBmatcher {
 public:
  static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
};
#include "descriptor_distance.inc"

// exposes the protected DistributeOctTree (orb_extractor.cc:542-742)
struct OrbExtractorOpen : public OrbExtractor {
  using OrbExtractor::OrbExtractor;
  using OrbExtractor::DistributeOctTree;
};
}  // namespace ORB_SLAM_FUSION

using ORB_SLAM_FUSION::OrbExtractorOpen;

extern "C" {

void* ref_create(int num_feats, float scale_factor, int num_levs, int ini_th, int min_th) {
  return new OrbExtractorOpen(num_feats, scale_factor, num_levs, ini_th, min_th);
}
void ref_destroy(void* h) { delete (OrbExtractorOpen*)h; }

// kps: 28-byte cv::KeyPoint records; desc: n x 32.  Returns operator()'s return value
// (mono count, -1 for an empty image); *n = number of keypoints.
int ref_extract(void* h, const uint8_t* img, int w, int h_, size_t stride, int lap0, int lap1,
                void* kps, uint8_t* desc, int cap, int* n) {
  OrbExtractorOpen* e = (OrbExtractorOpen*)h;
  cv::Mat im = (img && w > 0 && h_ > 0) ? cv::Mat(h_, w, CV_8UC1, (void*)img, stride) : cv::Mat();
  std::vector<cv::KeyPoint> k;
  cv::Mat d;
  std::vector<int> lap = {lap0, lap1};
  const int rc = (*e)(im, cv::Mat(), k, d, lap);
  *n = (int)k.size();
  if ((int)k.size() <= cap) {
    if (!k.empty()) std::memcpy(kps, k.data(), k.size() * sizeof(cv::KeyPoint));
    for (int i = 0; i < d.rows; i++) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
  }
  return rc;
}

// level lev of img_pyramid_ after ref_extract/ref_pyramid, copied with its 19-px border
int ref_level(void* h, int lev, uint8_t* dst, size_t dst_stride, int* w, int* h_) {
  OrbExtractorOpen* e = (OrbExtractorOpen*)h;
  const cv::Mat& m = e->img_pyramid_[lev];
  *w = m.cols;
  *h_ = m.rows;
  if (dst)
    for (int y = -19; y < m.rows + 19; y++)
      std::memcpy(dst + (size_t)(y + 19) * dst_stride, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - 19, (size_t)m.cols + 38);
  return 0;
}

void ref_pyramid(void* h, const uint8_t* img, int w, int h_, size_t stride) {
  ((OrbExtractorOpen*)h)->ComputePyramid(cv::Mat(h_, w, CV_8UC1, (void*)img, stride));
}

// DistributeOctTree on (x, y, response) int triples; writes the selected (x, y, response).
int ref_octree(void* h, const int* xyr, int n, int min_x, int max_x, int min_y, int max_y, int quota,
               int lev, int* out_xyr, int cap) {
  std::vector<cv::KeyPoint> in;
  in.reserve(n);
  for (int i = 0; i < n; i++)
    in.push_back(cv::KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2]));
  std::vector<cv::KeyPoint> out =
      ((OrbExtractorOpen*)h)->DistributeOctTree(in, min_x, max_x, min_y, max_y, quota, lev);
  for (size_t i = 0; i < out.size() && (int)i < cap; i++) {
    out_xyr[3 * i] = (int)out[i].pt.x;
    out_xyr[3 * i + 1] = (int)out[i].pt.y;
    out_xyr[3 * i + 2] = (int)out[i].response;
  }
  return (int)out.size();
}

void ref_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  OrbExtractorOpen* e = (OrbExtractorOpen*)h;
  const int L = e->GetLevels();
  std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(),
                     c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
  std::memcpy(scale, a.data(), 4 * L);
  std::memcpy(inv_scale, b.data(), 4 * L);
  std::memcpy(sigma2, c.data(), 4 * L);
  std::memcpy(inv_sigma2, d.data(), 4 * L);
}

int ref_hamming(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8U, (void*)a), mb(1, 32, CV_8U, (void*)b);
  return ORB_SLAM_FUSION::ORBmatcher::DescriptorDistance(ma, mb);
}

// Brute-force top-2 with the reference's DescriptorDistance as the inner kernel (the CPU
// baseline for the matching metric); queries split over nthreads.
void ref_knn2(const uint8_t* q, int nq, const uint8_t* d, long long nd, long long* idx, int* dist,
              int nthreads) {
  auto work = [&](int q0, int q1) {
    for (int qi = q0; qi < q1; qi++) {
      cv::Mat mq(1, 32, CV_8U, (void*)(q + 32 * (size_t)qi));
      int b0 = INT32_MAX, b1 = INT32_MAX;
      long long i0 = -1, i1 = -1;
      for (long long r = 0; r < nd; r++) {
        cv::Mat md(1, 32, CV_8U, (void*)(d + 32 * (size_t)r));
        const int dd = ORB_SLAM_FUSION::ORBmatcher::DescriptorDistance(mq, md);
        if (dd < b0) { b1 = b0; i1 = i0; b0 = dd; i0 = r; }
        else if (dd < b1) { b1 = dd; i1 = r; }
      }
      idx[2 * qi] = i0; idx[2 * qi + 1] = i1; dist[2 * qi] = b0; dist[2 * qi + 1] = b1;
    }
  };
  if (nthreads <= 1) { work(0, nq); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(work, (int)((long long)nq * t / nthreads), (int)((long long)nq * (t + 1) / nthreads));
  for (auto& t : th) t.join();
}

}  // extern "C"
