// Drop-in replacement for the reference's include/cam/orb_feature/orb_extractor.h (:44-104).
//
// Same namespace, class name, constructor, operator(), getters, public img_pyramid_ and
// ComputePyramid, so src/map/frame.cc (:170-176, 472-475, 834, 913-933), src/tracking.cc
// (:195-204, 804-812) and include/map/keyframe.h keep compiling unchanged.  All the work is done by
// the sm_100a kernels behind the C ABI of include/orbx.h; this class only owns the handle and
// converts between OpenCV containers and plain buffers.  There is no CPU fallback: construction
// throws std::runtime_error when no CUDA device / library is available.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <opencv2/opencv.hpp>
#include <vector>

struct orbx_extractor;  // include/orbx.h

namespace ORB_SLAM_FUSION {

class OrbExtractor {
 public:
  enum { kHarrisScore = 0, kFastScore = 1 };

  // Reads ORBX_DEVICE (CUDA ordinal, default 0) from the environment: an additive option, the five
  // reference arguments are unchanged (settings/EuRoC.yaml:85-98 via src/tracking.cc:189-204).
  OrbExtractor(int num_feats, float scale_factor, int num_levs, int ini_th_fast, int min_th_fast);
  ~OrbExtractor();
  OrbExtractor(const OrbExtractor&) = delete;
  OrbExtractor& operator=(const OrbExtractor&) = delete;

  // Compute the ORB features and descriptors on an image (orb_extractor.cc:1011-1091).
  // Mask is ignored, as in the reference.  Returns the number of non-lapping keypoints, -1 for an
  // empty image.
  int operator()(cv::InputArray img, cv::InputArray msk, std::vector<cv::KeyPoint>& kps, cv::OutputArray descs,
                 std::vector<int>& lapping_areas);

  int inline GetLevels() { return num_levs_; }
  float inline GetScaleFactor() { return scale_factor_; }
  std::vector<float> inline GetScaleFactors() { return scale_factors_; }
  std::vector<float> inline GetInverseScaleFactors() { return inv_scale_factors_; }
  std::vector<float> inline GetScaleSigmaSquares() { return lev_sigma_2_; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return inv_lev_sigma_2_; }

  // Level images of the last call, each a view into a buffer that carries the 19-px
  // BORDER_REFLECT_101 frame (orb_extractor.cc:1109-1114); read by Frame::ComputeStereoMatches.
  std::vector<cv::Mat> img_pyramid_;

  void ComputePyramid(cv::Mat img);

  // Additive: skip the device->host copy of the pyramid after operator() when no caller reads
  // img_pyramid_ (monocular / RGB-D tracking never does).
  void SetPyramidDownload(bool on) { download_pyramid_ = on; }
  orbx_extractor* handle() { return handle_; }

 protected:
  void DownloadPyramid();

  orbx_extractor* handle_;
  bool download_pyramid_;
  int num_feats_;
  double scale_factor_;
  int num_levs_;
  int ini_th_fast_;
  int min_th_fast_;
  std::vector<int> num_feats_per_lev_;
  std::vector<float> scale_factors_;
  std::vector<float> inv_scale_factors_;
  std::vector<float> lev_sigma_2_;
  std::vector<float> inv_lev_sigma_2_;
};

}  // namespace ORB_SLAM_FUSION

#endif
