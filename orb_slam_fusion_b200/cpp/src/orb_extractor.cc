// ORB_SLAM_FUSION::OrbExtractor over the B200 C ABI (include/orbx.h).  Replaces the reference's
// src/cam/orb_feature/orb_extractor.cc; see the header for what stays source-compatible.
#include "cam/orb_feature/orb_extractor.h"

#include <cassert>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orbx.h"

namespace ORB_SLAM_FUSION {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint must be the 28-byte record the kernels write");

OrbExtractor::OrbExtractor(int num_feats, float scale_factor, int num_levs, int ini_th_fast, int min_th_fast)
    : handle_(nullptr),
      download_pyramid_(true),
      num_feats_(num_feats),
      scale_factor_(scale_factor),
      num_levs_(num_levs),
      ini_th_fast_(ini_th_fast),
      min_th_fast_(min_th_fast) {
  const orbx_params p = {num_feats, scale_factor, num_levs, ini_th_fast, min_th_fast};
  const char* dev = std::getenv("ORBX_DEVICE");
  const int rc = orbx_create(&p, dev ? std::atoi(dev) : 0, 1, &handle_);
  if (rc != ORBX_OK) throw std::runtime_error("OrbExtractor: orbx_create failed with code " + std::to_string(rc));
  scale_factors_.resize(num_levs_);
  inv_scale_factors_.resize(num_levs_);
  lev_sigma_2_.resize(num_levs_);
  inv_lev_sigma_2_.resize(num_levs_);
  num_feats_per_lev_.resize(num_levs_);
  orbx_tables(handle_, scale_factors_.data(), inv_scale_factors_.data(), lev_sigma_2_.data(), inv_lev_sigma_2_.data(),
              num_feats_per_lev_.data());
  img_pyramid_.resize(num_levs_);
}

OrbExtractor::~OrbExtractor() { orbx_destroy(handle_); }

void OrbExtractor::DownloadPyramid() {
  for (int lev = 0; lev < num_levs_; ++lev) {
    int w = 0, h = 0;
    if (orbx_pyramid_level(handle_, lev, nullptr, 0, &w, &h) != ORBX_OK) throw std::runtime_error(orbx_last_error(handle_));
    cv::Mat temp(h + 2 * ORBX_EDGE, w + 2 * ORBX_EDGE, CV_8UC1);
    if (orbx_pyramid_level(handle_, lev, temp.data, temp.step, &w, &h) != ORBX_OK)
      throw std::runtime_error(orbx_last_error(handle_));
    img_pyramid_[lev] = temp(cv::Rect(ORBX_EDGE, ORBX_EDGE, w, h));  // same ROI-of-bordered-Mat as :1101-1102
  }
}

void OrbExtractor::ComputePyramid(cv::Mat img) {
  if (orbx_compute_pyramid(handle_, img.data, img.cols, img.rows, img.step) != ORBX_OK)
    throw std::runtime_error(orbx_last_error(handle_));
  DownloadPyramid();
}

int OrbExtractor::operator()(cv::InputArray img, cv::InputArray /*msk*/, std::vector<cv::KeyPoint>& kps,
                             cv::OutputArray descs, std::vector<int>& lapping_areas) {
  if (img.empty()) return -1;  // :1016
  cv::Mat image = img.getMat();
  assert(image.type() == CV_8UC1);  // :1019
  int cap = orbx_max_keypoints(handle_), n = 0, n_mono = 0;
  cv::Mat desc_buf;
  int rc = ORBX_OK;
  for (int attempt = 0; attempt < 2; ++attempt) {
    kps.resize(cap);
    desc_buf.create(cap, 32, CV_8U);
    rc = orbx_extract(handle_, image.data, image.cols, image.rows, image.step, lapping_areas[0], lapping_areas[1],
                                reinterpret_cast<orbx_kp*>(kps.data()), desc_buf.data, cap, &n, &n_mono);
    if (rc == ORBX_E_CAP) { cap = n; continue; }  // capacity grows once the image geometry is known
    break;
  }
  if (rc != ORBX_OK) throw std::runtime_error(orbx_last_error(handle_));  // includes a second ORBX_E_CAP
  kps.resize(n);
  if (n == 0) {
    descs.release();  // :1033-1034
  } else {
    descs.create(n, 32, CV_8U);
    cv::Mat out = descs.getMat();
    for (int i = 0; i < n; ++i) std::memcpy(out.ptr(i), desc_buf.ptr(i), 32);
  }
  if (download_pyramid_) DownloadPyramid();
  return n_mono;
}

}  // namespace ORB_SLAM_FUSION
