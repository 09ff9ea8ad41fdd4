// tests/cpp/matcher_facade_harness.cc -- TEST INFRASTRUCTURE.
//
// The product's ORBmatcher class (orb_slam_fusion_b200/cpp/src/orb_matcher.cc, compiled against the stand-in
// Frame / KeyFrame / MapPoint of tests/cpp/standin) behind the SAME extern "C" entry points as
// oracle/ref_frame_shim.cc, where the reference's own spliced method bodies run: tests/test_cpp_matcher.py loads
// both libraries, feeds them identical arrays and compares what the two ORBmatcher classes did to the frames --
// through the class API, pointers and all.
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "cam/orb_feature/orb_matcher.h"

using namespace ORB_SLAM_FUSION;

namespace ORB_SLAM_FUSION {
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
std::vector<std::tuple<int, int, int> > MapPoint::log;
// the fundamental matrix is a product of Eigen matrices in the reference (pinhole_model.cc:116-119); here it is given
void FundamentalMatrix(KeyFrame *pKF1, KeyFrame *, const Eigen::Matrix3f &, const Eigen::Vector3f &, float f12[9]) {
  for (int i = 0; i < 9; i++) f12[i] = pKF1->cam_->F12_given.m[i];
}
}  // namespace ORB_SLAM_FUSION

namespace {
void set_bounds(float min_x, float max_x, float min_y, float max_y) {
  Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);  // frame.cc:214-217
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
}
void fill_kf(KeyFrame &kf, std::vector<MapPoint> &mps, const void *kps, const uint8_t *desc, int n, const uint8_t *has_point,
             const uint32_t *nodes, const int *begin, int n_nodes, const uint32_t *feats, int total) {
  kf.mvKeysUn.assign((const cv::KeyPoint *)kps, (const cv::KeyPoint *)kps + n);
  kf.mvKeys = kf.mvKeysUn;
  kf.N = n;
  kf.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)desc);
  kf.mvpMapPoints.assign(n, (MapPoint *)nullptr);
  for (int i = 0; i < n; i++)
    if (!has_point || has_point[i]) kf.mvpMapPoints[i] = &mps[i];
  for (int j = 0; j < n_nodes; j++)
    kf.mFeatVec[nodes[j]].assign(feats + begin[j], feats + (j + 1 < n_nodes ? begin[j + 1] : total));
}
}  // namespace

extern "C" {

struct reff_track_point {
  float proj_x, proj_y, proj_xr, view_cos, depth;
  int level, in_view, bad;
};

int reff_search_by_projection(const void *keys_un, const uint8_t *desc, int n, const float *u_right, float min_x, float max_x,
                              float min_y, float max_y, const float *scale, int n_levels, const reff_track_point *pts,
                              const uint8_t *pt_desc, int n_pts, const uint8_t *pre_matched, float th, float nnratio, int far_points,
                              float th_far, int *assigned) {
  Frame F;
  F.N = n;
  F.mvKeysUn.assign((const cv::KeyPoint *)keys_un, (const cv::KeyPoint *)keys_un + n);
  F.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)desc);
  F.mvuRight.assign(n, -1.0f);
  if (u_right) F.mvuRight.assign(u_right, u_right + n);
  F.mvScaleFactors.assign(scale, scale + n_levels);
  set_bounds(min_x, max_x, min_y, max_y);
  MapPoint occupied;
  occupied.nObs = 1;
  F.mvpMapPoints.assign(n, (MapPoint *)nullptr);
  for (int i = 0; i < n; i++)
    if (pre_matched && pre_matched[i]) F.mvpMapPoints[i] = &occupied;
  std::vector<MapPoint> mps(n_pts > 0 ? n_pts : 1);
  std::vector<MapPoint *> vp;
  for (int i = 0; i < n_pts; i++) {
    MapPoint &m = mps[i];
    m.mTrackProjX = pts[i].proj_x;
    m.mTrackProjY = pts[i].proj_y;
    m.mTrackProjXR = pts[i].proj_xr;
    m.mTrackViewCos = pts[i].view_cos;
    m.mTrackDepth = pts[i].depth;
    m.mnTrackScaleLevel = pts[i].level;
    m.mbTrackInView = pts[i].in_view != 0;
    m.mbBad = pts[i].bad != 0;
    m.nObs = 1;
    m.mDescriptor = cv::Mat(1, 32, CV_8U, (void *)(pt_desc + 32 * (size_t)i)).clone();
    vp.push_back(&m);
  }
  ORBmatcher matcher(nnratio, true);
  const int nm = matcher.SearchByProjection(F, vp, th, far_points != 0, th_far);
  for (int i = 0; i < n; i++) {
    MapPoint *p = F.mvpMapPoints[i];
    assigned[i] = (p && p != &occupied) ? (int)(p - mps.data()) : -1;
  }
  return nm;
}

int reff_search_by_bow(const void *kps_kf, const uint8_t *desc_kf, int n_kf, const uint8_t *has_point_kf, const uint32_t *nodes_kf,
                       const int *begin_kf, int n_nodes_kf, const uint32_t *feats_kf, int total_kf, const void *kps_f,
                       const uint8_t *desc_f, int n_f, const uint32_t *nodes_f, const int *begin_f, int n_nodes_f,
                       const uint32_t *feats_f, int total_f, float nnratio, int check_orientation, int *match_of_f) {
  KeyFrame kf;
  Frame F;
  std::vector<MapPoint> mps(n_kf > 0 ? n_kf : 1);
  fill_kf(kf, mps, kps_kf, desc_kf, n_kf, has_point_kf, nodes_kf, begin_kf, n_nodes_kf, feats_kf, total_kf);
  F.N = n_f;
  F.mvKeys.assign((const cv::KeyPoint *)kps_f, (const cv::KeyPoint *)kps_f + n_f);
  F.mDescriptors = cv::Mat(n_f, 32, CV_8U, (void *)desc_f);
  for (int j = 0; j < n_nodes_f; j++)
    F.mFeatVec[nodes_f[j]].assign(feats_f + begin_f[j], feats_f + (j + 1 < n_nodes_f ? begin_f[j + 1] : total_f));
  std::vector<MapPoint *> matches;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int nm = matcher.SearchByBoW(&kf, F, matches);
  for (int i = 0; i < n_f; i++) match_of_f[i] = matches[i] ? (int)(matches[i] - mps.data()) : -1;
  return nm;
}

int reff_search_by_bow_kf(const void *kps1, const uint8_t *desc1, int n1, const uint8_t *has_point1, const uint32_t *nodes1,
                          const int *begin1, int n_nodes1, const uint32_t *feats1, int total1, const void *kps2, const uint8_t *desc2,
                          int n2, const uint8_t *has_point2, const uint32_t *nodes2, const int *begin2, int n_nodes2,
                          const uint32_t *feats2, int total2, float nnratio, int check_orientation, int *match_of_1) {
  KeyFrame k1, k2;
  std::vector<MapPoint> m1(n1 > 0 ? n1 : 1), m2(n2 > 0 ? n2 : 1);
  fill_kf(k1, m1, kps1, desc1, n1, has_point1, nodes1, begin1, n_nodes1, feats1, total1);
  fill_kf(k2, m2, kps2, desc2, n2, has_point2, nodes2, begin2, n_nodes2, feats2, total2);
  std::vector<MapPoint *> matches;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int nm = matcher.SearchByBoW(&k1, &k2, matches);
  for (int i = 0; i < n1; i++) match_of_1[i] = matches[i] ? (int)(matches[i] - m2.data()) : -1;
  return nm;
}

int reff_search_by_projection_last(const void *keys_un, const uint8_t *desc, int n, const float *u_right, float min_x, float max_x,
                                   float min_y, float max_y, const float *scale, int n_levels, float bf, float mb, const float *cam4,
                                   const float *t_cw, const float *t_lw, const void *last_keys, int n_last, const uint8_t *last_has_point,
                                   const uint8_t *last_outlier, const float *last_world, const uint8_t *last_desc,
                                   const uint8_t *pre_matched, float th, int mono, int check_orientation, int *assigned) {
  Frame F, L;
  GeometricCamera cam;
  cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
  F.cam_ = &cam;
  F.N = n;
  F.mvKeysUn.assign((const cv::KeyPoint *)keys_un, (const cv::KeyPoint *)keys_un + n);
  F.mvKeys = F.mvKeysUn;
  F.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)desc);
  F.mvuRight.assign(n, -1.0f);
  if (u_right) F.mvuRight.assign(u_right, u_right + n);
  F.mvScaleFactors.assign(scale, scale + n_levels);
  F.bf_ = bf;
  F.mb = mb;
  F.pose.t = Eigen::Vector3f(t_cw[0], t_cw[1], t_cw[2]);
  set_bounds(min_x, max_x, min_y, max_y);
  MapPoint occupied;
  occupied.nObs = 1;
  F.mvpMapPoints.assign(n, (MapPoint *)nullptr);
  for (int i = 0; i < n; i++)
    if (pre_matched && pre_matched[i]) F.mvpMapPoints[i] = &occupied;
  L.N = n_last;
  L.mvKeysUn.assign((const cv::KeyPoint *)last_keys, (const cv::KeyPoint *)last_keys + n_last);
  L.mvKeys = L.mvKeysUn;
  L.pose.t = Eigen::Vector3f(t_lw[0], t_lw[1], t_lw[2]);
  std::vector<MapPoint> mps(n_last > 0 ? n_last : 1);
  L.mvpMapPoints.assign(n_last, (MapPoint *)nullptr);
  L.mvbOutlier.assign(n_last, false);
  for (int i = 0; i < n_last; i++) {
    if (last_has_point[i]) L.mvpMapPoints[i] = &mps[i];
    L.mvbOutlier[i] = last_outlier[i] != 0;
    mps[i].nObs = 1;
    mps[i].world = Eigen::Vector3f(last_world[3 * i], last_world[3 * i + 1], last_world[3 * i + 2]);
    mps[i].mDescriptor = cv::Mat(1, 32, CV_8U, (void *)(last_desc + 32 * (size_t)i)).clone();
  }
  ORBmatcher matcher(0.9f, check_orientation != 0);
  const int nm = matcher.SearchByProjection(F, L, th, mono != 0);
  for (int i = 0; i < n; i++) {
    MapPoint *p = F.mvpMapPoints[i];
    assigned[i] = (p && p != &occupied) ? (int)(p - mps.data()) : -1;
  }
  return nm;
}

int reff_search_for_triangulation(const void *kps1, const uint8_t *desc1, int n1, const uint8_t *has_point1, const float *u_right1,
                                  const uint32_t *nodes1, const int *begin1, int n_nodes1, const uint32_t *feats1, int total1,
                                  const void *kps2, const uint8_t *desc2, int n2, const uint8_t *has_point2, const float *u_right2,
                                  const uint32_t *nodes2, const int *begin2, int n_nodes2, const uint32_t *feats2, int total2,
                                  const float *f12, const float *cam4, const float *c2, const float *scale, const float *sigma2,
                                  int n_levels, float nnratio, int only_stereo, int coarse, int check_orientation, float *ep_out,
                                  int *match_of_1) {
  KeyFrame k1, k2;
  std::vector<MapPoint> m1(n1 > 0 ? n1 : 1), m2(n2 > 0 ? n2 : 1);
  fill_kf(k1, m1, kps1, desc1, n1, has_point1, nodes1, begin1, n_nodes1, feats1, total1);
  fill_kf(k2, m2, kps2, desc2, n2, has_point2, nodes2, begin2, n_nodes2, feats2, total2);
  GeometricCamera cam1, cam2;
  cam1.fx = cam2.fx = cam4[0]; cam1.fy = cam2.fy = cam4[1]; cam1.cx = cam2.cx = cam4[2]; cam1.cy = cam2.cy = cam4[3];
  for (int i = 0; i < 9; i++) cam1.F12_given.m[i] = f12[i];
  k1.cam_ = &cam1; k2.cam_ = &cam2;
  k1.mvuRight.assign(u_right1, u_right1 + n1);
  k2.mvuRight.assign(u_right2, u_right2 + n2);
  k1.mvScaleFactors.assign(scale, scale + n_levels); k2.mvScaleFactors = k1.mvScaleFactors;
  k1.mvLevelSigma2.assign(sigma2, sigma2 + n_levels); k2.mvLevelSigma2 = k1.mvLevelSigma2;
  k2.pose.t = Eigen::Vector3f(c2[0], c2[1], c2[2]);
  const Eigen::Vector2f ep = cam2.Project(k2.pose * k1.GetCameraCenter());
  ep_out[0] = ep(0); ep_out[1] = ep(1);
  std::vector<std::pair<size_t, size_t> > pairs;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int nm = matcher.SearchForTriangulation(&k1, &k2, pairs, only_stereo != 0, coarse != 0);
  for (int i = 0; i < n1; i++) match_of_1[i] = -1;
  for (size_t j = 0; j < pairs.size(); j++) match_of_1[pairs[j].first] = (int)pairs[j].second;
  return nm;
}

// static ORBmatcher::DescriptorDistance on n row pairs
void facade_descriptor_distance(const uint8_t *a, const uint8_t *b, int n, int *out) {
  for (int i = 0; i < n; i++)
    out[i] = ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void *)(a + 32 * (size_t)i)), cv::Mat(1, 32, CV_8U, (void *)(b + 32 * (size_t)i)));
}

// the class-level entry points of the remaining ORBmatcher methods, shared with oracle/ref_frame_shim.cc
#include "matcher_harness.inc"

}  // extern "C"
