// oracle/ref_frame_shim.cc -- TEST INFRASTRUCTURE, not product code.
//
// The reference's own lines for the matcher-side rows of the hot path, compiled where they lie.
// frame.cc / mappoint.cc / orb_matcher.cc cannot be compiled as files here (they include Eigen, Sophus,
// DBoW2, boost), so the FUNCTION BODIES are spliced by sed into build intermediates under _ref/
// (git-ignored, see oracle/Makefile) and compiled inside the minimal class declarations below, which
// carry exactly the members those bodies touch, under the reference's names:
//   _ref/frame_stereo.inc        frame.cc:828-986      Frame::ComputeStereoMatches
//   _ref/frame_grid.inc          frame.cc:438-465      Frame::AssignFeaturesToGrid
//   _ref/frame_area.inc          frame.cc:679-759      Frame::GetFeaturesInArea, Frame::PosInGrid
//   _ref/mappoint_distinct.inc   mappoint.cc:365-433   MapPoint::ComputeDistinctiveDescriptors
//   _ref/matcher_consts.inc      orb_matcher.cc:35-40  TH_HIGH / TH_LOW / HISTO_LENGTH, constructor
//   _ref/matcher_project.inc     orb_matcher.cc:42-213 ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, ...),
//                                                      ORBmatcher::RadiusByViewingCos
//   _ref/matcher_bow.inc         orb_matcher.cc:215-389  ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)
//   _ref/matcher_bow_kf.inc      orb_matcher.cc:697-815  ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)
//   _ref/matcher_project_last.inc orb_matcher.cc:1518-1728 ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
//   _ref/matcher_triangulation.inc orb_matcher.cc:817-1040 ORBmatcher::SearchForTriangulation
//   _ref/pinhole_epipolar.inc    pinhole_model.cc:121-134  the epipolar-line test of Pinhole::EpipolarConstrain (F12 given)
//   _ref/matcher_fuse_loop.inc   orb_matcher.cc:1145-1190  the candidate loop of ORBmatcher::Fuse (spliced into reff_fuse_search)
//   _ref/matcher_maxima.inc      orb_matcher.cc:1841-1873 ORBmatcher::ComputeThreeMaxima
//   _ref/descriptor_distance.inc orb_matcher.cc:1877-1891
// Used by tests/test_oracle_vs_ref_frame.py to pin orc_stereo_rowband / orc_stereo_refine /
// orc_distinctive / orc_window_search(_stereo) / orc_search_by_bow of oracle/orb_oracle.c.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>

using namespace std;

#define FRAME_GRID_ROWS 48  // include/map/frame.h:40-41
#define FRAME_GRID_COLS 64

// Stand-ins for the two Eigen / Sophus types orb_matcher.cc:1518-1728 touches, with exactly the operations it uses.  The
// pose is a pure translation (the harness chooses the poses), so x3Dc = x3Dw + t is one float add per component.
namespace Eigen {
struct Vector3f {
  float v[3];
  Vector3f() : v{0, 0, 0} {}
  Vector3f(float a, float b, float c) : v{a, b, c} {}
  float operator()(int i) const { return v[i]; }
  Vector3f operator-(const Vector3f &o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vector3f operator/(float s) const { return Vector3f(v[0] / s, v[1] / s, v[2] / s); }
  float dot(const Vector3f &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
  float norm() const { return std::sqrt(dot(*this)); }
};
struct Vector2f {
  float v[2];
  Vector2f() : v{0, 0} {}
  Vector2f(float a, float b) : v{a, b} {}
  float operator()(int i) const { return v[i]; }
};
}  // namespace Eigen
namespace Eigen {
struct Matrix3f {
  float m[9];  // row-major
  Matrix3f() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  float operator()(int r, int c) const { return m[3 * r + c]; }
};
}  // namespace Eigen
namespace Sophus {
struct SE3f {
  Eigen::Vector3f t;
  SE3f() {}
  SE3f(const Eigen::Matrix3f &, const Eigen::Vector3f &tt) : t(tt) {}
  SE3f operator*(const SE3f &o) const { SE3f r; r.t = Eigen::Vector3f(t(0) + o.t(0), t(1) + o.t(1), t(2) + o.t(2)); return r; }
  Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
  SE3f inverse() const { SE3f r; r.t = Eigen::Vector3f(-t(0), -t(1), -t(2)); return r; }
  Eigen::Vector3f translation() const { return t; }
  Eigen::Vector3f operator*(const Eigen::Vector3f &p) const { return Eigen::Vector3f(p(0) + t(0), p(1) + t(1), p(2) + t(2)); }
};
template <class T>
struct Sim3 {  // scale + translation (the rotation of the harness' similarity is the identity)
  float s = 1;
  Eigen::Vector3f t;
  Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
  Eigen::Vector3f translation() const { return t; }
  float scale() const { return s; }
  Sim3 inverse() const { Sim3 r; r.s = 1 / s; r.t = Eigen::Vector3f(-t(0) / s, -t(1) / s, -t(2) / s); return r; }
  Eigen::Vector3f operator*(const Eigen::Vector3f &p) const { return Eigen::Vector3f(s * p(0) + t(0), s * p(1) + t(1), s * p(2) + t(2)); }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

namespace DBoW2 {  // 3rdparty/DBoW2/DBoW2/FeatureVector.h:24: a std::map from node id to feature indices
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;
}

namespace ORB_SLAM_FUSION {

class MapPoint;
class KeyFrame;
struct GeometricCamera {  // a pinhole camera (camera_models/pinhole_model.cc: fx * x / z + cx)
  float fx = 1, fy = 1, cx = 0, cy = 0;
  Eigen::Vector2f Project(const Eigen::Vector3f &p) { return Eigen::Vector2f(fx * p(0) / p(2) + cx, fy * p(1) / p(2) + cy); }
  // Pinhole::EpipolarConstrain (pinhole_model.cc:109-135): the fundamental matrix of :116-119 is a product of Eigen
  // matrices (K1^-T [t12]x R12 K2^-1); the harness supplies it, the line test below is the reference's own :121-134.
  Eigen::Matrix3f F12_given;
  bool EpipolarConstrain(GeometricCamera *cam_2, const cv::KeyPoint &kp_1, const cv::KeyPoint &kp_2, const Eigen::Matrix3f &R12_eig,
                         const Eigen::Vector3f &t12_eig, const float sigma_lev, const float unc) {
    const Eigen::Matrix3f &F12 = F12_given;
#include "pinhole_epipolar.inc"
  }
};

// the one member of OrbExtractor that frame.cc:834,913-933 reads
struct OrbExtractor {
  std::vector<cv::Mat> img_pyramid_;
};

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);
  static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
  int SearchByProjection(class Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3,
                         const bool bFarPoints = false, const float thFarPoints = 50.0f);
  int SearchByProjection(class Frame &CurrentFrame, const class Frame &LastFrame, const float th, const bool bMono);
  int SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<std::pair<size_t, size_t> > &vMatchedPairs,
                             const bool bOnlyStereo, const bool bCoarse = false);
  int SearchByBoW(KeyFrame *pKF, class Frame &F, std::vector<MapPoint *> &vpMapPointMatches);
  int SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12);
  int SearchByProjection(class Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th, const int ORBdist);
  int SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint *> &vpPoints, std::vector<MapPoint *> &vpMatched,
                         int th, float ratioHamming = 1.0);
  int SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint *> &vpPoints,
                         const std::vector<KeyFrame *> &vpPointsKFs, std::vector<MapPoint *> &vpMatched,
                         std::vector<KeyFrame *> &vpMatchedKF, int th, float ratioHamming = 1.0);
  int SearchForInitialization(class Frame &F1, class Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12,
                              int windowSize = 10);
  int SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12, const Sophus::Sim3f &S12, const float th);
  int Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th = 3.0, const bool bRight = false);
  int Fuse(KeyFrame *pKF, Sophus::Sim3f &Scw, const std::vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint);
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;

 protected:
  float RadiusByViewingCos(const float &viewCos);
  void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);
  float mfNNratio;
  bool mbCheckOrientation;
};

class Frame {  // include/map/frame.h: the members the spliced bodies use, same names and types
 public:
  void ComputeStereoMatches();
  void AssignFeaturesToGrid();
  bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);
  vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                   const int maxLevel = -1, const bool bRight = false) const;

  OrbExtractor *orb_extractor_left_ = nullptr, *orb_extractor_right_ = nullptr;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors, mDescriptorsRight;
  std::vector<float> mvScaleFactors, mvInvScaleFactors;
  float mb = 0, bf_ = 0;
  std::vector<MapPoint *> mvpMapPoints;
  static float mfGridElementWidthInv, mfGridElementHeightInv;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  int Nleft = -1, Nright = -1;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera *cam_ = nullptr, *cam2_ = nullptr;
  Sophus::SE3f pose, Trl;
  Sophus::SE3f GetPose() const { return pose; }
  Sophus::SE3f GetRelativePoseTrl() const { return Trl; }
  std::vector<bool> mvbOutlier;
};
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

class KeyFrame {  // include/map/keyframe.h: what mappoint.cc:365-433 and orb_matcher.cc:215-389 touch
 public:
  bool isBad() { return bad; }
  std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
  bool bad = false;
  cv::Mat mDescriptors;
  std::vector<MapPoint *> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  int NLeft = -1, N = 0;
  GeometricCamera *cam_ = nullptr, *cam2_ = nullptr;
  // orb_matcher.cc:817-1040
  MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
  std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  // KeyFrame::GetFeaturesInArea (keyframe.cc:729-773) visits the frame's grid exactly like Frame::GetFeaturesInArea without a level gate
  Frame *grid = nullptr;
  std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const bool bRight = false) const {
    return grid->GetFeaturesInArea(x, y, r, -1, -1, bRight);
  }
  Sophus::SE3f pose;  // Tcw, translation only
  // orb_matcher.cc:391-596, 1042-1516, 1730-1840
  float fx = 1, fy = 1, cx = 0, cy = 0, bf_ = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  bool IsInImage(const float &x, const float &y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }  // keyframe.cc:775-777
  std::set<MapPoint *> GetMapPoints();
  void AddMapPoint(MapPoint *mp, const size_t &idx);
  Eigen::Vector3f GetRightCameraCenter() { return pose.inverse().translation(); }
  Sophus::SE3f GetPose() { return pose; }
  Sophus::SE3f GetPoseInverse() { return pose.inverse(); }
  Sophus::SE3f GetRightPose() { return pose; }
  Sophus::SE3f GetRightPoseInverse() { return pose.inverse(); }
  Eigen::Vector3f GetCameraCenter() { return pose.inverse().translation(); }
};

class MapPoint {  // include/map/mappoint.h
 public:
  void ComputeDistinctiveDescriptors();
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  Eigen::Vector3f GetWorldPos() { return world; }
  Eigen::Vector3f world;
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }

  float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
  bool mbTrackInView = false, mbTrackInViewR = false;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
  float mTrackViewCos = 0, mTrackViewCosR = 0;

  std::map<KeyFrame *, std::tuple<int, int>> mObservations;
  cv::Mat mDescriptor;
  bool mbBad = false;
  int nObs = 0;
  std::mutex mMutexFeatures;
  // the rest of the MapPoint interface the other matcher methods call, with the harness' values behind it
  int id = 0, predicted_level = 0;
  Eigen::Vector3f normal;
  float min_dist = 0, max_dist = 1e30f;
  std::map<KeyFrame *, int> in_kf;
  Eigen::Vector3f GetNormal() { return normal; }
  float GetMaxDistanceInvariance() { return max_dist; }
  float GetMinDistanceInvariance() { return min_dist; }
  int PredictScale(const float &, KeyFrame *) { return predicted_level; }
  int PredictScale(const float &, Frame *) { return predicted_level; }
  bool IsInKeyFrame(KeyFrame *kf) { return in_kf.count(kf) != 0; }
  std::tuple<int, int> GetIndexInKeyFrame(KeyFrame *kf) {
    std::map<KeyFrame *, int>::iterator it = in_kf.find(kf);
    return std::make_tuple(it == in_kf.end() ? -1 : it->second, -1);
  }
  void AddObservation(KeyFrame *kf, int idx) { in_kf[kf] = idx; nObs++; log.push_back(std::make_tuple(1, id, idx)); }
  void Replace(MapPoint *other) { mbBad = true; log.push_back(std::make_tuple(2, id, other->id)); }
  static std::vector<std::tuple<int, int, int> > log;  // (1 AddObservation | 2 Replace | 3 AddMapPoint, point id, keypoint / other id)
};
std::vector<std::tuple<int, int, int> > MapPoint::log;
std::set<MapPoint *> KeyFrame::GetMapPoints() {
  std::set<MapPoint *> s;
  for (size_t i = 0; i < mvpMapPoints.size(); i++)
    if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
  return s;
}
void KeyFrame::AddMapPoint(MapPoint *mp, const size_t &idx) { mvpMapPoints[idx] = mp; MapPoint::log.push_back(std::make_tuple(3, mp->id, (int)idx)); }

#include "matcher_consts.inc"
#include "descriptor_distance.inc"
#include "frame_grid.inc"
#include "frame_area.inc"
#include "frame_stereo.inc"
#include "mappoint_distinct.inc"
#include "matcher_project.inc"
#include "matcher_project_last.inc"
#include "matcher_bow.inc"
#include "matcher_bow_kf.inc"
#include "matcher_triangulation.inc"
#include "matcher_maxima.inc"
#include "matcher_project_sim3.inc"
#include "matcher_project_sim3_kfs.inc"
#include "matcher_init.inc"
#include "matcher_fuse.inc"
#include "matcher_fuse_sim3.inc"
#include "matcher_sim3.inc"
#include "matcher_project_reloc.inc"

}  // namespace ORB_SLAM_FUSION

using namespace ORB_SLAM_FUSION;

// the spliced loop of Fuse calls DescriptorDistance unqualified (it is a member function there)
static inline int DescriptorDistance(const cv::Mat &a, const cv::Mat &b) { return ORBmatcher::DescriptorDistance(a, b); }

extern "C" {

struct reff_level {
  const uint8_t *px;  // pixel (0,0) of the level (the 19-px border lies around it, like img_pyramid_)
  int w, h;
  size_t stride;
};

// Frame::ComputeStereoMatches on caller-supplied keypoints, descriptors and the two pyramids.
void reff_stereo_matches(const reff_level *left, const reff_level *right, int n_levels, const void *kl, int nl,
                         const uint8_t *dl, const void *kr, int nr, const uint8_t *dr, const float *scale,
                         const float *inv_scale, float bf, float mb, float *u_right, float *depth) {
  OrbExtractor el, er;
  for (int l = 0; l < n_levels; l++) {
    el.img_pyramid_.push_back(cv::Mat(left[l].h, left[l].w, CV_8UC1, (void *)left[l].px, left[l].stride));
    er.img_pyramid_.push_back(cv::Mat(right[l].h, right[l].w, CV_8UC1, (void *)right[l].px, right[l].stride));
  }
  Frame F;
  F.orb_extractor_left_ = &el;
  F.orb_extractor_right_ = &er;
  F.N = nl;
  F.mvKeys.assign((const cv::KeyPoint *)kl, (const cv::KeyPoint *)kl + nl);
  F.mvKeysRight.assign((const cv::KeyPoint *)kr, (const cv::KeyPoint *)kr + nr);
  F.mDescriptors = cv::Mat(nl, 32, CV_8U, (void *)dl);
  F.mDescriptorsRight = cv::Mat(nr, 32, CV_8U, (void *)dr);
  F.mvScaleFactors.assign(scale, scale + n_levels);
  F.mvInvScaleFactors.assign(inv_scale, inv_scale + n_levels);
  F.bf_ = bf;
  F.mb = mb;
  F.ComputeStereoMatches();
  std::memcpy(u_right, F.mvuRight.data(), sizeof(float) * nl);
  std::memcpy(depth, F.mvDepth.data(), sizeof(float) * nl);
}

// MapPoint::ComputeDistinctiveDescriptors for point p = rows [offsets[p], offsets[p+1]) of desc: every row is
// the left observation of its own key frame (key frames allocated in one array, so the std::map iterates
// them in row order).  Writes the chosen 32-byte descriptor; chosen[p] = 0 where mDescriptor stays empty.
void reff_distinctive(const uint8_t *desc, const int *offsets, int n_points, uint8_t *out_desc, int *chosen) {
  for (int p = 0; p < n_points; p++) {
    const int n = offsets[p + 1] - offsets[p];
    std::vector<KeyFrame> kfs(n > 0 ? n : 1);
    MapPoint mp;
    for (int i = 0; i < n; i++) {
      kfs[i].mDescriptors = cv::Mat(1, 32, CV_8U, (void *)(desc + 32 * (size_t)(offsets[p] + i)));
      mp.mObservations[&kfs[i]] = std::make_tuple(0, -1);
    }
    mp.ComputeDistinctiveDescriptors();
    chosen[p] = !mp.mDescriptor.empty();
    if (chosen[p]) std::memcpy(out_desc + 32 * (size_t)p, mp.mDescriptor.data, 32);
  }
}

struct reff_track_point {  // the MapPoint fields SearchByProjection reads (orb_matcher.cc:53-92)
  float proj_x, proj_y, proj_xr, view_cos, depth;
  int level, in_view, bad;
};

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints) on a
// monocular / rectified-stereo frame (Nleft == -1): keys_un + descriptors + optional mvuRight, the 64x48 grid built
// by the reference's AssignFeaturesToGrid.  pre_matched[i] != 0: the frame keypoint already holds a map point
// with observations.  assigned[i] = index of the map point the call stored in F.mvpMapPoints[i] (-1: untouched).
int reff_search_by_projection(const void *keys_un, const uint8_t *desc, int n, const float *u_right, float min_x,
                              float max_x, float min_y, float max_y, const float *scale, int n_levels,
                              const reff_track_point *pts, const uint8_t *pt_desc, int n_pts, const uint8_t *pre_matched,
                              float th, float nnratio, int far_points, float th_far, int *assigned) {
  Frame F;
  F.N = n;
  F.mvKeysUn.assign((const cv::KeyPoint *)keys_un, (const cv::KeyPoint *)keys_un + n);
  F.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)desc);
  F.mvuRight.assign(n, -1.0f);
  if (u_right) F.mvuRight.assign(u_right, u_right + n);
  F.mvScaleFactors.assign(scale, scale + n_levels);
  Frame::mnMinX = min_x;
  Frame::mnMaxX = max_x;
  Frame::mnMinY = min_y;
  Frame::mnMaxY = max_y;
  // frame.cc:214-217
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
  F.AssignFeaturesToGrid();
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

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches): key frame = (keys, descriptors, has_point flags: the
// feature holds a map point that is not bad, FeatureVector as sorted nodes / group starts / feature indices), frame
// likewise.  match_of_f[i] = key-frame feature whose map point vpMapPointMatches[i] is, -1 = NULL.
int reff_search_by_bow(const void *kps_kf, const uint8_t *desc_kf, int n_kf, const uint8_t *has_point_kf, const uint32_t *nodes_kf,
                       const int *begin_kf, int n_nodes_kf, const uint32_t *feats_kf, int total_kf, const void *kps_f,
                       const uint8_t *desc_f, int n_f, const uint32_t *nodes_f, const int *begin_f, int n_nodes_f,
                       const uint32_t *feats_f, int total_f, float nnratio, int check_orientation, int *match_of_f) {
  KeyFrame kf;
  Frame F;
  kf.mvKeysUn.assign((const cv::KeyPoint *)kps_kf, (const cv::KeyPoint *)kps_kf + n_kf);
  kf.mvKeys = kf.mvKeysUn;
  kf.mDescriptors = cv::Mat(n_kf, 32, CV_8U, (void *)desc_kf);
  std::vector<MapPoint> mps(n_kf > 0 ? n_kf : 1);
  kf.mvpMapPoints.assign(n_kf, (MapPoint *)nullptr);
  for (int i = 0; i < n_kf; i++)
    if (!has_point_kf || has_point_kf[i]) kf.mvpMapPoints[i] = &mps[i];
  for (int j = 0; j < n_nodes_kf; j++)
    kf.mFeatVec[nodes_kf[j]].assign(feats_kf + begin_kf[j], feats_kf + (j + 1 < n_nodes_kf ? begin_kf[j + 1] : total_kf));
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

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12): match_of_1[i] = side-2 feature whose map point
// vpMatches12[i] is, -1 = NULL.
static void fill_kf(KeyFrame &kf, std::vector<MapPoint> &mps, const void *kps, const uint8_t *desc, int n, const uint8_t *has_point,
                    const uint32_t *nodes, const int *begin, int n_nodes, const uint32_t *feats, int total) {
  kf.mvKeysUn.assign((const cv::KeyPoint *)kps, (const cv::KeyPoint *)kps + n);
  kf.mvKeys = kf.mvKeysUn;
  kf.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)desc);
  kf.mvpMapPoints.assign(n, (MapPoint *)nullptr);
  for (int i = 0; i < n; i++)
    if (!has_point || has_point[i]) kf.mvpMapPoints[i] = &mps[i];
  for (int j = 0; j < n_nodes; j++)
    kf.mFeatVec[nodes[j]].assign(feats + begin[j], feats + (j + 1 < n_nodes ? begin[j + 1] : total));
}

int reff_search_by_bow_kf(const void *kps1, const uint8_t *desc1, int n1, const uint8_t *has_point1, const uint32_t *nodes1,
                          const int *begin1, int n_nodes1, const uint32_t *feats1, int total1, const void *kps2,
                          const uint8_t *desc2, int n2, const uint8_t *has_point2, const uint32_t *nodes2, const int *begin2,
                          int n_nodes2, const uint32_t *feats2, int total2, float nnratio, int check_orientation,
                          int *match_of_1) {
  KeyFrame k1, k2;
  std::vector<MapPoint> m1(n1 > 0 ? n1 : 1), m2(n2 > 0 ? n2 : 1);  // MapPoint holds a mutex: sized at construction
  fill_kf(k1, m1, kps1, desc1, n1, has_point1, nodes1, begin1, n_nodes1, feats1, total1);
  fill_kf(k2, m2, kps2, desc2, n2, has_point2, nodes2, begin2, n_nodes2, feats2, total2);
  std::vector<MapPoint *> matches;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int nm = matcher.SearchByBoW(&k1, &k2, matches);
  for (int i = 0; i < n1; i++) match_of_1[i] = matches[i] ? (int)(matches[i] - m2.data()) : -1;
  return nm;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono): the current frame as in
// reff_search_by_projection (+ bf, mb and a pinhole camera), the last frame as keypoints (octave, angle), per feature a
// map point or none (has_point), outlier flags, world positions and descriptors; both poses are pure translations.
// assigned[i] = last-frame feature whose map point CurrentFrame.mvpMapPoints[i] receives (-1: untouched).
int reff_search_by_projection_last(const void *keys_un, const uint8_t *desc, int n, const float *u_right, float min_x, float max_x,
                                   float min_y, float max_y, const float *scale, int n_levels, float bf, float mb, const float *cam4,
                                   const float *t_cw, const float *t_lw, const void *last_keys, int n_last,
                                   const uint8_t *last_has_point, const uint8_t *last_outlier, const float *last_world,
                                   const uint8_t *last_desc, const uint8_t *pre_matched, float th, int mono, int check_orientation,
                                   int *assigned) {
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
  Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
  F.AssignFeaturesToGrid();
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

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse): two key frames with one pinhole
// camera each (fill_kf + mvuRight, scale factors, level sigma^2); F12 (row-major) is handed to the camera stand-in, the
// poses are translations chosen so that the epipole cam_->Project(T2w * Cw) is (ep_x, ep_y) = cam(c2).
int reff_search_for_triangulation(const void *kps1, const uint8_t *desc1, int n1, const uint8_t *has_point1, const float *u_right1,
                                  const uint32_t *nodes1, const int *begin1, int n_nodes1, const uint32_t *feats1, int total1,
                                  const void *kps2, const uint8_t *desc2, int n2, const uint8_t *has_point2, const float *u_right2,
                                  const uint32_t *nodes2, const int *begin2, int n_nodes2, const uint32_t *feats2, int total2,
                                  const float *f12, const float *cam4, const float *c2 /* T2w * Cw */, const float *scale,
                                  const float *sigma2, int n_levels, float nnratio, int only_stereo, int coarse,
                                  int check_orientation, float *ep_out, int *match_of_1) {
  KeyFrame k1, k2;
  std::vector<MapPoint> m1(n1 > 0 ? n1 : 1), m2(n2 > 0 ? n2 : 1);
  fill_kf(k1, m1, kps1, desc1, n1, has_point1, nodes1, begin1, n_nodes1, feats1, total1);
  fill_kf(k2, m2, kps2, desc2, n2, has_point2, nodes2, begin2, n_nodes2, feats2, total2);
  GeometricCamera cam1, cam2;
  cam1.fx = cam2.fx = cam4[0]; cam1.fy = cam2.fy = cam4[1]; cam1.cx = cam2.cx = cam4[2]; cam1.cy = cam2.cy = cam4[3];
  for (int i = 0; i < 9; i++) cam1.F12_given.m[i] = f12[i];
  k1.cam_ = &cam1; k2.cam_ = &cam2;
  k1.N = n1; k2.N = n2;
  k1.mvuRight.assign(u_right1, u_right1 + n1);
  k2.mvuRight.assign(u_right2, u_right2 + n2);
  k1.mvScaleFactors.assign(scale, scale + n_levels); k2.mvScaleFactors = k1.mvScaleFactors;
  k1.mvLevelSigma2.assign(sigma2, sigma2 + n_levels); k2.mvLevelSigma2 = k1.mvLevelSigma2;
  // Cw = -t1 with t1 = 0, so Cw = 0 and C2 = T2w * Cw = t2 = c2
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

// The search of ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight) for one key frame (NLeft == -1) and n_pts projected map
// points: uv, ur, nPredictedLevel and radius are given (the geometry of :1060-1131 stays with the caller); the candidate
// loop is the reference's own lines 1145-1190.  best_idx / best_dist per point (-1 / 256 when nothing passes).
void reff_fuse_search(const void *keys_un, const uint8_t *desc, int n, const float *u_right, float min_x, float max_x, float min_y,
                      float max_y, const float *inv_sigma2, int n_levels, const float *pt_u, const float *pt_v, const float *pt_ur,
                      const float *pt_radius, const int *pt_level, const uint8_t *pt_desc, int n_pts, int *best_idx, int *best_dist) {
  Frame F;
  F.N = n;
  F.mvKeysUn.assign((const cv::KeyPoint *)keys_un, (const cv::KeyPoint *)keys_un + n);
  Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
  F.AssignFeaturesToGrid();
  KeyFrame kf;
  KeyFrame *pKF = &kf;
  kf.grid = &F;
  kf.mvKeysUn = F.mvKeysUn;
  kf.mvKeys = F.mvKeysUn;
  kf.mDescriptors = cv::Mat(n, 32, CV_8U, (void *)desc);
  kf.mvuRight.assign(n, -1.0f);
  if (u_right) kf.mvuRight.assign(u_right, u_right + n);
  kf.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + n_levels);
  const bool bRight = false;
  for (int p = 0; p < n_pts; p++) {
    const Eigen::Vector2f uv(pt_u[p], pt_v[p]);
    const float ur = pt_ur[p];
    const int nPredictedLevel = pt_level[p];
    const float radius = pt_radius[p];
    const vector<size_t> vIndices = pKF->GetFeaturesInArea(uv(0), uv(1), radius, bRight);  // :1133-1134
    MapPoint mp;
    MapPoint *pMP = &mp;
    mp.mDescriptor = cv::Mat(1, 32, CV_8U, (void *)(pt_desc + 32 * (size_t)p)).clone();
    const cv::Mat dMP = pMP->GetDescriptor();  // :1143
#include "matcher_fuse_loop.inc"
    best_idx[p] = bestIdx;
    best_dist[p] = bestDist;
  }
}

// Frame::GetFeaturesInArea after AssignFeaturesToGrid: the visiting order of the grid lookup.
int reff_features_in_area(const void *keys_un, int n, float min_x, float max_x, float min_y, float max_y, float x, float y,
                          float r, int min_level, int max_level, int *out, int cap) {
  Frame F;
  F.N = n;
  F.mvKeysUn.assign((const cv::KeyPoint *)keys_un, (const cv::KeyPoint *)keys_un + n);
  Frame::mnMinX = min_x;
  Frame::mnMaxX = max_x;
  Frame::mnMinY = min_y;
  Frame::mnMaxY = max_y;
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(max_x - min_x);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(max_y - min_y);
  F.AssignFeaturesToGrid();
  const vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
  return (int)v.size();
}

// the class-level entry points of the remaining ORBmatcher methods, shared with tests/cpp/matcher_facade_harness.cc
#include "matcher_harness.inc"

}  // extern "C"
