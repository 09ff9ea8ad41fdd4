// tests/cpp/standin/standin_types.h -- TEST INFRASTRUCTURE, not product code.
//
// Stand-ins for the reference headers ORBmatcher's drop-in (orb_slam_fusion_b200/cpp/src/orb_matcher.cc) includes --
// map/frame.h, map/keyframe.h, map/mappoint.h, sophus/sim3.hpp -- which need Eigen, Sophus, DBoW2 and boost, none of
// which exist in this image.  They carry the public members and methods the matcher touches, under the reference's
// names and types (include/map/frame.h, keyframe.h, mappoint.h), so the SAME source compiles here and against the
// real headers.  Poses are translations and cameras pinhole, like the stand-ins of oracle/ref_frame_shim.cc that the
// reference's own spliced bodies run on: the two sides of tests/test_cpp_matcher.py see the same numbers.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>

#define FRAME_GRID_ROWS 48  // include/map/frame.h:40-41
#define FRAME_GRID_COLS 64

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
struct Matrix3f {
  float m[9];  // row-major
  Matrix3f() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  float operator()(int r, int c) const { return m[3 * r + c]; }
};
}  // namespace Eigen

namespace Sophus {
struct SE3f {  // translation only
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
struct Sim3 {  // scale + translation
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

namespace DBoW2 {  // 3rdparty/DBoW2/DBoW2/FeatureVector.h:24
typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;
}

namespace ORB_SLAM_FUSION {

class MapPoint;
class KeyFrame;
class Frame;

struct GeometricCamera {  // a pinhole camera (camera_models/pinhole_model.cc)
  float fx = 1, fy = 1, cx = 0, cy = 0;
  Eigen::Vector2f Project(const Eigen::Vector3f &p) { return Eigen::Vector2f(fx * p(0) / p(2) + cx, fy * p(1) / p(2) + cy); }
  Eigen::Matrix3f F12_given;  // the harness supplies F12 (a product of Eigen matrices in the reference)
};

class Frame {
 public:
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors, mDescriptorsRight;
  std::vector<float> mvScaleFactors, mvInvScaleFactors;
  float mb = 0, bf_ = 0;
  std::vector<MapPoint *> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  static float mfGridElementWidthInv, mfGridElementHeightInv;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY;
  int Nleft = -1, Nright = -1;
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera *cam_ = nullptr, *cam2_ = nullptr;
  Sophus::SE3f pose;
  Sophus::SE3f GetPose() const { return pose; }

  // the frame grid (frame.cc:438-465, 679-759), restated for the stand-in: keypoints in insertion order per cell,
  // cells visited column-major
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
  void AssignFeaturesToGrid() {
    for (int i = 0; i < FRAME_GRID_COLS; i++)
      for (int j = 0; j < FRAME_GRID_ROWS; j++) mGrid[i][j].clear();
    for (int i = 0; i < N; i++) {
      const cv::KeyPoint &kp = mvKeysUn[i];
      const int px = (int)round((kp.pt.x - mnMinX) * mfGridElementWidthInv), py = (int)round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
      if (px < 0 || px >= FRAME_GRID_COLS || py < 0 || py >= FRAME_GRID_ROWS) continue;
      mGrid[px][py].push_back(i);
    }
  }
  std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1,
                                        const bool = false) const {
    std::vector<size_t> out;
    const int x0 = std::max(0, (int)floor((x - mnMinX - r) * mfGridElementWidthInv));
    const int x1 = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + r) * mfGridElementWidthInv));
    const int y0 = std::max(0, (int)floor((y - mnMinY - r) * mfGridElementHeightInv));
    const int y1 = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (x0 >= FRAME_GRID_COLS || x1 < 0 || y0 >= FRAME_GRID_ROWS || y1 < 0) return out;
    const bool levels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = x0; ix <= x1; ix++)
      for (int iy = y0; iy <= y1; iy++)
        for (size_t j = 0; j < mGrid[ix][iy].size(); j++) {
          const cv::KeyPoint &kp = mvKeysUn[mGrid[ix][iy][j]];
          if (levels && (kp.octave < minLevel || (maxLevel >= 0 && kp.octave > maxLevel))) continue;
          if (fabs(kp.pt.x - x) < r && fabs(kp.pt.y - y) < r) out.push_back(mGrid[ix][iy][j]);
        }
    return out;
  }
};

class MapPoint {
 public:
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  Eigen::Vector3f GetWorldPos() { return world; }
  Eigen::Vector3f GetNormal() { return normal; }
  bool isBad() { return mbBad; }
  int Observations() { return nObs; }
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
  void Replace(MapPoint *other) { mbBad = true; replaced_by = other; log.push_back(std::make_tuple(2, id, other->id)); }

  float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
  bool mbTrackInView = false, mbTrackInViewR = false;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
  float mTrackViewCos = 0, mTrackViewCosR = 0;

  // stand-in state
  int id = 0;
  Eigen::Vector3f world, normal;
  cv::Mat mDescriptor;
  bool mbBad = false;
  int nObs = 0, predicted_level = 0;
  float min_dist = 0, max_dist = 1e30f;
  std::map<KeyFrame *, int> in_kf;
  MapPoint *replaced_by = nullptr;
  static std::vector<std::tuple<int, int, int> > log;  // (1 = AddObservation | 2 = Replace | 3 = AddMapPoint, point id, keypoint / other id)
};

class KeyFrame {
 public:
  std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
  std::set<MapPoint *> GetMapPoints() {
    std::set<MapPoint *> s;
    for (size_t i = 0; i < mvpMapPoints.size(); i++)
      if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
    return s;
  }
  MapPoint *GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
  void AddMapPoint(MapPoint *mp, const size_t &idx) { mvpMapPoints[idx] = mp; MapPoint::log.push_back(std::make_tuple(3, mp->id, (int)idx)); }
  bool IsInImage(const float &x, const float &y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
  Sophus::SE3f GetPose() { return pose; }
  Sophus::SE3f GetPoseInverse() { return pose.inverse(); }
  Eigen::Vector3f GetCameraCenter() { return pose.inverse().translation(); }

  cv::Mat mDescriptors;
  std::vector<MapPoint *> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  int NLeft = -1, N = 0;
  GeometricCamera *cam_ = nullptr, *cam2_ = nullptr;
  std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  float fx = 1, fy = 1, cx = 0, cy = 0, bf_ = 0;
  int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
  int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
  float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  Sophus::SE3f pose;
  Frame *grid = nullptr;  // (harness bookkeeping shared with oracle/ref_frame_shim.cc; the product's class never reads it)
};

}  // namespace ORB_SLAM_FUSION
