// Drop-in replacement for the reference's include/cam/orb_feature/orb_matcher.h (:36-129).
//
// Same namespace, class name, constructor, the static DescriptorDistance, the thirteen Search* / Fuse
// methods with the reference's signatures and defaults, the three public constants and the two protected
// members, so every call site -- tracking.cc:2053, 2192, 2687, 2909, 2978; localmapping.cc:382, 677-708;
// loopclosing.cc:571, 671, 702, 877, 2003, 2044 -- keeps compiling unchanged.  The bodies
// (cpp/src/orb_matcher.cc) gather the fields a search reads into plain arrays, run the search on the GPU
// through the C ABI of include/orbx.h (liborbx_b200.so) and scatter the resulting indices back into the
// callers' pointer vectors.  ORBmatcher objects are created on the stack all over the reference
// (`ORBmatcher matcher(0.9, true);`), so the class itself owns nothing: the GPU handle is one per host
// thread, created on first use.  ORBX_DEVICE selects the CUDA ordinal, as for OrbExtractor.
//
// There is no CPU fallback.  Frames of a two-camera rig (Frame::Nleft != -1 / KeyFrame::NLeft != -1) are
// not supported by the GPU searches yet: those calls throw std::runtime_error instead of silently taking
// another path.  SearchForInitialization (monocular map initialisation, tracking.cc:1823), whose matches
// depend on each other through a running per-keypoint best distance, is plain host code.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <set>
#include <utility>
#include <vector>

#include "map/frame.h"
#include "map/keyframe.h"
#include "map/mappoint.h"
#include "sophus/sim3.hpp"

namespace ORB_SLAM_FUSION {

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);

  // Hamming distance of two 256-bit descriptors (one row each).  Single pairs are eight popcounts on the
  // host; the batched form is orbm_hamming_pairs (include/orbx.h).
  static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

  // Tracking::SearchLocalPoints (tracking.cc:2687): projected local map points against the frame.
  int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3,
                         const bool bFarPoints = false, const float thFarPoints = 50.0f);

  // Tracking::TrackWithMotionModel (tracking.cc:2192, 2203).
  int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);

  // Tracking::Relocalization (tracking.cc:2978, 2992).
  int SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound,
                         const float th, const int ORBdist);

  // Loop detection (loopclosing.cc:877).
  int SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint *> &vpPoints,
                         std::vector<MapPoint *> &vpMatched, int th, float ratioHamming = 1.0);

  // Place recognition (loopclosing.cc:671, 702).
  int SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Scw, const std::vector<MapPoint *> &vpPoints,
                         const std::vector<KeyFrame *> &vpPointsKFs, std::vector<MapPoint *> &vpMatched,
                         std::vector<KeyFrame *> &vpMatchedKF, int th, float ratioHamming = 1.0);

  // Bag-of-words guided matching (tracking.cc:2053, 2909; loopclosing.cc:571).
  int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint *> &vpMapPointMatches);
  int SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12);

  // Monocular map initialisation (tracking.cc:1823).
  int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched,
                              std::vector<int> &vnMatches12, int windowSize = 10);

  // LocalMapping::CreateNewMapPoints (localmapping.cc:382).
  int SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<std::pair<size_t, size_t> > &vMatchedPairs,
                             const bool bOnlyStereo, const bool bCoarse = false);

  // Sim3 refinement of a loop candidate (loopclosing.cc:2003).
  int SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint *> &vpMatches12, const Sophus::Sim3f &S12,
                   const float th);

  // LocalMapping::SearchInNeighbors (localmapping.cc:677-708) and loop fusion (loopclosing.cc:2044).
  int Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th = 3.0, const bool bRight = false);
  int Fuse(KeyFrame *pKF, Sophus::Sim3f &Scw, const std::vector<MapPoint *> &vpPoints, float th,
           std::vector<MapPoint *> &vpReplacePoint);

 public:
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;
#ifdef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

 protected:
  float RadiusByViewingCos(const float &viewCos);

  void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);

  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM_FUSION

#endif  // ORBMATCHER_H
