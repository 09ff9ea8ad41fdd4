// GPU counterparts of the data-parallel loops around ORBmatcher::DescriptorDistance
// (orb_matcher.cc:1877-1891) over the C ABI of include/orbx.h.  The reference's ORBmatcher class
// (orb_matcher.h:36-129) keeps its 13 pointer-chasing Search*/Fuse methods; the call sites that are
// pure descriptor work switch to these (INTEGRATION.md shows the three-line patches):
//   Frame::ComputeStereoMatches       frame.cc:836-900   -> StereoRowBand
//   Frame::ComputeStereoFishEyeMatches frame.cc:1154-1162 -> KnnMatch2 + RatioTest
//   SearchByProjection inner loop      orb_matcher.cc:66-113 -> WindowSearch (greedy claim stays on the host)
//   SearchByProjection(Frame, MapPoints) orb_matcher.cc:42-134 -> SearchByProjection (whole function, greedy claim on the device)
//   ORBmatcher::SearchByBoW(KF, Frame) orb_matcher.cc:215-389 -> SearchByBoW (whole function, greedy claim included)
#ifndef ORBMATCHER_GPU_H
#define ORBMATCHER_GPU_H

#include <cstdint>
#include <map>
#include <utility>
#include <opencv2/opencv.hpp>
#include <vector>

struct orbm_matcher;  // include/orbx.h
struct orbx_extractor;

namespace ORB_SLAM_FUSION {

class ORBmatcherGpu {
 public:
  static const int TH_LOW = 50;  // orb_matcher.cc:35-37
  static const int TH_HIGH = 100;
  static const int HISTO_LENGTH = 30;

  explicit ORBmatcherGpu(int device = 0);
  ~ORBmatcherGpu();
  ORBmatcherGpu(const ORBmatcherGpu&) = delete;
  ORBmatcherGpu& operator=(const ORBmatcherGpu&) = delete;

  // ORBmatcher::DescriptorDistance for row i of a against row i of b (n x 32 CV_8U each).
  std::vector<int> DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

  // cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2): per query the two nearest train
  // rows ordered by (distance, index); fewer entries when train has fewer than 2 rows.
  void KnnMatch2(const cv::Mat& query, const cv::Mat& train, std::vector<std::vector<cv::DMatch> >& matches);

  // frame.cc:836-900: best right keypoint and distance per left keypoint (-1 / TH_HIGH when none).
  void StereoRowBand(const std::vector<cv::KeyPoint>& keys_left, const cv::Mat& desc_left,
                     const std::vector<cv::KeyPoint>& keys_right, const cv::Mat& desc_right,
                     const std::vector<float>& scale_factors, int n_rows, float min_d, float max_d,
                     std::vector<int>& best_idx_right, std::vector<int>& best_dist);

  // Frame::ComputeStereoMatches (frame.cc:828-986) on the device, for the frames the two extractors
  // processed last (the reference reads their img_pyramid_): fills mvuRight / mvDepth (-1 = no match).
  // bf and mb are Frame::bf_ and Frame::mb (minZ = mb, maxD = bf / minZ, frame.cc:853-856).
  void ComputeStereoMatches(orbx_extractor* left, orbx_extractor* right, const std::vector<cv::KeyPoint>& keys_left,
                            const cv::Mat& desc_left, const std::vector<cv::KeyPoint>& keys_right, const cv::Mat& desc_right,
                            const std::vector<float>& scale_factors, int n_rows, float bf, float mb,
                            std::vector<float>& u_right, std::vector<float>& depth);

  // MapPoint::ComputeDistinctiveDescriptors (mappoint.cc:365-428) for many map points at once:
  // observations[p] holds the descriptor rows of point p (the vDescriptors of :367-398); returns for
  // each point the index of the row with the least median distance to the others (-1 if empty).
  std::vector<int> ComputeDistinctiveDescriptors(const std::vector<std::vector<cv::Mat> >& observations);

  // ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th, bFarPoints, thFarPoints)
  // (orb_matcher.cc:42-134, Nleft == -1; Tracking::SearchLocalPoints, tracking.cc:2687) as a whole.  TrackedPoint carries
  // the MapPoint fields the function reads (:50-70, :90): mTrackProjX/Y, mTrackProjXR, mTrackViewCos, mTrackDepth,
  // mnTrackScaleLevel, mbTrackInView, isBad(); point_desc row p = pMP->GetDescriptor().  keys_un / desc / u_right =
  // F.mvKeysUn / F.mDescriptors / F.mvuRight (u_right may be empty: monocular), already_matched[i] = F.mvpMapPoints[i]
  // holds a point with observations.  assigned_point[i] = index into `points` of the map point F.mvpMapPoints[i]
  // receives (-1: untouched); returns nmatches.  The map points are taken to have observations (local map points do).
  struct TrackedPoint { float proj_x, proj_y, proj_xr, view_cos, depth; int level; bool in_view, bad; };
  int SearchByProjection(const std::vector<cv::KeyPoint>& keys_un, const cv::Mat& desc, const std::vector<float>& u_right,
                         const std::vector<float>& scale_factors, float min_x, float min_y, float grid_inv_w, float grid_inv_h,
                         int grid_cols, int grid_rows, const std::vector<TrackedPoint>& points, const cv::Mat& point_desc,
                         const std::vector<uint8_t>& already_matched, float th, bool far_points, float th_far_points, float nnratio,
                         std::vector<int>& assigned_point);

  // ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (orb_matcher.cc:1518-1728,
  // Nleft == -1; Tracking::TrackWithMotionModel) after the projection of :1539-1566, which stays host geometry:
  // ProjectedPoint p = one map point of the last frame that projects into the current one, in LastFrame order:
  // (u, v) = cam_->Project(Tcw * x3Dw), invzc, last_octave / last_angle = LastFrame's keypoint; point_desc row p =
  // pMP->GetDescriptor().  forward / backward = bForward / bBackward of :1535-1536.  assigned_point[i] = index into
  // `points` of the map point CurrentFrame.mvpMapPoints[i] receives (-1: untouched); returns nmatches.
  struct ProjectedPoint { float u, v, invzc; int last_octave; float last_angle; };
  int SearchByProjectionLastFrame(const std::vector<cv::KeyPoint>& keys_un, const cv::Mat& desc, const std::vector<float>& u_right,
                                  const std::vector<float>& scale_factors, float bf, float min_x, float min_y, float grid_inv_w,
                                  float grid_inv_h, int grid_cols, int grid_rows, const std::vector<ProjectedPoint>& points,
                                  const cv::Mat& point_desc, const std::vector<uint8_t>& already_matched, float th, bool forward,
                                  bool backward, bool check_orientation, std::vector<int>& assigned_point);

  // ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (orb_matcher.cc:215-389, Nleft == -1):
  // keys_kf = pKF->mvKeysUn, desc_kf = pKF->mDescriptors, has_point_kf[i] = pKF's feature i holds a map point that is
  // not bad, featvec_kf = pKF->mFeatVec; keys_f = F.mvKeys, desc_f = F.mDescriptors, featvec_f = F.mFeatVec
  // (DBoW2::FeatureVector is this std::map).  match_of_f[i] = index of the key-frame feature whose map point
  // vpMapPointMatches[i] becomes (-1 = NULL); returns nmatches.  At most 2048 features per side.
  int SearchByBoW(const std::vector<cv::KeyPoint>& keys_kf, const cv::Mat& desc_kf, const std::vector<uint8_t>& has_point_kf,
                  const std::map<unsigned int, std::vector<unsigned int> >& featvec_kf, const std::vector<cv::KeyPoint>& keys_f,
                  const cv::Mat& desc_f, const std::map<unsigned int, std::vector<unsigned int> >& featvec_f, float nnratio,
                  bool check_orientation, std::vector<int>& match_of_f);

  // ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (orb_matcher.cc:697-815, loop closing): both
  // sides need a good map point; match_of_1[i] = feature of pKF2 whose map point vpMatches12[i] becomes (-1 = NULL).
  int SearchByBoW(const std::vector<cv::KeyPoint>& keys_un1, const cv::Mat& desc1, const std::vector<uint8_t>& has_point1,
                  const std::map<unsigned int, std::vector<unsigned int> >& featvec1, const std::vector<cv::KeyPoint>& keys_un2,
                  const cv::Mat& desc2, const std::vector<uint8_t>& has_point2,
                  const std::map<unsigned int, std::vector<unsigned int> >& featvec2, float nnratio, bool check_orientation,
                  std::vector<int>& match_of_1);

  // ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (orb_matcher.cc:817-1040) for key
  // frames with one pinhole camera: keys_un / desc / has_point (GetMapPoint(i) != NULL) / u_right (mvuRight) / featvec of
  // both key frames, f12 = the F12 of Pinhole::EpipolarConstrain (pinhole_model.cc:116-119), row-major, epipole =
  // pKF2->cam_->Project(T2w * Cw) (:829-830), scale_factors / level_sigma2 = mvScaleFactors / mvLevelSigma2.
  int SearchForTriangulation(const std::vector<cv::KeyPoint>& keys_un1, const cv::Mat& desc1, const std::vector<uint8_t>& has_point1,
                             const std::vector<float>& u_right1, const std::map<unsigned int, std::vector<unsigned int> >& featvec1,
                             const std::vector<cv::KeyPoint>& keys_un2, const cv::Mat& desc2, const std::vector<uint8_t>& has_point2,
                             const std::vector<float>& u_right2, const std::map<unsigned int, std::vector<unsigned int> >& featvec2,
                             const float f12[9], float epipole_x, float epipole_y, const std::vector<float>& scale_factors,
                             const std::vector<float>& level_sigma2, bool only_stereo, bool coarse, bool check_orientation,
                             std::vector<std::pair<size_t, size_t> >& matched_pairs);

  struct Window { float u, v, r; int min_level, max_level; };
  struct WindowBest { int best_dist, best_idx, best_level, best_dist2, best_level2; };
  // orb_matcher.cc:66-113 with Frame::GetFeaturesInArea (frame.cc:679-746) for a batch of projections.
  void WindowSearch(const std::vector<cv::KeyPoint>& keys_un, const cv::Mat& desc, float min_x, float min_y,
                    float grid_inv_w, float grid_inv_h, int grid_cols, int grid_rows, const std::vector<Window>& windows,
                    const cv::Mat& window_desc, const std::vector<uint8_t>* already_matched, std::vector<WindowBest>& out,
                    // stereo gate of orb_matcher.cc:89-92 / 1586-1590: Frame::mvuRight, and per window the projected
                    // right coordinate and the largest accepted |difference| (all three or none)
                    const std::vector<float>* u_right = nullptr, const std::vector<float>* window_u_right = nullptr,
                    const std::vector<float>* window_max_err = nullptr);

  orbm_matcher* handle() { return m_; }  // for callers that go to the C ABI directly (cpp/src/orb_matcher.cc)

 private:
  int SearchByBoWImpl(bool keyframes, const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1,
                      const std::vector<uint8_t>& has_point1, const std::map<unsigned int, std::vector<unsigned int> >& featvec1,
                      const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2, const std::vector<uint8_t>* has_point2,
                      const std::map<unsigned int, std::vector<unsigned int> >& featvec2, float nnratio, bool check_orientation,
                      std::vector<int>& match);
  orbm_matcher* m_;
};

}  // namespace ORB_SLAM_FUSION

#endif
