// ORB_SLAM_FUSION::ORBmatcher over the B200 C ABI (include/orbx.h).  Replaces the reference's
// src/cam/orb_feature/orb_matcher.cc (CMakeLists.txt:81); see the header for what stays source-compatible.
//
// Every method has the same shape: (1) walk the reference's entry tests on the host, in the reference's order,
// and gather what the search reads -- window (u, v, r, level gate), descriptor, flags -- into plain arrays;
// (2) ONE call into liborbx_b200.so runs the search for all of them on the GPU, greedy claims and the rotation
// histogram included where the reference has them; (3) scatter the returned indices into the callers' pointer
// vectors and do the map bookkeeping (Replace / AddObservation) in the reference's order.  Line numbers cite
// the reference's orb_matcher.cc.
#include "cam/orb_feature/orb_matcher.h"

#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

#include "cam/orb_feature/orb_matcher_gpu.h"
#include "orbx.h"

using std::vector;

namespace ORB_SLAM_FUSION {

const int ORBmatcher::TH_HIGH = 100;  // :35-37
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

namespace {

// One GPU matcher per host thread (tracking, local mapping and loop closing each run their own matchers
// concurrently; a handle is not re-entrant), created on first use and kept for the life of the thread.
ORBmatcherGpu &gpu() {
  static thread_local std::unique_ptr<ORBmatcherGpu> g;
  if (!g) {
    const char *dev = std::getenv("ORBX_DEVICE");
    g.reset(new ORBmatcherGpu(dev ? std::atoi(dev) : 0));
  }
  return *g;
}

void check(int rc, const char *what) {
  if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBmatcher::") + what + ": " + orbm_last_error(gpu().handle()));
}

[[noreturn]] void two_camera(const char *what) {
  throw std::runtime_error(std::string("ORBmatcher::") + what +
                           ": two-camera frames (Nleft != -1) are not supported by the GPU matcher");
}

orbm_grid_geom frame_grid() {  // frame.cc:199-217; the members are static
  const orbm_grid_geom g = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv,
                            FRAME_GRID_COLS, FRAME_GRID_ROWS};
  return g;
}
orbm_grid_geom keyframe_grid(const KeyFrame *kf) {  // keyframe.cc:729-773 reads these
  const orbm_grid_geom g = {(float)kf->mnMinX, (float)kf->mnMinY, kf->mfGridElementWidthInv, kf->mfGridElementHeightInv,
                            (int32_t)kf->mnGridCols, (int32_t)kf->mnGridRows};
  return g;
}

vector<uint8_t> dense(const cv::Mat &m, int n) {
  vector<uint8_t> v((size_t)(n > 0 ? n : 0) * 32);
  for (int i = 0; i < n; ++i) std::memcpy(v.data() + 32 * (size_t)i, m.ptr(i), 32);
  return v;
}

// the projected map points of one call, in the reference's visiting order
struct Windows {
  vector<int> origin;  // index into the caller's point vector
  vector<orbm_window_query> q;
  vector<uint8_t> desc;
  vector<float> angle, u_right, max_err;
  void add(int from, float u, float v, float r, int min_level, int max_level, MapPoint *mp, float ang = 0.f, float ur = 0.f) {
    const orbm_window_query w = {u, v, r, min_level, max_level};
    origin.push_back(from);
    q.push_back(w);
    const cv::Mat d = mp->GetDescriptor();
    desc.insert(desc.end(), d.ptr(0), d.ptr(0) + 32);
    angle.push_back(ang);
    u_right.push_back(ur);
    max_err.push_back(r);
  }
  int size() const { return (int)q.size(); }
};

// orbm_search_by_projection_last over explicit windows: nearest free keypoint within th_high per window, in
// order, claims blocking the later windows, optional rotation histogram (:1706-1725 and its siblings).
int claim_nearest(const char *what, const vector<cv::KeyPoint> &keys, const cv::Mat &descriptors, const orbm_grid_geom &g,
                  const Windows &w, const vector<uint8_t> &taken, const vector<float> *kp_u_right, int th_high, bool orientation,
                  vector<int32_t> &assigned) {
  const int n = (int)keys.size();
  assigned.assign((size_t)(n > 0 ? n : 1), -1);
  if (w.size() == 0 || n == 0) return 0;
  const vector<uint8_t> d = dense(descriptors, n);
  int32_t nm = 0;
  check(orbm_search_by_projection_last(gpu().handle(), reinterpret_cast<const orbx_kp *>(keys.data()), d.data(), n, &g, w.q.data(),
                                       w.desc.data(), w.angle.data(), w.size(), taken.data(), kp_u_right ? kp_u_right->data() : nullptr,
                                       kp_u_right ? w.u_right.data() : nullptr, kp_u_right ? w.max_err.data() : nullptr, th_high,
                                       orientation ? 1 : 0, assigned.data(), &nm, ORBX_MEM_HOST, nullptr),
        what);
  return nm;
}

// best keypoint per window, windows independent (no claims)
void nearest(const char *what, const vector<cv::KeyPoint> &keys, const cv::Mat &descriptors, const orbm_grid_geom &g, const Windows &w,
             vector<orbm_window_result> &out) {
  const int n = (int)keys.size();
  out.resize((size_t)w.size());
  if (w.size() == 0) return;
  const vector<uint8_t> d = dense(descriptors, n);
  check(orbm_window_search(gpu().handle(), reinterpret_cast<const orbx_kp *>(keys.data()), d.data(), n, &g, w.q.data(), w.desc.data(),
                           w.size(), nullptr, out.data(), ORBX_MEM_HOST, nullptr),
        what);
}

void require_observed(MapPoint *mp, const char *what) {
  // A claim blocks the later map points of the call only when the claiming point has observations (:86-87, :1581-1583).
  // Points without (the temporal points of localisation-only tracking) would need the reference's overwrite semantics.
  if (mp->Observations() <= 0)
    throw std::runtime_error(std::string("ORBmatcher::") + what + ": map point without observations (localisation-only mode) is not supported");
}

}  // namespace

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {  // :1877-1891
  const uint8_t *pa = a.ptr<uint8_t>(), *pb = b.ptr<uint8_t>();
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, pa + 4 * i, 4);
    std::memcpy(&y, pb + 4 * i, 4);
    dist += __builtin_popcount(x ^ y);
  }
  return dist;
}

float ORBmatcher::RadiusByViewingCos(const float &viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }  // :208-213

void ORBmatcher::ComputeThreeMaxima(vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3) {  // :1841-1873
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// ---- :42-134 -------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint *> &vpMapPoints, const float th, const bool bFarPoints,
                                   const float thFarPoints) {
  if (F.Nleft != -1) two_camera("SearchByProjection(Frame, MapPoints)");
  vector<ORBmatcherGpu::TrackedPoint> pts;
  vector<int> origin;
  cv::Mat pdesc;
  vector<uint8_t> pd;
  for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
    MapPoint *pMP = vpMapPoints[iMP];
    if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;  // :52
    if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;  // :54
    if (pMP->isBad()) continue;                                  // :56
    if (!pMP->mbTrackInView) continue;                           // :58 (the right-camera half belongs to Nleft != -1)
    require_observed(pMP, "SearchByProjection(Frame, MapPoints)");
    const ORBmatcherGpu::TrackedPoint t = {pMP->mTrackProjX,     pMP->mTrackProjY,       pMP->mTrackProjXR, pMP->mTrackViewCos,
                                           pMP->mTrackDepth,     pMP->mnTrackScaleLevel, true,              false};
    pts.push_back(t);
    origin.push_back((int)iMP);
    const cv::Mat d = pMP->GetDescriptor();
    pd.insert(pd.end(), d.ptr(0), d.ptr(0) + 32);
  }
  if (pts.empty()) return 0;
  pdesc = cv::Mat((int)pts.size(), 32, CV_8U, pd.data());
  vector<uint8_t> taken((size_t)F.N, 0);
  for (int i = 0; i < F.N; ++i) taken[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;  // :86-87
  const orbm_grid_geom g = frame_grid();
  vector<int> assigned;
  // far-point / in-view / bad tests are done above, so the gather of ORBmatcherGpu sees only accepted points
  const int nm = gpu().SearchByProjection(F.mvKeysUn, F.mDescriptors, F.mvuRight, F.mvScaleFactors, g.min_x, g.min_y, g.inv_w, g.inv_h,
                                          g.cols, g.rows, pts, pdesc, taken, th, false, 0.f, mfNNratio, assigned);
  for (int i = 0; i < F.N; ++i)
    if (assigned[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[origin[assigned[i]]];  // :121
  return nm;
}

// ---- :1518-1728 ------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono) {
  if (CurrentFrame.Nleft != -1 || LastFrame.Nleft != -1) two_camera("SearchByProjection(CurrentFrame, LastFrame)");
  const Sophus::SE3f Tcw = CurrentFrame.GetPose();
  const Eigen::Vector3f twc = Tcw.inverse().translation();
  const Sophus::SE3f Tlw = LastFrame.GetPose();
  const Eigen::Vector3f tlc = Tlw * twc;
  const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;   // :1535-1536
  const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;

  Windows w;
  for (int i = 0; i < LastFrame.N; i++) {
    MapPoint *pMP = LastFrame.mvpMapPoints[i];
    if (!pMP || LastFrame.mvbOutlier[i]) continue;
    const Eigen::Vector3f x3Dw = pMP->GetWorldPos();
    const Eigen::Vector3f x3Dc = Tcw * x3Dw;
    const float invzc = 1.0 / x3Dc(2);
    if (invzc < 0) continue;
    const Eigen::Vector2f uv = CurrentFrame.cam_->Project(x3Dc);
    if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
    if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
    const int nLastOctave = LastFrame.mvKeys[i].octave;
    const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];  // :1566
    require_observed(pMP, "SearchByProjection(CurrentFrame, LastFrame)");
    // :1568-1576: forward motion looks at the same or finer... coarser levels only, backward at the finer ones
    const int lo = bForward ? nLastOctave : (bBackward ? 0 : nLastOctave - 1);
    const int hi = bForward ? -1 : (bBackward ? nLastOctave : nLastOctave + 1);
    w.add(i, uv(0), uv(1), radius, lo, hi, pMP, LastFrame.mvKeysUn[i].angle, uv(0) - CurrentFrame.bf_ * invzc);  // :1587
  }
  vector<uint8_t> taken((size_t)CurrentFrame.N, 0);
  for (int i = 0; i < CurrentFrame.N; ++i)
    taken[i] = CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0;  // :1581-1583
  vector<int32_t> assigned;
  const int nm = claim_nearest("SearchByProjection(CurrentFrame, LastFrame)", CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, frame_grid(),
                               w, taken, &CurrentFrame.mvuRight, TH_HIGH, mbCheckOrientation, assigned);
  for (int i = 0; i < CurrentFrame.N; ++i)
    if (assigned[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[w.origin[assigned[i]]];  // :1611
  return nm;
}

// ---- :1730-1840 ------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th,
                                   const int ORBdist) {
  if (CurrentFrame.Nleft != -1) two_camera("SearchByProjection(CurrentFrame, KeyFrame)");
  const Sophus::SE3f Tcw = CurrentFrame.GetPose();
  const Eigen::Vector3f Ow = Tcw.inverse().translation();
  const vector<MapPoint *> vpMPs = pKF->GetMapPointMatches();
  Windows w;
  for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
    MapPoint *pMP = vpMPs[i];
    if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
    const Eigen::Vector3f x3Dw = pMP->GetWorldPos();
    const Eigen::Vector3f x3Dc = Tcw * x3Dw;
    const Eigen::Vector2f uv = CurrentFrame.cam_->Project(x3Dc);
    if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
    if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
    const Eigen::Vector3f PO = x3Dw - Ow;
    const float dist3D = PO.norm();
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
    const float radius = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
    w.add((int)i, uv(0), uv(1), radius, nPredictedLevel - 1, nPredictedLevel + 1, pMP, pKF->mvKeysUn[i].angle);  // :1776-1777, :1806
  }
  vector<uint8_t> taken((size_t)CurrentFrame.N, 0);
  for (int i = 0; i < CurrentFrame.N; ++i) taken[i] = CurrentFrame.mvpMapPoints[i] != nullptr;  // :1791
  vector<int32_t> assigned;
  const int nm = claim_nearest("SearchByProjection(CurrentFrame, KeyFrame)", CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, frame_grid(), w,
                               taken, nullptr, ORBdist, mbCheckOrientation, assigned);
  for (int i = 0; i < CurrentFrame.N; ++i)
    if (assigned[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[w.origin[assigned[i]]];
  return nm;
}

// ---- :391-488 and :490-596 ---------------------------------------------------------------------------------------
namespace {
int sim3_projection(const char *what, KeyFrame *pKF, Sophus::Sim3f &Scw, const vector<MapPoint *> &vpPoints,
                    const vector<KeyFrame *> *vpPointsKFs, vector<MapPoint *> &vpMatched, vector<KeyFrame *> *vpMatchedKF, int th,
                    float ratioHamming, int th_low) {
  if (pKF->NLeft != -1) two_camera(what);
  const Sophus::SE3f Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
  const Eigen::Vector3f Ow = Tcw.inverse().translation();
  std::set<MapPoint *> spAlreadyFound(vpMatched.begin(), vpMatched.end());
  spAlreadyFound.erase(static_cast<MapPoint *>(NULL));
  Windows w;
  for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {
    MapPoint *pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    const Eigen::Vector3f p3Dw = pMP->GetWorldPos();
    const Eigen::Vector3f p3Dc = Tcw * p3Dw;
    if (p3Dc(2) < 0.0) continue;
    const Eigen::Vector2f uv = pKF->cam_->Project(p3Dc);
    if (!pKF->IsInImage(uv(0), uv(1))) continue;
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    const Eigen::Vector3f PO = p3Dw - Ow;
    const float dist = PO.norm();
    if (dist < minDistance || dist > maxDistance) continue;
    const Eigen::Vector3f Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist) continue;
    const int nPredictedLevel = pMP->PredictScale(dist, pKF);
    const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
    // KeyFrame::GetFeaturesInArea has no level gate; the loop's own test (:465) is the gate [level-1, level]
    w.add(iMP, uv(0), uv(1), radius, nPredictedLevel - 1, nPredictedLevel, pMP);
  }
  const int n = (int)pKF->mvKeysUn.size();
  vector<uint8_t> taken((size_t)n, 0);
  for (int i = 0; i < n && i < (int)vpMatched.size(); ++i) taken[i] = vpMatched[i] != nullptr;  // :461
  // bestDist <= TH_LOW * ratioHamming (:483) on integers == bestDist <= floor of the float product
  const int th_high = (int)std::floor((float)th_low * ratioHamming);
  vector<int32_t> assigned;
  const int nm = claim_nearest(what, pKF->mvKeysUn, pKF->mDescriptors, keyframe_grid(pKF), w, taken, nullptr, th_high, false, assigned);
  for (int i = 0; i < n; ++i)
    if (assigned[i] >= 0) {
      vpMatched[i] = vpPoints[w.origin[assigned[i]]];
      if (vpMatchedKF) (*vpMatchedKF)[i] = (*vpPointsKFs)[w.origin[assigned[i]]];
    }
  return nm;
}
}  // namespace

int ORBmatcher::SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Scw, const vector<MapPoint *> &vpPoints,
                                   vector<MapPoint *> &vpMatched, int th, float ratioHamming) {
  return sim3_projection("SearchByProjection(KeyFrame, Sim3)", pKF, Scw, vpPoints, nullptr, vpMatched, nullptr, th, ratioHamming, TH_LOW);
}

int ORBmatcher::SearchByProjection(KeyFrame *pKF, Sophus::Sim3<float> &Scw, const vector<MapPoint *> &vpPoints,
                                   const vector<KeyFrame *> &vpPointsKFs, vector<MapPoint *> &vpMatched,
                                   vector<KeyFrame *> &vpMatchedKF, int th, float ratioHamming) {
  return sim3_projection("SearchByProjection(KeyFrame, Sim3, KFs)", pKF, Scw, vpPoints, &vpPointsKFs, vpMatched, &vpMatchedKF, th,
                         ratioHamming, TH_LOW);
}

// ---- :215-389 --------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint *> &vpMapPointMatches) {
  if (F.Nleft != -1) two_camera("SearchByBoW(KeyFrame, Frame)");
  const vector<MapPoint *> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = vector<MapPoint *>(F.N, static_cast<MapPoint *>(NULL));
  vector<uint8_t> has((size_t)pKF->mvKeysUn.size(), 0);
  for (size_t i = 0; i < has.size() && i < vpMapPointsKF.size(); ++i) has[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();  // :246-250
  vector<int> match_of_f;
  const int nm = gpu().SearchByBoW(pKF->mvKeysUn, pKF->mDescriptors, has, pKF->mFeatVec, F.mvKeys, F.mDescriptors, F.mFeatVec, mfNNratio,
                                   mbCheckOrientation, match_of_f);
  for (int i = 0; i < F.N && i < (int)match_of_f.size(); ++i)
    if (match_of_f[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[match_of_f[i]];
  return nm;
}

// ---- :697-815 --------------------------------------------------------------------------------------------------
int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12) {
  if (pKF1->NLeft != -1 || pKF2->NLeft != -1) two_camera("SearchByBoW(KeyFrame, KeyFrame)");
  const vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
  vpMatches12 = vector<MapPoint *>(vpMapPoints1.size(), static_cast<MapPoint *>(NULL));
  vector<uint8_t> has1(pKF1->mvKeysUn.size(), 0), has2(pKF2->mvKeysUn.size(), 0);
  for (size_t i = 0; i < has1.size() && i < vpMapPoints1.size(); ++i) has1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();
  for (size_t i = 0; i < has2.size() && i < vpMapPoints2.size(); ++i) has2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
  vector<int> match_of_1;
  const int nm = gpu().SearchByBoW(pKF1->mvKeysUn, pKF1->mDescriptors, has1, pKF1->mFeatVec, pKF2->mvKeysUn, pKF2->mDescriptors, has2,
                                   pKF2->mFeatVec, mfNNratio, mbCheckOrientation, match_of_1);
  for (size_t i = 0; i < vpMatches12.size() && i < match_of_1.size(); ++i)
    if (match_of_1[i] >= 0) vpMatches12[i] = vpMapPoints2[match_of_1[i]];
  return nm;
}

// ---- :597-695 (host: every match depends on the running best distance of its keypoint, vMatchedDistance) ----------
int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12,
                                        int windowSize) {
  int nmatches = 0;
  const size_t n1 = F1.mvKeysUn.size(), n2 = F2.mvKeysUn.size();
  vnMatches12.assign(n1, -1);
  vector<int> bins[30];
  const float factor = HISTO_LENGTH / 360.0f;
  vector<int> held_dist(n2, INT_MAX), back(n2, -1);
  for (size_t i1 = 0; i1 < n1; i1++) {
    if (F1.mvKeysUn[i1].octave > 0) continue;  // level 0 only
    const vector<size_t> cand = F2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, 0, 0);
    int d_best = INT_MAX, d_second = INT_MAX, i_best = -1;
    const cv::Mat d1 = F1.mDescriptors.row((int)i1);
    for (size_t c = 0; c < cand.size(); ++c) {
      const int d = DescriptorDistance(d1, F2.mDescriptors.row((int)cand[c]));
      if (held_dist[cand[c]] <= d) continue;
      if (d < d_best) { d_second = d_best; d_best = d; i_best = (int)cand[c]; }
      else if (d < d_second) d_second = d;
    }
    if (d_best > TH_LOW || !(d_best < (float)d_second * mfNNratio)) continue;
    if (back[i_best] >= 0) { vnMatches12[back[i_best]] = -1; nmatches--; }
    vnMatches12[i1] = i_best;
    back[i_best] = (int)i1;
    held_dist[i_best] = d_best;
    nmatches++;
    if (mbCheckOrientation) {
      float rot = F1.mvKeysUn[i1].angle - F2.mvKeysUn[i_best].angle;
      if (rot < 0.0) rot += 360.0f;
      int bin = (int)round(rot * factor);
      if (bin == HISTO_LENGTH) bin = 0;
      bins[bin].push_back((int)i1);
    }
  }
  if (mbCheckOrientation) {
    int a = -1, b = -1, c = -1;
    ComputeThreeMaxima(bins, HISTO_LENGTH, a, b, c);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == a || i == b || i == c) continue;
      for (size_t j = 0; j < bins[i].size(); j++)
        if (vnMatches12[bins[i][j]] >= 0) { vnMatches12[bins[i][j]] = -1; nmatches--; }
    }
  }
  for (size_t i1 = 0; i1 < n1; i1++)
    if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
  return nmatches;
}

// ---- :817-1040 -------------------------------------------------------------------------------------------------
// F12 of Pinhole::EpipolarConstrain (pinhole_model.cc:116-119): K1^-T [t12]x R12 K2^-1, row-major.
#ifndef ORBM_FACADE_CUSTOM_F12
static void FundamentalMatrix(KeyFrame *pKF1, KeyFrame *pKF2, const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12, float f12[9]) {
  const Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
  const Eigen::Matrix3f K1 = pKF1->cam_->ToKEig(), K2 = pKF2->cam_->ToKEig();
  const Eigen::Matrix3f F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) f12[3 * r + c] = F12(r, c);
}
#else
void FundamentalMatrix(KeyFrame *pKF1, KeyFrame *pKF2, const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12, float f12[9]);
#endif

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, vector<std::pair<size_t, size_t> > &vMatchedPairs,
                                       const bool bOnlyStereo, const bool bCoarse) {
  if (pKF1->NLeft != -1 || pKF2->NLeft != -1 || pKF1->cam2_ || pKF2->cam2_) two_camera("SearchForTriangulation");
  const Sophus::SE3f T1w = pKF1->GetPose(), T2w = pKF2->GetPose(), Tw2 = pKF2->GetPoseInverse();
  const Eigen::Vector3f Cw = pKF1->GetCameraCenter();
  const Eigen::Vector3f C2 = T2w * Cw;
  const Eigen::Vector2f ep = pKF2->cam_->Project(C2);  // :829-830
  const Sophus::SE3f T12 = T1w * Tw2;
  float f12[9];
  FundamentalMatrix(pKF1, pKF2, T12.rotationMatrix(), T12.translation(), f12);
  const int n1 = (int)pKF1->mvKeysUn.size(), n2 = (int)pKF2->mvKeysUn.size();
  vector<uint8_t> has1((size_t)n1, 0), has2((size_t)n2, 0);
  for (int i = 0; i < n1; ++i) has1[i] = pKF1->GetMapPoint(i) != nullptr;  // :883-886
  for (int i = 0; i < n2; ++i) has2[i] = pKF2->GetMapPoint(i) != nullptr;  // :911
  return gpu().SearchForTriangulation(pKF1->mvKeysUn, pKF1->mDescriptors, has1, pKF1->mvuRight, pKF1->mFeatVec, pKF2->mvKeysUn,
                                      pKF2->mDescriptors, has2, pKF2->mvuRight, pKF2->mFeatVec, f12, ep(0), ep(1), pKF1->mvScaleFactors,
                                      pKF2->mvLevelSigma2, bOnlyStereo, bCoarse, mbCheckOrientation, vMatchedPairs);
}

// ---- :1320-1516 ------------------------------------------------------------------------------------------------
int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12, const Sophus::Sim3f &S12, const float th) {
  if (pKF1->NLeft != -1 || pKF2->NLeft != -1) two_camera("SearchBySim3");
  const float &fx = pKF1->fx, &fy = pKF1->fy, &cx = pKF1->cx, &cy = pKF1->cy;
  const Sophus::SE3f T1w = pKF1->GetPose(), T2w = pKF2->GetPose();
  const Sophus::Sim3f S21 = S12.inverse();
  const vector<MapPoint *> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
  const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
  vector<bool> done1(N1, false), done2(N2, false);
  for (int i = 0; i < N1; i++) {
    MapPoint *pMP = vpMatches12[i];
    if (!pMP) continue;
    done1[i] = true;
    const int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
    if (idx2 >= 0 && idx2 < N2) done2[idx2] = true;
  }
  // the two directions: points of one key frame projected into the other through the similarity
  auto project = [&](const vector<MapPoint *> &pts, const vector<bool> &done, const Sophus::SE3f &Tsrc, const Sophus::Sim3f &S, KeyFrame *into,
                     Windows &w) {
    for (int i = 0; i < (int)pts.size(); i++) {
      MapPoint *pMP = pts[i];
      if (!pMP || done[i] || pMP->isBad()) continue;
      const Eigen::Vector3f p3Dw = pMP->GetWorldPos();
      const Eigen::Vector3f p_src = Tsrc * p3Dw;
      const Eigen::Vector3f p = S * p_src;
      if (p(2) < 0.0) continue;
      const float invz = 1.0 / p(2);
      const float x = p(0) * invz, y = p(1) * invz;
      const float u = fx * x + cx, v = fy * y + cy;
      if (!into->IsInImage(u, v)) continue;
      const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
      const float dist3D = p.norm();
      if (dist3D < minDistance || dist3D > maxDistance) continue;
      const int nPredictedLevel = pMP->PredictScale(dist3D, into);
      const float radius = th * into->mvScaleFactors[nPredictedLevel];
      w.add(i, u, v, radius, nPredictedLevel - 1, nPredictedLevel, pMP);
    }
  };
  Windows w12, w21;
  project(vpMapPoints1, done1, T1w, S21, pKF2, w12);
  project(vpMapPoints2, done2, T2w, S12, pKF1, w21);
  vector<orbm_window_result> r12, r21;
  nearest("SearchBySim3", pKF2->mvKeysUn, pKF2->mDescriptors, keyframe_grid(pKF2), w12, r12);
  nearest("SearchBySim3", pKF1->mvKeysUn, pKF1->mDescriptors, keyframe_grid(pKF1), w21, r21);
  vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
  for (int k = 0; k < w12.size(); ++k)
    if (r12[k].best_idx >= 0 && r12[k].best_dist <= TH_HIGH) vnMatch1[w12.origin[k]] = r12[k].best_idx;
  for (int k = 0; k < w21.size(); ++k)
    if (r21[k].best_idx >= 0 && r21[k].best_dist <= TH_HIGH) vnMatch2[w21.origin[k]] = r21[k].best_idx;
  int nFound = 0;
  for (int i1 = 0; i1 < N1; i1++) {  // :1500-1513: keep the mutual matches
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && vnMatch2[idx2] == i1) {
      vpMatches12[i1] = vpMapPoints2[idx2];
      nFound++;
    }
  }
  return nFound;
}

// ---- :1042-1212 ------------------------------------------------------------------------------------------------
int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th, const bool bRight) {
  if (bRight || pKF->NLeft != -1) two_camera("Fuse");
  const Sophus::SE3f Tcw = pKF->GetPose();
  const Eigen::Vector3f Ow = pKF->GetCameraCenter();
  const float &bf = pKF->bf_;
  // The search of a map point does not depend on what the loop did for the points before it; only the tests on the
  // map point itself (bad, already in the key frame: Replace / AddObservation of an earlier iteration can change both)
  // and the bookkeeping do.  So: search for every point that passes the state-free tests now, re-test and book-keep in order.
  Windows w;
  const int nMPs = (int)vpMapPoints.size();
  for (int i = 0; i < nMPs; i++) {
    MapPoint *pMP = vpMapPoints[i];
    if (!pMP) continue;
    const Eigen::Vector3f p3Dw = pMP->GetWorldPos();
    const Eigen::Vector3f p3Dc = Tcw * p3Dw;
    if (p3Dc(2) < 0.0f) continue;
    const float invz = 1 / p3Dc(2);
    const Eigen::Vector2f uv = pKF->cam_->Project(p3Dc);
    if (!pKF->IsInImage(uv(0), uv(1))) continue;
    const float ur = uv(0) - bf * invz;
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    const Eigen::Vector3f PO = p3Dw - Ow;
    const float dist3D = PO.norm();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const Eigen::Vector3f Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist3D) continue;
    const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
    const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
    w.add(i, uv(0), uv(1), radius, nPredictedLevel - 1, nPredictedLevel, pMP, 0.f, ur);
  }
  const int n = (int)pKF->mvKeysUn.size();
  vector<orbm_window_result> res((size_t)w.size());
  if (w.size() > 0) {
    const vector<uint8_t> d = dense(pKF->mDescriptors, n);
    const orbm_grid_geom g = keyframe_grid(pKF);
    check(orbm_window_search_fuse(gpu().handle(), reinterpret_cast<const orbx_kp *>(pKF->mvKeysUn.data()), d.data(), n, &g, w.q.data(),
                                  w.desc.data(), w.size(), pKF->mvuRight.data(), w.u_right.data(), pKF->mvInvLevelSigma2.data(),
                                  (int)pKF->mvInvLevelSigma2.size(), res.data(), ORBX_MEM_HOST, nullptr),
          "Fuse");
  }
  int nFused = 0;
  for (int k = 0; k < w.size(); ++k) {
    MapPoint *pMP = vpMapPoints[w.origin[k]];
    if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;  // :1078-1084, evaluated when the reference would
    if (res[k].best_idx < 0 || res[k].best_dist > TH_LOW) continue;
    const int bestIdx = res[k].best_idx;
    MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx);  // :1191-1206
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

// ---- :1214-1318 ------------------------------------------------------------------------------------------------
int ORBmatcher::Fuse(KeyFrame *pKF, Sophus::Sim3f &Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint) {
  if (pKF->NLeft != -1) two_camera("Fuse(Sim3)");
  const Sophus::SE3f Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale());
  const Eigen::Vector3f Ow = Tcw.inverse().translation();
  const std::set<MapPoint *> spAlreadyFound = pKF->GetMapPoints();
  Windows w;
  const int nPoints = (int)vpPoints.size();
  for (int iMP = 0; iMP < nPoints; iMP++) {
    MapPoint *pMP = vpPoints[iMP];
    if (spAlreadyFound.count(pMP)) continue;
    const Eigen::Vector3f p3Dw = pMP->GetWorldPos();
    const Eigen::Vector3f p3Dc = Tcw * p3Dw;
    if (p3Dc(2) < 0.0f) continue;
    const Eigen::Vector2f uv = pKF->cam_->Project(p3Dc);
    if (!pKF->IsInImage(uv(0), uv(1))) continue;
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    const Eigen::Vector3f PO = p3Dw - Ow;
    const float dist3D = PO.norm();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const Eigen::Vector3f Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist3D) continue;
    const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
    const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
    w.add(iMP, uv(0), uv(1), radius, nPredictedLevel - 1, nPredictedLevel, pMP);
  }
  vector<orbm_window_result> res;
  nearest("Fuse(Sim3)", pKF->mvKeysUn, pKF->mDescriptors, keyframe_grid(pKF), w, res);
  int nFused = 0;
  for (int k = 0; k < w.size(); ++k) {
    MapPoint *pMP = vpPoints[w.origin[k]];
    if (pMP->isBad()) continue;  // :1240, evaluated in order (nothing in this loop makes a point bad, but stay literal)
    if (res[k].best_idx < 0 || res[k].best_dist > TH_LOW) continue;
    const int bestIdx = res[k].best_idx;
    MapPoint *pMPinKF = pKF->GetMapPoint(bestIdx);  // :1304-1312
    if (pMPinKF) {
      if (!pMPinKF->isBad()) vpReplacePoint[w.origin[k]] = pMPinKF;
    } else {
      pMP->AddObservation(pKF, bestIdx);
      pKF->AddMapPoint(pMP, bestIdx);
    }
    nFused++;
  }
  return nFused;
}

}  // namespace ORB_SLAM_FUSION
