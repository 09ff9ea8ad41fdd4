#include "cam/orb_feature/orb_matcher_gpu.h"

#include <climits>
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orbx.h"

namespace ORB_SLAM_FUSION {

const int ORBmatcherGpu::TH_HIGH;  // orb_matcher.cc:35-37
const int ORBmatcherGpu::TH_LOW;
const int ORBmatcherGpu::HISTO_LENGTH;

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_kp), "cv::KeyPoint layout");
static_assert(sizeof(ORBmatcherGpu::Window) == sizeof(orbm_window_query), "window layout");
static_assert(sizeof(ORBmatcherGpu::WindowBest) == sizeof(orbm_window_result), "result layout");

namespace {
// descriptors as one dense n x 32 block (cv::Mat rows may be strided views)
std::vector<uint8_t> dense_rows(const cv::Mat& m) {
  std::vector<uint8_t> v((size_t)m.rows * 32);
  for (int i = 0; i < m.rows; ++i) std::memcpy(v.data() + 32 * (size_t)i, m.ptr(i), 32);
  return v;
}
void check(orbm_matcher* m, int rc) {
  if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBmatcherGpu: ") + orbm_last_error(m));
}
}  // namespace

ORBmatcherGpu::ORBmatcherGpu(int device) : m_(nullptr) {
  const int rc = orbm_create(device, &m_);
  if (rc != ORBX_OK) throw std::runtime_error("ORBmatcherGpu: orbm_create failed with code " + std::to_string(rc));
}

ORBmatcherGpu::~ORBmatcherGpu() { orbm_destroy(m_); }

std::vector<int> ORBmatcherGpu::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
  const std::vector<uint8_t> da = dense_rows(a), db = dense_rows(b);
  std::vector<int> out((size_t)a.rows);
  check(m_, orbm_hamming_pairs(m_, da.data(), db.data(), a.rows, out.data(), ORBX_MEM_HOST, nullptr));
  return out;
}

void ORBmatcherGpu::KnnMatch2(const cv::Mat& query, const cv::Mat& train, std::vector<std::vector<cv::DMatch> >& matches) {
  const std::vector<uint8_t> q = dense_rows(query), t = dense_rows(train);
  std::vector<int64_t> idx((size_t)query.rows * 2);
  std::vector<int32_t> dist((size_t)query.rows * 2);
  check(m_, orbm_knn2(m_, q.data(), query.rows, t.data(), train.rows, 0, idx.data(), dist.data(), ORBX_MEM_HOST, nullptr));
  matches.assign((size_t)query.rows, std::vector<cv::DMatch>());
  for (int i = 0; i < query.rows; ++i)
    for (int k = 0; k < 2; ++k)
      if (idx[2 * i + k] >= 0) matches[i].push_back(cv::DMatch(i, (int)idx[2 * i + k], (float)dist[2 * i + k]));
}

void ORBmatcherGpu::StereoRowBand(const std::vector<cv::KeyPoint>& kl, const cv::Mat& dl, const std::vector<cv::KeyPoint>& kr,
                                  const cv::Mat& dr, const std::vector<float>& sf, int n_rows, float min_d, float max_d,
                                  std::vector<int>& best_idx_right, std::vector<int>& best_dist) {
  const std::vector<uint8_t> l = dense_rows(dl), r = dense_rows(dr);
  best_idx_right.assign(kl.size(), -1);
  best_dist.assign(kl.size(), TH_HIGH);
  check(m_, orbm_stereo_rowband(m_, reinterpret_cast<const orbx_kp*>(kl.data()), l.data(), (int)kl.size(),
                                reinterpret_cast<const orbx_kp*>(kr.data()), r.data(), (int)kr.size(), sf.data(),
                                (int)sf.size(), n_rows, min_d, max_d, best_idx_right.data(), best_dist.data(),
                                ORBX_MEM_HOST, nullptr));
}

void ORBmatcherGpu::ComputeStereoMatches(orbx_extractor* left, orbx_extractor* right, const std::vector<cv::KeyPoint>& kl,
                                         const cv::Mat& dl, const std::vector<cv::KeyPoint>& kr, const cv::Mat& dr,
                                         const std::vector<float>& sf, int n_rows, float bf, float mb,
                                         std::vector<float>& u_right, std::vector<float>& depth) {
  const float min_d = 0.f, max_d = bf / mb;  // frame.cc:853-856
  std::vector<int> best_idx, best_dist;
  StereoRowBand(kl, dl, kr, dr, sf, n_rows, min_d, max_d, best_idx, best_dist);
  u_right.assign(kl.size(), -1.0f);
  depth.assign(kl.size(), -1.0f);
  std::vector<int> sad(kl.size(), -1);
  const int th_orb_dist = (TH_HIGH + TH_LOW) / 2;  // frame.cc:832
  check(m_, orbm_stereo_refine(m_, left, right, reinterpret_cast<const orbx_kp*>(kl.data()), (int)kl.size(),
                               reinterpret_cast<const orbx_kp*>(kr.data()), (int)kr.size(), best_idx.data(), best_dist.data(),
                               th_orb_dist, min_d, max_d, bf, u_right.data(), depth.data(), sad.data(), ORBX_MEM_HOST, nullptr));
}

std::vector<int> ORBmatcherGpu::ComputeDistinctiveDescriptors(const std::vector<std::vector<cv::Mat> >& obs) {
  std::vector<int32_t> offsets(obs.size() + 1, 0);
  int max_rows = 0;
  for (size_t p = 0; p < obs.size(); ++p) {
    offsets[p + 1] = offsets[p] + (int32_t)obs[p].size();
    if ((int)obs[p].size() > max_rows) max_rows = (int)obs[p].size();
  }
  std::vector<uint8_t> desc((size_t)offsets.back() * 32);
  for (size_t p = 0; p < obs.size(); ++p)
    for (size_t i = 0; i < obs[p].size(); ++i) std::memcpy(desc.data() + 32 * (size_t)(offsets[p] + i), obs[p][i].ptr(0), 32);
  std::vector<int> best((size_t)obs.size(), -1), median((size_t)obs.size(), 0);
  check(m_, orbm_distinctive(m_, desc.data(), offsets.data(), (int)obs.size(), max_rows, best.data(), median.data(),
                             ORBX_MEM_HOST, nullptr));
  return best;
}

int ORBmatcherGpu::SearchByProjection(const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc, const std::vector<float>& u_right,
                                      const std::vector<float>& sf, float min_x, float min_y, float inv_w, float inv_h, int cols,
                                      int rows, const std::vector<TrackedPoint>& points, const cv::Mat& point_desc,
                                      const std::vector<uint8_t>& already, float th, bool far_points, float th_far, float nnratio,
                                      std::vector<int>& assigned_point) {
  const int n = (int)keys.size();
  assigned_point.assign((size_t)n, -1);
  const bool bFactor = th != 1.0;  // orb_matcher.cc:48
  std::vector<int> origin;
  std::vector<orbm_window_query> windows;
  std::vector<float> q_ur, q_err;
  std::vector<uint8_t> qd;
  const std::vector<uint8_t> pd = dense_rows(point_desc);
  for (size_t p = 0; p < points.size(); ++p) {
    const TrackedPoint& mp = points[p];
    if (!mp.in_view) continue;                          // :52 (mbTrackInViewR belongs to the two-camera rig)
    if (far_points && mp.depth > th_far) continue;      // :54
    if (mp.bad) continue;                               // :56
    float r = mp.view_cos > 0.998 ? 2.5f : 4.0f;        // RadiusByViewingCos :208-213
    if (bFactor) r *= th;                               // :64
    const int lev = mp.level < 0 ? 0 : (mp.level >= (int)sf.size() ? (int)sf.size() - 1 : mp.level);
    const float radius = r * sf[lev];                   // :67-68
    const orbm_window_query w = {mp.proj_x, mp.proj_y, radius, mp.level - 1, mp.level};
    windows.push_back(w);
    q_ur.push_back(mp.proj_xr);
    q_err.push_back(radius);                            // :91
    qd.insert(qd.end(), pd.begin() + 32 * p, pd.begin() + 32 * (p + 1));
    origin.push_back((int)p);
  }
  const std::vector<uint8_t> d = dense_rows(desc);
  const orbm_grid_geom g = {min_x, min_y, inv_w, inv_h, cols, rows};
  std::vector<int32_t> assigned((size_t)std::max(n, 1), -1);
  int32_t nm = 0;
  const bool stereo = !u_right.empty();
  check(m_, orbm_search_by_projection(m_, reinterpret_cast<const orbx_kp*>(keys.data()), d.data(), n, &g, windows.data(), qd.data(),
                                      (int)windows.size(), already.empty() ? nullptr : already.data(),
                                      stereo ? u_right.data() : nullptr, stereo ? q_ur.data() : nullptr,
                                      stereo ? q_err.data() : nullptr, TH_HIGH, nnratio, assigned.data(), &nm, ORBX_MEM_HOST, nullptr));
  for (int i = 0; i < n; ++i)
    if (assigned[i] >= 0) assigned_point[i] = origin[assigned[i]];
  return nm;
}

int ORBmatcherGpu::SearchByProjectionLastFrame(const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc,
                                               const std::vector<float>& u_right, const std::vector<float>& sf, float bf, float min_x,
                                               float min_y, float inv_w, float inv_h, int cols, int rows,
                                               const std::vector<ProjectedPoint>& points, const cv::Mat& point_desc,
                                               const std::vector<uint8_t>& already, float th, bool forward, bool backward,
                                               bool check_orientation, std::vector<int>& assigned_point) {
  const int n = (int)keys.size(), nq = (int)points.size();
  assigned_point.assign((size_t)n, -1);
  std::vector<orbm_window_query> windows((size_t)nq);
  std::vector<float> q_ur((size_t)nq), q_err((size_t)nq), q_angle((size_t)nq);
  for (int p = 0; p < nq; ++p) {
    const ProjectedPoint& pt = points[p];
    const int o = pt.last_octave < 0 ? 0 : (pt.last_octave >= (int)sf.size() ? (int)sf.size() - 1 : pt.last_octave);
    const float radius = th * sf[o];                                            // :1566
    const orbm_window_query w = {pt.u, pt.v, radius, forward ? o : (backward ? 0 : o - 1),  // :1568-1576
                                 forward ? -1 : (backward ? o : o + 1)};
    windows[p] = w;
    q_ur[p] = pt.u - bf * pt.invzc;                                             // :1587
    q_err[p] = radius;                                                          // :1589
    q_angle[p] = pt.last_angle;
  }
  const std::vector<uint8_t> d = dense_rows(desc), qd = dense_rows(point_desc);
  const orbm_grid_geom g = {min_x, min_y, inv_w, inv_h, cols, rows};
  std::vector<int32_t> assigned((size_t)std::max(n, 1), -1);
  int32_t nm = 0;
  const bool stereo = !u_right.empty();
  check(m_, orbm_search_by_projection_last(m_, reinterpret_cast<const orbx_kp*>(keys.data()), d.data(), n, &g, windows.data(), qd.data(),
                                           q_angle.data(), nq, already.empty() ? nullptr : already.data(),
                                           stereo ? u_right.data() : nullptr, stereo ? q_ur.data() : nullptr,
                                           stereo ? q_err.data() : nullptr, TH_HIGH, check_orientation ? 1 : 0, assigned.data(), &nm,
                                           ORBX_MEM_HOST, nullptr));
  for (int i = 0; i < n; ++i) assigned_point[i] = assigned[i];
  return nm;
}

int ORBmatcherGpu::SearchByBoWImpl(bool keyframes, const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1,
                                   const std::vector<uint8_t>& has_point1,
                                   const std::map<unsigned int, std::vector<unsigned int> >& featvec1,
                                   const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2,
                                   const std::vector<uint8_t>* has_point2,
                                   const std::map<unsigned int, std::vector<unsigned int> >& featvec2, float nnratio,
                                   bool check_orientation, std::vector<int>& match_out) {
  typedef std::map<unsigned int, std::vector<unsigned int> > FeatVec;
  const int n1 = (int)keys1.size(), n2 = (int)keys2.size();
  const int cap = std::max(1, std::max(n1, n2));
  const int n_out = keyframes ? n1 : n2;  // vpMatches12 is indexed by side 1, vpMapPointMatches by the frame
  match_out.assign((size_t)n_out, -1);
  // a pool of two frames in the [frame][cap] layout of orbm_search_by_bow: 0 = side 1, 1 = side 2
  std::vector<orbx_kp> kps(2 * (size_t)cap);
  std::vector<uint8_t> desc(2 * (size_t)cap * 32, 0), has(2 * (size_t)cap, 0);
  std::vector<uint32_t> nodes(2 * (size_t)cap, 0), feats(2 * (size_t)cap, 0);
  std::vector<int32_t> begin(2 * (size_t)cap, 0);
  int32_t fv_n[2] = {0, 0}, fv_total[2] = {0, 0};
  const int32_t npf[2] = {n1, n2};
  if (n1) std::memcpy(kps.data(), keys1.data(), sizeof(orbx_kp) * (size_t)n1);
  if (n2) std::memcpy(kps.data() + cap, keys2.data(), sizeof(orbx_kp) * (size_t)n2);
  const std::vector<uint8_t> d1 = dense_rows(desc1), d2 = dense_rows(desc2);
  if (n1) std::memcpy(desc.data(), d1.data(), 32 * (size_t)n1);
  if (n2) std::memcpy(desc.data() + 32 * (size_t)cap, d2.data(), 32 * (size_t)n2);
  for (int i = 0; i < n1; ++i) has[i] = i < (int)has_point1.size() ? has_point1[i] : 0;
  for (int i = 0; i < n2; ++i) has[(size_t)cap + i] = has_point2 ? (i < (int)has_point2->size() ? (*has_point2)[i] : 0) : 1;
  const FeatVec* fv[2] = {&featvec1, &featvec2};
  for (int s = 0; s < 2; ++s) {
    const size_t o = (size_t)s * cap;
    for (FeatVec::const_iterator it = fv[s]->begin(); it != fv[s]->end(); ++it) {
      size_t cnt = 0;
      for (size_t j = 0; j < it->second.size(); ++j) cnt += it->second[j] < (unsigned)npf[s];
      if (fv_total[s] + (int)cnt > cap || fv_n[s] >= cap) throw std::runtime_error("ORBmatcherGpu::SearchByBoW: FeatureVector larger than the frame");
      nodes[o + fv_n[s]] = it->first;
      begin[o + fv_n[s]] = fv_total[s];
      ++fv_n[s];
      for (size_t j = 0; j < it->second.size(); ++j)
        if (it->second[j] < (unsigned)npf[s]) feats[o + fv_total[s]++] = it->second[j];
    }
  }
  const int32_t p1 = 0, p2 = 1;
  std::vector<int32_t> match((size_t)cap, -1);
  int32_t nm = 0;
  check(m_, (keyframes ? orbm_search_by_bow_kf : orbm_search_by_bow)(
                m_, kps.data(), desc.data(), cap, 2, npf, nodes.data(), begin.data(), fv_n, feats.data(), fv_total, has.data(), &p1, &p2,
                1, nnratio, check_orientation ? 1 : 0, match.data(), &nm, ORBX_MEM_HOST, nullptr));
  for (int i = 0; i < n_out; ++i) match_out[i] = match[i];
  return nm;
}

int ORBmatcherGpu::SearchByBoW(const std::vector<cv::KeyPoint>& keys_kf, const cv::Mat& desc_kf,
                               const std::vector<uint8_t>& has_point_kf,
                               const std::map<unsigned int, std::vector<unsigned int> >& featvec_kf,
                               const std::vector<cv::KeyPoint>& keys_f, const cv::Mat& desc_f,
                               const std::map<unsigned int, std::vector<unsigned int> >& featvec_f, float nnratio,
                               bool check_orientation, std::vector<int>& match_of_f) {
  return SearchByBoWImpl(false, keys_kf, desc_kf, has_point_kf, featvec_kf, keys_f, desc_f, nullptr, featvec_f, nnratio,
                         check_orientation, match_of_f);
}

int ORBmatcherGpu::SearchByBoW(const std::vector<cv::KeyPoint>& keys_un1, const cv::Mat& desc1,
                               const std::vector<uint8_t>& has_point1,
                               const std::map<unsigned int, std::vector<unsigned int> >& featvec1,
                               const std::vector<cv::KeyPoint>& keys_un2, const cv::Mat& desc2,
                               const std::vector<uint8_t>& has_point2,
                               const std::map<unsigned int, std::vector<unsigned int> >& featvec2, float nnratio,
                               bool check_orientation, std::vector<int>& match_of_1) {
  return SearchByBoWImpl(true, keys_un1, desc1, has_point1, featvec1, keys_un2, desc2, &has_point2, featvec2, nnratio,
                         check_orientation, match_of_1);
}

int ORBmatcherGpu::SearchForTriangulation(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1,
                                          const std::vector<uint8_t>& has_point1, const std::vector<float>& u_right1,
                                          const std::map<unsigned int, std::vector<unsigned int> >& featvec1,
                                          const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2,
                                          const std::vector<uint8_t>& has_point2, const std::vector<float>& u_right2,
                                          const std::map<unsigned int, std::vector<unsigned int> >& featvec2, const float f12[9],
                                          float epipole_x, float epipole_y, const std::vector<float>& sf,
                                          const std::vector<float>& sigma2, bool only_stereo, bool coarse, bool check_orientation,
                                          std::vector<std::pair<size_t, size_t> >& matched_pairs) {
  typedef std::map<unsigned int, std::vector<unsigned int> > FeatVec;
  const int n1 = (int)keys1.size(), n2 = (int)keys2.size();
  const int cap = std::max(1, std::max(n1, n2));
  matched_pairs.clear();
  std::vector<orbx_kp> kps(2 * (size_t)cap);
  std::vector<uint8_t> desc(2 * (size_t)cap * 32, 0), has(2 * (size_t)cap, 1);
  std::vector<float> ur(2 * (size_t)cap, -1.0f);
  std::vector<uint32_t> nodes(2 * (size_t)cap, 0), feats(2 * (size_t)cap, 0);
  std::vector<int32_t> begin(2 * (size_t)cap, 0);
  int32_t fv_n[2] = {0, 0}, fv_total[2] = {0, 0};
  const int32_t npf[2] = {n1, n2};
  if (n1) std::memcpy(kps.data(), keys1.data(), sizeof(orbx_kp) * (size_t)n1);
  if (n2) std::memcpy(kps.data() + cap, keys2.data(), sizeof(orbx_kp) * (size_t)n2);
  const std::vector<uint8_t> d1 = dense_rows(desc1), d2 = dense_rows(desc2);
  if (n1) std::memcpy(desc.data(), d1.data(), 32 * (size_t)n1);
  if (n2) std::memcpy(desc.data() + 32 * (size_t)cap, d2.data(), 32 * (size_t)n2);
  for (int i = 0; i < n1; ++i) { has[i] = i < (int)has_point1.size() ? has_point1[i] : 0; ur[i] = i < (int)u_right1.size() ? u_right1[i] : -1.0f; }
  for (int i = 0; i < n2; ++i) {
    has[(size_t)cap + i] = i < (int)has_point2.size() ? has_point2[i] : 0;
    ur[(size_t)cap + i] = i < (int)u_right2.size() ? u_right2[i] : -1.0f;
  }
  const FeatVec* fv[2] = {&featvec1, &featvec2};
  for (int s = 0; s < 2; ++s) {
    const size_t o = (size_t)s * cap;
    for (FeatVec::const_iterator it = fv[s]->begin(); it != fv[s]->end(); ++it) {
      size_t cnt = 0;
      for (size_t j = 0; j < it->second.size(); ++j) cnt += it->second[j] < (unsigned)npf[s];
      if (fv_total[s] + (int)cnt > cap || fv_n[s] >= cap) throw std::runtime_error("ORBmatcherGpu::SearchForTriangulation: FeatureVector larger than the key frame");
      nodes[o + fv_n[s]] = it->first;
      begin[o + fv_n[s]] = fv_total[s];
      ++fv_n[s];
      for (size_t j = 0; j < it->second.size(); ++j)
        if (it->second[j] < (unsigned)npf[s]) feats[o + fv_total[s]++] = it->second[j];
    }
  }
  const int32_t p1 = 0, p2 = 1;
  const float ep[2] = {epipole_x, epipole_y};
  std::vector<int32_t> match((size_t)cap, -1);
  int32_t nm = 0;
  check(m_, orbm_search_for_triangulation(m_, kps.data(), desc.data(), cap, 2, npf, nodes.data(), begin.data(), fv_n, feats.data(), fv_total,
                                          has.data(), ur.data(), &p1, &p2, 1, f12, ep, sf.data(), sigma2.data(), (int)sf.size(),
                                          only_stereo ? 1 : 0, coarse ? 1 : 0, check_orientation ? 1 : 0, match.data(), &nm,
                                          ORBX_MEM_HOST, nullptr));
  for (int i = 0; i < n1; ++i)  // :1030-1037
    if (match[i] >= 0) matched_pairs.push_back(std::make_pair((size_t)i, (size_t)match[i]));
  return nm;
}

void ORBmatcherGpu::WindowSearch(const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc, float min_x, float min_y,
                                 float inv_w, float inv_h, int cols, int rows, const std::vector<Window>& windows,
                                 const cv::Mat& window_desc, const std::vector<uint8_t>* already, std::vector<WindowBest>& out,
                                 const std::vector<float>* u_right, const std::vector<float>* window_u_right,
                                 const std::vector<float>* window_max_err) {
  const std::vector<uint8_t> d = dense_rows(desc), qd = dense_rows(window_desc);
  const orbm_grid_geom g = {min_x, min_y, inv_w, inv_h, cols, rows};
  out.resize(windows.size());
  check(m_, orbm_window_search_stereo(m_, reinterpret_cast<const orbx_kp*>(keys.data()), d.data(), (int)keys.size(), &g,
                                      reinterpret_cast<const orbm_window_query*>(windows.data()), qd.data(), (int)windows.size(),
                                      already ? already->data() : nullptr, u_right ? u_right->data() : nullptr,
                                      window_u_right ? window_u_right->data() : nullptr,
                                      window_max_err ? window_max_err->data() : nullptr,
                                      reinterpret_cast<orbm_window_result*>(out.data()), ORBX_MEM_HOST, nullptr));
}

}  // namespace ORB_SLAM_FUSION
