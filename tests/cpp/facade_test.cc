// tests/cpp/facade_test.cc -- exercises the C++ drop-in classes (orb_slam_fusion_b200/cpp) the way
// the reference's Frame does (frame.cc:467-476, 834, 1154), compiled against the mini-cv stand-in
// for OpenCV (oracle/minicv, test infrastructure).  Writes its outputs to a binary file that
// tests/test_cpp_facade.py compares with the oracle.
//   facade_test <in.raw> <w> <h> <num_feats> <lap0> <lap1> <out.bin> [<vocabulary.txt>]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cam/orb_feature/orb_extractor.h"
#include "cam/orb_feature/orb_matcher_gpu.h"
#include "cam/orb_feature/orb_vocabulary_gpu.h"

using namespace ORB_SLAM_FUSION;

static void put(FILE* f, const void* p, size_t n) { fwrite(p, 1, n, f); }

int main(int argc, char** argv) {
  if (argc != 8 && argc != 9) return 2;
  const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]);
  std::vector<int> lap = {atoi(argv[5]), atoi(argv[6])};
  std::vector<uint8_t> buf((size_t)w * h);
  FILE* fi = fopen(argv[1], "rb");
  if (!fi || fread(buf.data(), 1, buf.size(), fi) != buf.size()) return 3;
  fclose(fi);
  cv::Mat img(h, w, CV_8UC1, buf.data());

  OrbExtractor* extractor = new OrbExtractor(nf, 1.2f, 8, 20, 7);  // tracking.cc:195
  std::vector<cv::KeyPoint> keys;
  cv::Mat desc;
  const int mono = (*extractor)(img, cv::Mat(), keys, desc, lap);  // frame.cc:472-475
  std::vector<cv::KeyPoint> none;
  cv::Mat nodesc(3, 32, CV_8U);
  const int empty_rc = (*extractor)(cv::Mat(), cv::Mat(), none, nodesc, lap);  // :1016

  FILE* fo = fopen(argv[7], "wb");
  const int32_t hdr[6] = {mono, (int32_t)keys.size(), desc.rows, desc.cols, empty_rc, extractor->GetLevels()};
  put(fo, hdr, sizeof(hdr));
  put(fo, keys.data(), keys.size() * sizeof(cv::KeyPoint));
  for (int i = 0; i < desc.rows; i++) put(fo, desc.ptr(i), 32);
  const std::vector<float> sf = extractor->GetScaleFactors();
  put(fo, sf.data(), sf.size() * sizeof(float));
  // img_pyramid_ as Frame::ComputeStereoMatches reads it (frame.cc:834, 913-931): rows/cols and pixels
  for (int lev = 0; lev < extractor->GetLevels(); lev++) {
    const cv::Mat& m = extractor->img_pyramid_[lev];
    const int32_t wh[2] = {m.cols, m.rows};
    put(fo, wh, sizeof(wh));
    for (int y = -19; y < m.rows + 19; y++) put(fo, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - 19, (size_t)m.cols + 38);
  }
  // brute-force 2-NN of the first 64 descriptors against all of them (frame.cc:1154 pattern)
  ORBmatcherGpu matcher(0);
  std::vector<std::vector<cv::DMatch> > matches;
  const int nq = desc.rows < 64 ? desc.rows : 64;
  matcher.KnnMatch2(desc.rowRange(0, nq), desc, matches);
  for (int i = 0; i < nq; i++) {
    int32_t rec[4] = {-1, -1, -1, -1};
    for (size_t k = 0; k < matches[i].size(); k++) { rec[2 * k] = matches[i][k].trainIdx; rec[2 * k + 1] = (int32_t)matches[i][k].distance; }
    put(fo, rec, sizeof(rec));
  }
  // stereo row band of the frame against itself: every keypoint must find itself at distance 0
  std::vector<int> bi, bd;
  matcher.StereoRowBand(keys, desc, keys, desc, sf, h, 0.f, 40.f, bi, bd);
  put(fo, bi.data(), bi.size() * sizeof(int));
  put(fo, bd.data(), bd.size() * sizeof(int));
  {
    // ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th = 3) (orb_matcher.cc:42-134): every keypoint projected as a
    // map point one pixel off, every fifth one not in view, every seventh keypoint already matched
    std::vector<ORBmatcherGpu::TrackedPoint> pts(keys.size());
    std::vector<uint8_t> already(keys.size(), 0);
    for (size_t i = 0; i < keys.size(); i++) {
      const ORBmatcherGpu::TrackedPoint t = {keys[i].pt.x + 1.0f, keys[i].pt.y - 1.0f, 0.f, i % 2 ? 1.0f : 0.9f, 10.f, keys[i].octave,
                                             i % 5 != 0, false};
      pts[i] = t;
      already[i] = i % 7 == 0;
    }
    std::vector<int> assigned;
    const int32_t nproj = matcher.SearchByProjection(keys, desc, std::vector<float>(), sf, 0.f, 0.f, 64.f / w, 48.f / h, 64, 48, pts, desc,
                                                     already, 3.0f, false, 50.f, 0.8f, assigned);
    put(fo, &nproj, 4);
    put(fo, assigned.data(), assigned.size() * sizeof(int));
    // ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th = 7, bMono) (orb_matcher.cc:1518-1728): the frame against itself
    // moved by (2, -1) px, neutral motion case, orientation check on
    std::vector<ORBmatcherGpu::ProjectedPoint> lp(keys.size());
    for (size_t i = 0; i < keys.size(); i++) {
      const ORBmatcherGpu::ProjectedPoint t = {keys[i].pt.x + 2.0f, keys[i].pt.y - 1.0f, 0.1f, keys[i].octave, keys[i].angle};
      lp[i] = t;
    }
    std::vector<int> assigned_last;
    const int32_t nlast = matcher.SearchByProjectionLastFrame(keys, desc, std::vector<float>(), sf, 47.9f, 0.f, 0.f, 64.f / w, 48.f / h, 64, 48,
                                                              lp, desc, already, 7.0f, false, false, true, assigned_last);
    put(fo, &nlast, 4);
    put(fo, assigned_last.data(), assigned_last.size() * sizeof(int));
  }
  if (argc == 9) {
    // Frame::ComputeBoW (frame.cc:761-766): toDescriptorVector + transform(vCurrentDesc, mBowVec, mFeatVec, 4)
    ORBVocabularyGpu voc(0);
    if (!voc.loadFromTextFile(argv[8])) return 4;
    if (ORBVocabularyGpu(0).loadFromTextFile("/nonexistent/voc.txt")) return 5;
    std::vector<cv::Mat> vdesc;
    for (int i = 0; i < desc.rows; i++) vdesc.push_back(desc.row(i));
    std::map<unsigned int, double> bow;
    std::map<unsigned int, std::vector<unsigned int> > fvec;
    voc.transform(vdesc, bow, fvec, 4);
    const int32_t cnt[3] = {(int32_t)voc.size(), (int32_t)bow.size(), (int32_t)fvec.size()};
    put(fo, cnt, sizeof(cnt));
    for (std::map<unsigned int, double>::const_iterator it = bow.begin(); it != bow.end(); ++it) {
      put(fo, &it->first, 4);
      put(fo, &it->second, 8);
    }
    for (std::map<unsigned int, std::vector<unsigned int> >::const_iterator it = fvec.begin(); it != fvec.end(); ++it) {
      const uint32_t hd[2] = {it->first, (uint32_t)it->second.size()};
      put(fo, hd, sizeof(hd));
      put(fo, it->second.data(), it->second.size() * 4);
    }
    const uint32_t w0 = desc.rows ? voc.transform(desc.row(0)) : 0;
    put(fo, &w0, 4);
    // ORBmatcher::SearchByBoW (orb_matcher.cc:215-389) of the frame against itself: every feature that holds a map
    // point (every third does not) must claim itself
    std::vector<uint8_t> has_point(keys.size(), 1);
    for (size_t i = 0; i < has_point.size(); i += 3) has_point[i] = 0;
    std::vector<int> match_of_f;
    const int32_t nmatch = matcher.SearchByBoW(keys, desc, has_point, fvec, keys, desc, fvec, 0.7f, true, match_of_f);
    put(fo, &nmatch, 4);
    put(fo, match_of_f.data(), match_of_f.size() * sizeof(int));
    // the key-frame / key-frame form (orb_matcher.cc:697-815)
    std::vector<int> match_of_1;
    const int32_t nmatch_kf = matcher.SearchByBoW(keys, desc, has_point, fvec, keys, desc, has_point, fvec, 0.8f, true, match_of_1);
    put(fo, &nmatch_kf, 4);
    put(fo, match_of_1.data(), match_of_1.size() * sizeof(int));
    // ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040) of the frame against itself: horizontal epipolar lines,
    // epipole far away, every third feature already holds a map point
    std::vector<uint8_t> with_point(keys.size(), 0);
    for (size_t i = 0; i < with_point.size(); i += 3) with_point[i] = 1;
    std::vector<float> no_right(keys.size(), -1.0f);
    const float f12[9] = {0.f, 0.f, 0.f, 0.f, 0.f, -2.4e-4f, 0.f, 2.4e-4f, 0.f};
    std::vector<std::pair<size_t, size_t> > tri_pairs;
    const int32_t ntri = matcher.SearchForTriangulation(keys, desc, with_point, no_right, fvec, keys, desc, with_point, no_right, fvec, f12,
                                                        -5000.f, 200.f, sf, extractor->GetScaleSigmaSquares(), false, false, true, tri_pairs);
    put(fo, &ntri, 4);
    std::vector<int> tri(keys.size(), -1);
    for (size_t j = 0; j < tri_pairs.size(); j++) tri[tri_pairs[j].first] = (int)tri_pairs[j].second;
    put(fo, tri.data(), tri.size() * sizeof(int));
  }
  fclose(fo);
  delete extractor;
  return 0;
}
