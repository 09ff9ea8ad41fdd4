// ORBVocabularyGpu: the reference's ORBVocabulary calls on the data path, forwarded to the C ABI (orbv_*).
#include "cam/orb_feature/orb_vocabulary_gpu.h"

#include <cstring>
#include <stdexcept>

#include "orbx.h"

namespace ORB_SLAM_FUSION {

ORBVocabularyGpu::ORBVocabularyGpu(int device) : device_(device), v_(nullptr) {}

ORBVocabularyGpu::~ORBVocabularyGpu() { orbv_destroy(v_); }

bool ORBVocabularyGpu::loadFromTextFile(const std::string& filename) {
  orbv_destroy(v_);
  v_ = nullptr;
  return orbv_load_text(device_, filename.c_str(), &v_) == ORBX_OK;
}

unsigned int ORBVocabularyGpu::size() const {
  int32_t words = 0;
  if (v_) orbv_info(v_, nullptr, nullptr, nullptr, nullptr, nullptr, &words);
  return (unsigned int)words;
}

bool ORBVocabularyGpu::empty() const { return size() == 0; }

void ORBVocabularyGpu::transform(const cv::Mat& descriptors, std::map<unsigned int, double>& v,
                                 std::map<unsigned int, std::vector<unsigned int> >& fv, int levelsup) const {
  v.clear();
  fv.clear();
  const int n = descriptors.rows;
  if (!v_ || n == 0) return;
  std::vector<uint8_t> rows((size_t)n * 32);
  for (int i = 0; i < n; i++) std::memcpy(&rows[32 * (size_t)i], descriptors.ptr(i), 32);
  std::vector<uint32_t> ids(n), nodes(n), feats(n);
  std::vector<double> vals(n);
  std::vector<int32_t> begin(n);
  int32_t nb = 0, nf = 0, tot = 0;
  if (orbv_transform(v_, rows.data(), n, nullptr, 1, levelsup, ids.data(), vals.data(), &nb, nodes.data(), begin.data(), &nf,
                     feats.data(), &tot, ORBX_MEM_HOST, nullptr) != ORBX_OK)
    throw std::runtime_error(orbv_last_error(v_));
  for (int j = 0; j < nb; j++) v.insert(v.end(), std::make_pair(ids[j], vals[j]));  // already in increasing word id
  for (int j = 0; j < nf; j++) {
    const int e = j + 1 < nf ? begin[j + 1] : tot;
    fv.insert(fv.end(), std::make_pair(nodes[j], std::vector<unsigned int>(feats.begin() + begin[j], feats.begin() + e)));
  }
}

void ORBVocabularyGpu::transform(const std::vector<cv::Mat>& features, std::map<unsigned int, double>& v,
                                 std::map<unsigned int, std::vector<unsigned int> >& fv, int levelsup) const {
  cv::Mat all((int)features.size(), 32, CV_8U);
  for (size_t i = 0; i < features.size(); i++) std::memcpy(all.ptr((int)i), features[i].ptr(0), 32);
  transform(all, v, fv, levelsup);
}

unsigned int ORBVocabularyGpu::transform(const cv::Mat& feature) const {
  if (empty()) return 0;  // :991-993
  uint32_t word = 0, node = 0;
  double w = 0;
  if (orbv_features(v_, feature.ptr(0), 1, 0, &word, &w, &node, ORBX_MEM_HOST, nullptr) != ORBX_OK)
    throw std::runtime_error(orbv_last_error(v_));
  return word;
}

}  // namespace ORB_SLAM_FUSION
