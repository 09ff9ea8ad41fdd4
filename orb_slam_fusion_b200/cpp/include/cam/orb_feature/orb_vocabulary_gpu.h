// GPU counterpart of the reference's ORBVocabulary (include/cam/orb_feature/orb_vocabulary.h =
// DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB>) for the calls on the data path, over
// the C ABI of include/orbx.h (orbv_*):
//   loadFromTextFile   3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1246-1330   (System start-up)
//   transform          :1056-1118 as called by Frame::ComputeBoW (src/map/frame.cc:761-766) and KeyFrame::ComputeBoW
// DBoW2::BowVector is a std::map<WordId, WordValue> and DBoW2::FeatureVector a
// std::map<NodeId, std::vector<unsigned int>> (BowVector.h:47, FeatureVector.h:24), so the reference's
// mBowVec / mFeatVec members bind to the std::map references below unchanged.
#ifndef ORB_VOCABULARY_GPU_H
#define ORB_VOCABULARY_GPU_H

#include <map>
#include <opencv2/opencv.hpp>
#include <string>
#include <vector>

struct orbv_vocab;  // include/orbx.h

namespace ORB_SLAM_FUSION {

class ORBVocabularyGpu {
 public:
  explicit ORBVocabularyGpu(int device = 0);
  ~ORBVocabularyGpu();
  ORBVocabularyGpu(const ORBVocabularyGpu&) = delete;
  ORBVocabularyGpu& operator=(const ORBVocabularyGpu&) = delete;

  bool loadFromTextFile(const std::string& filename);  // false on failure, like the reference
  unsigned int size() const;                           // number of words
  bool empty() const;

  // transform(features, v, fv, levelsup): features = Converter::toDescriptorVector(mDescriptors), 1 x 32 CV_8U rows
  void transform(const std::vector<cv::Mat>& features, std::map<unsigned int, double>& v,
                 std::map<unsigned int, std::vector<unsigned int> >& fv, int levelsup) const;
  // the same on the N x 32 descriptor matrix itself (no per-row Mats)
  void transform(const cv::Mat& descriptors, std::map<unsigned int, double>& v,
                 std::map<unsigned int, std::vector<unsigned int> >& fv, int levelsup) const;
  // transform(feature) -> word id (:989-998)
  unsigned int transform(const cv::Mat& feature) const;

  orbv_vocab* handle() const { return v_; }

 private:
  int device_;
  orbv_vocab* v_;
};

}  // namespace ORB_SLAM_FUSION

#endif
