// oracle/ref_bow_shim.cc -- TEST INFRASTRUCTURE, not product code.
//
// C entry points over the reference's own vendored DBoW2 (3rdparty/DBoW2: TemplatedVocabulary.h,
// FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp, DUtils), compiled unmodified where
// it lies on the mini-cv shim (oracle/minicv; boost::serialization and cv::FileStorage are
// declaration-only stubs there).  Built by `make -C oracle ref` into _ref/libbow_ref.so; used to pin
// oracle/bow_oracle.c and to generate tests/golden/bow_*.npz.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;  // include/map/orb_vocabulary.h

extern "C" {

void* refbow_load_text(const char* path) {
  ORBVocabulary* v = new ORBVocabulary();
  if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
  return v;
}
void refbow_destroy(void* h) { delete (ORBVocabulary*)h; }
int refbow_words(void* h) { return (int)((ORBVocabulary*)h)->size(); }

// Frame::ComputeBoW (src/map/frame.cc:761-766): transform(vCurrentDesc, mBowVec, mFeatVec, levelsup)
void refbow_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* bow_n,
                      uint32_t* fv_nodes, int32_t* fv_begin, int* fv_n, uint32_t* fv_feats, int* fv_total) {
  ORBVocabulary* v = (ORBVocabulary*)h;
  std::vector<cv::Mat> feats((size_t)n);
  for (int i = 0; i < n; i++) {  // Converter::toDescriptorVector: one 1 x 32 row per feature
    feats[i] = cv::Mat(1, 32, CV_8U);
    std::memcpy(feats[i].data, desc + 32 * (size_t)i, 32);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  v->transform(feats, bv, fv, levelsup);
  int nb = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++nb) {
    bow_ids[nb] = it->first;
    bow_vals[nb] = it->second;
  }
  *bow_n = nb;
  int nf = 0, pos = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++nf) {
    fv_nodes[nf] = it->first;
    fv_begin[nf] = pos;
    for (size_t j = 0; j < it->second.size(); j++) fv_feats[pos++] = it->second[j];
  }
  *fv_n = nf;
  *fv_total = pos;
}

// transform(feature) -> word id (:989-998)
void refbow_words_of(void* h, const uint8_t* desc, int n, uint32_t* word_id) {
  ORBVocabulary* v = (ORBVocabulary*)h;
  for (int i = 0; i < n; i++) {
    cv::Mat f(1, 32, CV_8U);
    std::memcpy(f.data, desc + 32 * (size_t)i, 32);
    word_id[i] = v->transform(f);
  }
}

}  // extern "C"
