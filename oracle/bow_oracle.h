/* oracle/bow_oracle.h -- TEST INFRASTRUCTURE (CPU restatement), not product code.
 *
 * Bag-of-words transform of a frame's descriptors, restating the reference's vendored DBoW2:
 *   TemplatedVocabulary::loadFromTextFile   3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1246-1330
 *   TemplatedVocabulary::transform (1 feat) 3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1179
 *   TemplatedVocabulary::transform (frame)  3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1056-1118
 *   BowVector::addWeight / addIfNotExist / normalize   3rdparty/DBoW2/DBoW2/BowVector.cpp:30-70
 *   FeatureVector::addFeature               3rdparty/DBoW2/DBoW2/FeatureVector.cpp:28-38
 *   FORB::distance                          3rdparty/DBoW2/DBoW2/FORB.cpp:71-88
 * as called by Frame::ComputeBoW (src/map/frame.cc:761-766, levelsup = 4).
 * Pinned by tests/test_oracle_bow.py against the reference's own DBoW2 sources compiled on the
 * mini-cv shim (oracle/_ref/libbow_ref.so) and against tests/golden/bow_*.npz made from it.
 */
#ifndef BOW_ORACLE_H
#define BOW_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_vocab orc_vocab;

/* Node arrays indexed by node id (0 = root; parent[0], is_leaf[0], desc row 0, weight[0] unused).
 * Children are visited in increasing node id (loadFromTextFile pushes them in file order);
 * word ids are handed out to leaves in increasing node id (:1319-1324).
 * scoring: 0 L1_NORM .. 5 DOT_PRODUCT; weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY (BowVector.h:31-44). */
orc_vocab* orc_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                            const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
/* the text format of loadFromTextFile; NULL if the file cannot be read or the header is out of range */
orc_vocab* orc_vocab_load_text(const char* path);
void orc_vocab_destroy(orc_vocab* v);
int orc_vocab_nodes(const orc_vocab* v);
int orc_vocab_words(const orc_vocab* v);
/* copies the node arrays out (each may be NULL) */
void orc_vocab_arrays(const orc_vocab* v, int* k, int* L, int* scoring, int* weighting, int32_t* parent, uint8_t* is_leaf,
                      uint8_t* desc, double* weight);

/* transform(feature, word_id, weight, nid, levelsup) for n features (:1139-1179).  node_id is the node on
 * the path at level L - levelsup (0 = root when that level is <= 0); when the path ends in a leaf above
 * that level the reference leaves *nid unset -- the restatement reports that leaf. */
void orc_bow_features(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                      uint32_t* node_id);

/* transform(features, BowVector, FeatureVector, levelsup) for one frame (:1056-1118).
 * bow_ids/bow_vals: the BowVector in increasing word id (capacity n), *bow_n entries.
 * fv_nodes/fv_begin: the FeatureVector's node ids in increasing order and the start of each node's
 * feature list inside fv_feats (capacity n each), *fv_n nodes; fv_feats: feature indices, grouped by node,
 * increasing inside a group; *fv_total = number of features that were not stopped (weight > 0). */
void orc_bow_transform(const orc_vocab* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals,
                       int* bow_n, uint32_t* fv_nodes, int32_t* fv_begin, int* fv_n, uint32_t* fv_feats, int* fv_total);

/* Deterministic synthetic vocabulary (SURVEY.md 8(d) style, splitmix64): a complete k-ary tree of depth L in
 * breadth-first node order; a child's descriptor is its parent's with ~256 >> level random bits flipped
 * (the root's children are random rows); leaf weights are (splitmix % 997 + 1) / 64.0, every 29th word is
 * "stopped" (weight 0).  Arrays as orc_vocab_create takes them; returns the node count
 * ((k^(L+1) - 1) / (k - 1)), also when the output pointers are NULL. */
int orc_synth_vocab(int k, int L, uint64_t seed, int32_t* parent, uint8_t* is_leaf, uint8_t* desc, double* weight);
/* writes the arrays in loadFromTextFile's format WITHOUT a final newline (a trailing newline makes
 * the reference's loader append a phantom child of the root with an unset descriptor, :1288-1296) */
int orc_vocab_save_text(const char* path, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                        const uint8_t* is_leaf, const uint8_t* desc, const double* weight);

#ifdef __cplusplus
}
#endif
#endif
