/*
 * oracle/orb_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's ORB front-end hot path:
 *   OrbExtractor (src/cam/orb_feature/orb_extractor.cc:407-465, 744-849, 1011-1117)
 *   ORBmatcher::DescriptorDistance (src/cam/orb_feature/orb_matcher.cc:1877-1891)
 *   and the data-parallel matcher call patterns (frame.cc:836-900, 1154-1162,
 *   orb_matcher.cc:66-113).
 * Pinned against cv2 4.13.0 (tests/test_oracle_cv2.py), against the committed
 * golden vectors (tests/golden/) and against the reference's own orb_extractor.cc
 * compiled on the mini-cv shim (oracle/_ref, tests/test_oracle_vs_ref.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may link or call this.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

typedef struct {
  int num_feats;
  float scale_factor;
  int num_levs;
  int ini_th_fast;
  int min_th_fast;
} orc_params;

/* Same field order and size (28 B) as cv::KeyPoint. */
typedef struct {
  float x, y, size, angle, response;
  int octave, class_id;
} orc_kp;

typedef struct orc_extractor orc_extractor;

enum { ORC_TRIG_LIBM = 0, ORC_TRIG_CR = 1 };

orc_extractor* orc_create(const orc_params* p);
void orc_destroy(orc_extractor* e);
/* trig mode for the rBRIEF steering: ORC_TRIG_LIBM = cosf/sinf as the reference
 * (orb_extractor.cc:106); ORC_TRIG_CR = (float)cos((double)a) (correctly rounded). */
void orc_set_trig(orc_extractor* e, int mode);

/* A.1 tables; arrays must hold num_levs entries (umax: 16). */
void orc_tables(const orc_extractor* e, float* scale, float* inv_scale, float* sigma2,
                float* inv_sigma2, int* quota, int* umax);

/* orb_extractor.cc:1093-1117.  Levels are kept with the 19-px REFLECT_101 border. */
int orc_compute_pyramid(orc_extractor* e, const uint8_t* img, int w, int h, size_t stride);
/* pointer to pixel (0,0) of level lev (inside its bordered buffer) */
const uint8_t* orc_level(const orc_extractor* e, int lev, int* w, int* h, size_t* stride);
/* blurred copy of level lev as used for descriptors (valid after orc_extract for
 * levels that had keypoints; recomputed on demand otherwise) */
const uint8_t* orc_blurred_level(orc_extractor* e, int lev, size_t* stride);

/* orb_extractor.cc:1011-1091.  Returns 0 on success, -1 for an empty image,
 * -2 if cap is too small.  *n_mono is the reference's return value. */
int orc_extract(orc_extractor* e, const uint8_t* img, int w, int h, size_t stride,
                int lap0, int lap1, orc_kp* kps, uint8_t* desc, int cap, int* n, int* n_mono);

/* Stage intermediates of the last orc_extract call (per level).
 * candidates: FAST survivors in to_dist_kps order, coords relative to (16,16)
 *             (orb_extractor.cc:816-823) as (x, y, response) int triples.
 * selected:   DistributeOctTree output order, level coords incl. the +16 shift,
 *             with angle (orb_extractor.cc:834-848). */
int orc_candidates(const orc_extractor* e, int lev, int* xyr, int cap);
int orc_selected(const orc_extractor* e, int lev, orc_kp* out, int cap);

/* Stand-alone stages for stage-wise parity */
int orc_fast_grid(const uint8_t* lvl, int w, int h, size_t stride, int ini_th, int min_th,
                  int* xyr, int cap);
int orc_octree(const int* xyr, int n, int min_x, int max_x, int min_y, int max_y, int quota,
               int* out_index, int cap);
float orc_ic_angle(const uint8_t* lvl, size_t stride, int cx, int cy);
void orc_rbrief(const uint8_t* blurred, size_t stride, int cx, int cy, float angle_deg,
                int trig_mode, uint8_t desc[32]);

/* ---- matching ---- */
int orc_hamming(const uint8_t* a, const uint8_t* b); /* orb_matcher.cc:1877-1891 */

/* cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) as used at frame.cc:1154: per query the two
 * nearest train rows ordered by (distance, index).  idx/dist are [nq][2]; missing
 * entries (nd < 2) are idx=-1, dist=INT32_MAX.  nthreads>1 splits queries over threads. */
void orc_knn2(const uint8_t* q, int nq, const uint8_t* d, int64_t nd, int64_t* idx, int* dist,
              int nthreads);
/* frame.cc:1162 ratio test: accept iff two neighbours and d0 < d1*ratio */
int orc_ratio_accept(int d0, int d1, int have2, double ratio);

/* frame.cc:836-900: best right keypoint per left keypoint in the row band.
 * best_dist[i]=TH_HIGH(100) and best_idx[i]=-1 when nothing beats the threshold
 * (the reference leaves bestIdxR=0 in that case but never reads it). */
void orc_stereo_rowband(const orc_kp* kl, const uint8_t* dl, int nl, const orc_kp* kr,
                        const uint8_t* dr, int nr, const float* scale_factors, int n_rows,
                        float min_d, float max_d, int* best_idx, int* best_dist);

/* frame.cc:903-985, the rest of Frame::ComputeStereoMatches: for every left
 * keypoint whose row-band match has best_dist < th_orb_dist, an 11x11 SAD search over 11 horizontal
 * shifts on the keypoint's pyramid level, parabola sub-pixel fit, disparity gate, then the
 * median-distance outlier cut over all accepted matches.  Levels are given WITH their 19-px
 * REFLECT_101 border (pointer to pixel (0,0), like img_pyramid_).  Outputs: u_right[i], depth[i]
 * (-1 where there is no stereo match), sad[i] (the SAD of accepted matches, -1 otherwise).
 * Pinned: frame.cc as a file cannot be compiled here (Eigen/Sophus/DBoW2), but its lines 828-986 are spliced
 * into oracle/ref_frame_shim.cc and compiled; tests/test_oracle_vs_ref_frame.py compares bit for bit. */
typedef struct {
  const uint8_t* px;
  int w, h;
  size_t stride;
} orc_level_view;
void orc_stereo_refine(const orc_level_view* left, const orc_level_view* right, int n_levels,
                       const orc_kp* kl, int nl, const orc_kp* kr, const int* best_idx,
                       const int* best_dist, const float* scale_factors, const float* inv_scale_factors,
                       int th_orb_dist, float min_d, float max_d, float bf, float* u_right, float* depth,
                       int* sad);

/* MapPoint::ComputeDistinctiveDescriptors (mappoint.cc:365-428) for a batch of
 * map points: point p owns descriptor rows [offsets[p], offsets[p+1]); all-pairs Hamming distances,
 * per row the median (sorted row incl. the 0 self-distance, element int(0.5*(N-1))), the first row
 * with the least median wins.  best_idx[p] is relative to offsets[p] (-1 for a point without rows),
 * best_median[p] its median.  Pinned on mappoint.cc:365-433 spliced into oracle/ref_frame_shim.cc
 * (tests/test_oracle_vs_ref_frame.py). */
void orc_distinctive(const uint8_t* desc, const int* offsets, int n_points, int* best_idx, int* best_median);

/* Frame grid (frame.cc:438-465 AssignFeaturesToGrid + :679-746 GetFeaturesInArea) and the
 * best / second-best inner loop of SearchByProjection (orb_matcher.cc:66-113).  Pinned on frame.cc:438-465,
 * :679-759 and orb_matcher.cc:42-213 spliced into oracle/ref_frame_shim.cc (tests/test_oracle_vs_ref_frame.py
 * runs the reference's whole SearchByProjection against this search + the greedy claim restated in the test). */
typedef struct {
  float min_x, min_y, inv_w, inv_h; /* mnMinX, mnMinY, mfGridElementWidthInv/HeightInv */
  int cols, rows;                   /* FRAME_GRID_COLS=64, FRAME_GRID_ROWS=48 */
} orc_grid_geom;

typedef struct {
  float u, v, r;
  int min_level, max_level;
} orc_window_query;

typedef struct {
  int best_dist, best_idx, best_level, best_dist2, best_level2;
} orc_window_result;

void orc_window_search(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                       const orc_window_query* q, const uint8_t* qdesc, int nq,
                       const uint8_t* skip /* n bytes or NULL */, orc_window_result* out);
/* the same with the stereo gate of orb_matcher.cc:89-92 / :1586-1590: a keypoint with a right coordinate
 * (kp_u_right[i] > 0) is skipped when |q_u_right[q] - kp_u_right[i]| > q_max_err[q] */
void orc_window_search_stereo(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                              const orc_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                              const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                              orc_window_result* out);

/* The search of ORBmatcher::Fuse (orb_matcher.cc:1130-1187): the window search with the chi-square reprojection gate of
 * :1159-1178 (7.8 with a right coordinate >= 0, else 5.99) instead of the tracker's stereo gate; no skip flags.
 * Pinned on orb_matcher.cc:1144-1187 spliced into oracle/ref_frame_shim.cc. */
void orc_window_search_fuse(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g, const orc_window_query* q,
                            const uint8_t* qdesc, int nq, const float* kp_u_right, const float* q_u_right,
                            const float* inv_level_sigma2, orc_window_result* out);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) (orb_matcher.cc:42-134, Nleft == -1) as a whole:
 * the window search per map point in order + the greedy claim.  assigned[i] = query stored in F.mvpMapPoints[i] or -1;
 * returns nmatches.  Pinned on orb_matcher.cc:42-213 spliced into oracle/ref_frame_shim.cc. */
int orc_search_by_projection(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                             const orc_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                             const float* kp_u_right, const float* q_u_right, const float* q_max_err, int th_high,
                             float nnratio, int* assigned);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (orb_matcher.cc:1518-1728,
 * Nleft == -1) after the projection: nearest unclaimed keypoint per window within th_high, greedy claim, rotation
 * histogram.  Pinned on orb_matcher.cc:1518-1728 spliced into oracle/ref_frame_shim.cc. */
int orc_search_by_projection_last(const orc_kp* kps, const uint8_t* desc, int n, const orc_grid_geom* g,
                                  const orc_window_query* q, const uint8_t* qdesc, const float* q_angle, int nq,
                                  const uint8_t* skip, const float* kp_u_right, const float* q_u_right,
                                  const float* q_max_err, int th_high, int check_orientation, int* assigned);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (orb_matcher.cc:215-386; Nleft == -1): per shared
 * vocabulary node, every key-frame feature with a map point takes the best unclaimed frame feature of the node
 * (<= TH_LOW, ratio test), in order; then the 30-bin rotation histogram keeps the three dominant bins
 * (ComputeThreeMaxima :1841-1873).  FeatureVectors as sorted node ids / group starts / feature indices.
 * match_of_f[i] = key-frame feature matched to frame feature i or -1; returns nmatches.
 * Pinned on orb_matcher.cc:215-386 + :1841-1873 spliced into oracle/ref_frame_shim.cc. */
int orc_search_by_bow(const orc_kp* kps_kf, const uint8_t* desc_kf, const uint8_t* has_point_kf,
                      const uint32_t* nodes_kf, const int* begin_kf, int n_nodes_kf, const uint32_t* feats_kf, int total_kf,
                      const orc_kp* kps_f, const uint8_t* desc_f, int n_f,
                      const uint32_t* nodes_f, const int* begin_f, int n_nodes_f, const uint32_t* feats_f, int total_f,
                      float nnratio, int check_orientation, int* match_of_f);

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (orb_matcher.cc:697-815; loop closing): both sides need
 * a good map point, strict < TH_LOW, each side-2 feature claimed once; match_of_1[i] = side-2 feature or -1.
 * Pinned on orb_matcher.cc:697-815 spliced into oracle/ref_frame_shim.cc. */
int orc_search_by_bow_kf(const orc_kp* kps1, const uint8_t* desc1, const uint8_t* has_point1, int n1,
                         const uint32_t* nodes1, const int* begin1, int n_nodes1, const uint32_t* feats1, int total1,
                         const orc_kp* kps2, const uint8_t* desc2, const uint8_t* has_point2, int n2,
                         const uint32_t* nodes2, const int* begin2, int n_nodes2, const uint32_t* feats2, int total2,
                         float nnratio, int check_orientation, int* match_of_1);

/* ORBmatcher::SearchForTriangulation (orb_matcher.cc:817-1040; LocalMapping::CreateNewMapPoints) for key frames with one
 * pinhole camera: bag-of-words guided nearest match between features WITHOUT map points, epipole and epipolar gates
 * (F12 row-major, ep = epipole in image 2), greedy claim, rotation histogram.  match_of_1[i] = feature of key frame 2 or
 * -1; returns nmatches.  Pinned on orb_matcher.cc:817-1040 + pinhole_model.cc:121-134 spliced into oracle/ref_frame_shim.cc. */
int orc_search_for_triangulation(const orc_kp* kps1, const uint8_t* desc1, const uint8_t* has_point1, const float* u_right1, int n1,
                                 const uint32_t* nodes1, const int* begin1, int n_nodes1, const uint32_t* feats1, int total1,
                                 const orc_kp* kps2, const uint8_t* desc2, const uint8_t* has_point2, const float* u_right2, int n2,
                                 const uint32_t* nodes2, const int* begin2, int n_nodes2, const uint32_t* feats2, int total2,
                                 const float* f12, const float* ep, const float* scale_factors, const float* level_sigma2,
                                 int only_stereo, int coarse, int check_orientation, int* match_of_1);

/* ---- deterministic synthetic inputs ---- */
uint64_t orc_splitmix64(uint64_t x);
void orc_synth_blocks_v1(uint8_t* img, int w, int h, size_t stride, uint64_t seed, uint64_t frame,
                         int shift_x, uint64_t noise_seed);
void orc_synth_uniform_v1(uint8_t* img, int w, int h, size_t stride, uint64_t seed, uint64_t frame);
void orc_synth_descriptors(uint8_t* rows, int64_t first, int64_t n, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
