/*
 * orbx.h -- C ABI of liborbx_b200.so: the B200 (sm_100a) ORB front end.
 *
 * This is the drop-in boundary for the reference's ORB hot path.  Each entry point names the
 * reference interface it replaces (paths relative to the orb_slam_fusion tree).  Only plain
 * pointers and sizes cross the boundary; there are no C++ or torch types, no exceptions and no
 * exit().  All functions return ORBX_OK (0) or a negative ORBX_E_* code; outputs are written
 * only on success.
 *
 * Ownership: the caller owns every buffer it passes in; a handle owns its device memory,
 * pinned staging memory and one CUDA stream.  Threading: distinct handles may be used
 * concurrently from different host threads (the reference runs the left and right extractor
 * on two threads, src/map/frame.cc:179-182); one handle is not re-entrant, exactly like the
 * reference object (it mutates img_pyramid_).
 *
 * There is no CPU fallback: every function that computes needs a CUDA device.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_E_EMPTY (-1)       /* empty image: OrbExtractor::operator() returns -1 (orb_extractor.cc:1016) */
#define ORBX_E_ARG (-2)         /* bad argument */
#define ORBX_E_CAP (-3)         /* an output buffer is too small */
#define ORBX_E_CUDA (-4)        /* a CUDA call failed; see orbx_last_error() */
#define ORBX_E_NOMEM (-5)
#define ORBX_E_UNSUPPORTED (-6) /* geometry / parameter outside what the kernels support */

#define ORBX_MEM_HOST 0
#define ORBX_MEM_DEVICE 1
#define ORBX_MEM_HOST_ASYNC 2 /* pinned host memory; the call only enqueues, outputs are valid after orbx_sync */

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE 19 /* kEdgeThreshold: border kept around every pyramid level (orb_extractor.cc:74) */

/* The five constructor arguments of OrbExtractor (include/cam/orb_feature/orb_extractor.h:48-49;
 * read from OrbExtractor.* in settings/EuRoC.yaml:85-98 by src/tracking.cc:189-204). */
typedef struct {
  int num_feats;
  float scale_factor;
  int num_levs;
  int ini_th_fast;
  int min_th_fast;
} orbx_params;

/* Same field order and size (28 bytes) as cv::KeyPoint, so a std::vector<cv::KeyPoint> can be
 * filled with one memcpy.  pt in level-0 pixels, size = int(31*scale), angle in degrees,
 * response = FAST score, octave = pyramid level, class_id = -1 (orb_extractor.cc:834-843). */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orbx_kp;

typedef struct orbx_extractor orbx_t;

/* ------------------------------------------------------------------ extractor */

/* OrbExtractor::OrbExtractor (orb_extractor.cc:407-465).  `device` is the CUDA ordinal;
 * `max_batch` is the largest number of frames one orbx_extract_batch call may carry
 * (1 is enough for the operator() drop-in).  Device buffers are sized lazily for the first
 * image geometry and re-sized when it changes. */
int orbx_create(const orbx_params* params, int device, int max_batch, orbx_t** out);
void orbx_destroy(orbx_t* h);
/* Message for the last failing call on this handle (never NULL). */
const char* orbx_last_error(const orbx_t* h);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (orb_extractor.h:64-74) and the per-level feature quotas num_feats_per_lev_
 * (orb_extractor.cc:433-444).  Arrays hold num_levs entries; NULL pointers are skipped. */
int orbx_tables(const orbx_t* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                int* quota);
/* Upper bound on keypoints per frame: per level quota + 3 (orb_extractor.cc:713), or the 4 children
 * of every initial node where that is more (tiny quotas).  Exact once an image geometry is known,
 * conservative before. */
int orbx_max_keypoints(const orbx_t* h);

/* OrbExtractor::operator() (orb_extractor.cc:1011-1091) for one host image, blocking.
 *   img/w/h/stride : CV_8UC1 image; NULL or zero size -> ORBX_E_EMPTY.
 *   lap0, lap1     : lapping_areas[0..1] (orb_extractor.cc:1075-1076).
 *   kps, desc      : host outputs, `cap` entries / cap*32 bytes.
 *   *n             : number of keypoints; *n_mono: the reference's return value (monoIndex).
 * If cap < *n the call fails with ORBX_E_CAP and *n still reports the needed size. */
int orbx_extract(orbx_t* h, const uint8_t* img, int w, int h_, size_t stride, int lap0, int lap1,
                 orbx_kp* kps, uint8_t* desc, int cap, int* n, int* n_mono);

/* The same call in two halves, so that one host thread can have the left and the right image of a stereo pair in
 * flight on two handles at once (the reference runs them on two threads, frame.cc:179-182): _begin copies the image
 * and enqueues the whole pipeline on the handle's stream, _end waits and hands the result over exactly like
 * orbx_extract.  orbx_extract == _begin + _end. */
int orbx_extract_begin(orbx_t* h, const uint8_t* img, int w, int h_, size_t stride, int lap0, int lap1);
int orbx_extract_end(orbx_t* h, orbx_kp* kps, uint8_t* desc, int cap, int* n, int* n_mono);

/* OrbExtractor::ComputePyramid (orb_extractor.cc:1093-1117; public, orb_extractor.h:78). */
int orbx_compute_pyramid(orbx_t* h, const uint8_t* img, int w, int h_, size_t stride);

/* img_pyramid_[lev] of the last single-frame call (public member, orb_extractor.h:76, read by
 * Frame::ComputeStereoMatches, frame.cc:834,913-931).  Writes the level WITH its 19-px
 * BORDER_REFLECT_101 frame into dst ((h+38) rows of (w+38) bytes, row pitch dst_stride) and
 * returns the interior size in *w,*h.  dst == NULL only queries the size. */
int orbx_pyramid_level(orbx_t* h, int lev, uint8_t* dst, size_t dst_stride, int* w, int* h_);

/* Batched operator(): n_frames images of equal size, frame f at imgs + f*frame_stride.
 * `mem` says where ALL pointers of this call live (ORBX_MEM_HOST: pageable or pinned host
 * memory, the call blocks until the outputs are written; ORBX_MEM_HOST_ASYNC: pinned host memory,
 * the call returns once everything is enqueued so consecutive calls pipeline their copies and
 * kernels, outputs are valid after orbx_sync; ORBX_MEM_DEVICE: device memory on
 * the handle's GPU, the call only enqueues work on `stream` (NULL = the handle's stream) and
 * returns; use orbx_sync).  Outputs: kps[f*cap + i], desc[(f*cap + i)*32], n[f], n_mono[f].
 * Frames with more than cap keypoints report n[f] = -(needed) and write nothing for f.
 * n_frames may exceed max_batch: the call loops over chunks.
 * ORBX_MEM_DEVICE frames whose base address, row_stride and frame_stride are multiples of 16 bytes are
 * read IN PLACE (no copy into the handle's pyramid: level 0 is the caller's buffer); like any input of an
 * asynchronous call they must stay unchanged until the call's work has completed on the stream.  After
 * such a call the handle holds no level-0 plane: orbx_pyramid_level(0) and the level-0 image of
 * orbx_stage_download return ORBX_E_ARG until the next call that copies its frames. */
int orbx_extract_batch(orbx_t* h, const uint8_t* imgs, int n_frames, int w, int h_, size_t row_stride,
                       size_t frame_stride, int mem, int lap0, int lap1, orbx_kp* kps, uint8_t* desc,
                       int cap, int32_t* n, int32_t* n_mono, void* stream);
int orbx_sync(orbx_t* h);

/* Number of kernels the handle has launched so far (bench.py's gpu_launches claim). */
long long orbx_launch_count(const orbx_t* h);

/* Debug counter: FAST candidates the detector dropped on this handle's device since the last reset because a
 * per-tile or per-level list was full.  The capacities are upper bounds by construction (3x3 non-maximum suppression
 * keeps at most every other pixel of every other row), so the value is 0; the parity tests assert that.  Blocks until
 * the handle's streams are idle. */
int orbx_debug_dropped(orbx_t* h, long long* dropped, int reset);

/* Per-stage device time, measured with CUDA events recorded on the launching stream between the
 * stages of every enqueued chunk while profiling is on.  ms[k] accumulates milliseconds of stage k
 * (ORBX_STAGE_T_*), *chunks the number of chunks (kernel sequences) measured; `reset` clears the
 * accumulators after reading.  Blocks until the measured chunks have finished. */
#define ORBX_N_STAGES 5
#define ORBX_STAGE_T_IMPORT 0    /* copy into the padded level-0 plane */
#define ORBX_STAGE_T_PYRAMID 1   /* 7 bilinear resizes */
#define ORBX_STAGE_T_FAST_BLUR 2 /* one kernel per tile: 7x7 Gaussian + grid FAST / NMS / candidate lists */
#define ORBX_STAGE_T_OCTREE 3    /* ini/min retry filter + DistributeOctTree */
#define ORBX_STAGE_T_DESCRIBE 4  /* slot plan + IC_Angle + rBRIEF + keypoint records */
int orbx_set_profiling(orbx_t* h, int on);
int orbx_stage_times(orbx_t* h, double* ms, long long* chunks, int reset);

/* Stage intermediates of frame `frame` of the last call, for stage-wise parity tests
 * (SURVEY.md section 7 step 3).  All write to HOST memory.
 *   ORBX_STAGE_LEVEL   : u8 level pixels, w*h bytes, no border        -> *count = w*h
 *   ORBX_STAGE_BLUR    : u8 blurred level, w*h bytes                   -> *count = w*h
 *   ORBX_STAGE_CAND    : int32 (x, y, response) triples of the FAST survivors handed to the
 *                        quadtree, coordinates relative to (16,16) as orb_extractor.cc:816-823,
 *                        in UNSPECIFIED order (sort before comparing)  -> *count = triples
 *   ORBX_STAGE_SELECTED: int32 (x, y, response) of the quadtree output in the reference's
 *                        order, level coordinates                      -> *count = triples */
#define ORBX_STAGE_LEVEL 0
#define ORBX_STAGE_BLUR 1
#define ORBX_STAGE_CAND 2
#define ORBX_STAGE_SELECTED 3
int orbx_stage_download(orbx_t* h, int frame, int stage, int lev, void* dst, size_t dst_bytes,
                        int* count);

/* Deterministic synthetic frames, generated on the device for benchmarks:
 * kind 0 = blocks-v1, 1 = uniform-v1.  Frame f uses (seed, first_frame + f).  `dst` is device
 * memory, n_frames * frame_stride bytes. */
int orbx_synth_frames(int device, int kind, uint8_t* dst, int n_frames, int w, int h_, size_t row_stride,
                      size_t frame_stride, uint64_t seed, uint64_t first_frame, int shift_x,
                      uint64_t noise_seed, void* stream);

/* ------------------------------------------------------------------ matcher */

typedef struct orbm_matcher orbm_t;

int orbm_create(int device, orbm_t** out);
void orbm_destroy(orbm_t* m);
const char* orbm_last_error(const orbm_t* m);
int orbm_sync(orbm_t* m);
long long orbm_launch_count(const orbm_t* m);
/* Test / debug switches.  ORBM_OPT_CLAIM_SEQUENTIAL = 1: orbm_search_by_projection resolves the greedy claim one map
 * point at a time (the kernel very large frames fall back to) instead of 32 at a time. */
#define ORBM_OPT_CLAIM_SEQUENTIAL 1
int orbm_set_option(orbm_t* m, int option, int value);

/* ORBmatcher::DescriptorDistance (orb_matcher.cc:1877-1891) for n independent pairs:
 * out[i] = popcount(a[i] xor b[i]) over 256 bits. */
int orbm_hamming_pairs(orbm_t* m, const uint8_t* a, const uint8_t* b, int64_t n, int32_t* out, int mem,
                       void* stream);

/* Brute-force 2-nearest-neighbour search, cv::BFMatcher(NORM_HAMMING).knnMatch(q, db, k=2) as
 * called at frame.cc:1154: for each of nq queries the two nearest of nd database rows, ordered by
 * (distance, index) so the lowest index wins ties.  idx/dist are [nq][2]; idx is the database
 * row plus db_index_base (lets a shard report global rows); missing neighbours (nd < 2) are
 * idx = -1, dist = INT32_MAX. */
int orbm_knn2(orbm_t* m, const uint8_t* q, int nq, const uint8_t* db, int64_t nd, int64_t db_index_base,
              int64_t* idx, int32_t* dist, int mem, void* stream);

/* Merge n_parts partial top-2 lists ([n_parts][nq][2], e.g. the all-gathered per-GPU results of
 * a sharded database) into the global top-2 by the same (distance, index) order. */
int orbm_top2_merge(orbm_t* m, const int64_t* idx_parts, const int32_t* dist_parts, int n_parts, int nq,
                    int64_t* idx, int32_t* dist, int mem, void* stream);

/* Lowe ratio test of frame.cc:1162: accept[i] = have two neighbours && d0 < d1 * ratio
 * (float distances, double product). */
int orbm_ratio_test(orbm_t* m, const int64_t* idx, const int32_t* dist, int nq, double ratio,
                    uint8_t* accept, int mem, void* stream);

/* ---- the database sharded over the GPUs of one node (BASELINE config 5) ----
 * The reference has no multi-GPU path; this is frame.cc:1154-1162 (knnMatch(k=2) + ratio) over a database whose
 * rows are split into one contiguous slice per rank.  NCCL is bound at run time (the library does not link it): the
 * process's own libnccl.so.2 if one is loaded, else the system's; ORBX_NCCL_LIB overrides.  A host that has no NCCL
 * headers can bootstrap a communicator with the three helpers below and any channel that carries 128 bytes from
 * rank 0 to the others (MPI, a file, torch.distributed); a host that already owns an ncclComm_t passes it as is. */
#ifndef NCCL_H_
typedef struct ncclComm* ncclComm_t; /* the same declaration as nccl.h:33 */
#endif
#define ORBM_NCCL_ID_BYTES 128 /* NCCL_UNIQUE_ID_BYTES */
int orbm_nccl_unique_id(uint8_t id[ORBM_NCCL_ID_BYTES]);  /* ncclGetUniqueId, called on one rank */
/* ncclCommInitRank on `device`; collective over the n_ranks callers */
int orbm_nccl_comm_create(const uint8_t id[ORBM_NCCL_ID_BYTES], int n_ranks, int rank, int device, ncclComm_t* comm);
int orbm_nccl_comm_destroy(ncclComm_t comm);
int orbm_nccl_version(void); /* ncclGetVersion of the bound library, 0 if none can be loaded */

/* Collective over the ranks of `comm`, same q / nq / ratio on every rank: rank r holds database rows
 * [db_index_base, db_index_base + nd_local) in db_local.  Each rank finds its local top-2 per query (the kernel of
 * orbm_knn2), packs them as 64-bit keys (distance << 40 | global row), ONE ncclAllGather of nq * 16 bytes per rank
 * runs on the call's stream, and one kernel merges the gathered keys by their integer order -- the (distance, row)
 * order of knnMatch, associative, hence bit-identical to orbm_knn2 over the whole database -- and applies the ratio
 * test of frame.cc:1162.  Every rank receives the full result: idx / dist [nq][2] as orbm_knn2 (global rows),
 * accept[nq] as orbm_ratio_test (may be NULL).  comm == NULL: a single rank (no NCCL needed).  Global rows < 2^40. */
int orbm_knn2_sharded(orbm_t* m, ncclComm_t comm, const uint8_t* q, int nq, const uint8_t* db_local, int64_t nd_local,
                      int64_t db_index_base, double ratio, int64_t* idx, int32_t* dist, uint8_t* accept, int mem,
                      void* stream);

/* Row-band stereo search of Frame::ComputeStereoMatches (frame.cc:836-900): for every left
 * keypoint the right keypoint of minimum distance among those whose row band
 * [floor(yR - 2*sf[octR]), ceil(yR + 2*sf[octR])] holds int(yL), with octave within +-1 and
 * uL - max_d <= uR <= uL - min_d; best_dist starts at TH_HIGH = 100 (strict <, lowest right
 * index wins ties); best_idx = -1 where nothing beats it. */
int orbm_stereo_rowband(orbm_t* m, const orbx_kp* kl, const uint8_t* dl, int nl, const orbx_kp* kr,
                        const uint8_t* dr, int nr, const float* scale_factors, int n_levels, int n_rows,
                        float min_d, float max_d, int32_t* best_idx, int32_t* best_dist, int mem,
                        void* stream);

/* The rest of Frame::ComputeStereoMatches (frame.cc:903-985): for every left
 * keypoint with best_dist < th_orb_dist ((TH_HIGH + TH_LOW) / 2 = 75 at frame.cc:832), the 11x11 SAD
 * search over 11 shifts on the keypoint's pyramid level of the two extractors' LAST frames (the
 * reference reads orb_extractor_left_/right_->img_pyramid_, frame.cc:913-931), parabola sub-pixel
 * fit, disparity gate [min_d, max_d), and the median-distance outlier cut.  Outputs per left
 * keypoint: u_right (mvuRight), depth (mvDepth = bf / disparity), both -1 without a match, and the
 * SAD distance of accepted matches (-1 otherwise).  Both extractors must live on the matcher's device. */
int orbm_stereo_refine(orbm_t* m, const orbx_t* left, const orbx_t* right, const orbx_kp* kl, int nl, const orbx_kp* kr,
                       int nr, const int32_t* best_idx, const int32_t* best_dist, int th_orb_dist, float min_d, float max_d,
                       float bf, float* u_right, float* depth, int32_t* sad, int mem, void* stream);

/* Frame::ComputeStereoMatches (frame.cc:828-986) as ONE call for the frames the two extractors processed last with
 * orbx_extract / orbx_extract_end: their keypoints, descriptors and pyramids are still on the device, so nothing is
 * uploaded; row band search (:836-900) -> SAD refinement (:903-972) -> median cut (:974-985) run back to back on the
 * matcher's stream and one copy brings mvuRight / mvDepth back.  minZ = mb, minD = 0, maxD = bf / minZ (:853-856);
 * thOrbDist = (TH_HIGH + TH_LOW) / 2 (:832).  u_right / depth: host arrays of `cap` floats, cap >= the left extractor's
 * keypoint count (else ORBX_E_CAP); *nl returns that count. */
int orbm_stereo_matches_last(orbm_t* m, const orbx_t* left, const orbx_t* right, float bf, float mb, float* u_right,
                             float* depth, int cap, int* nl);

/* MapPoint::ComputeDistinctiveDescriptors (mappoint.cc:365-428) for a batch of
 * map points: point p owns descriptor rows [offsets[p], offsets[p+1]) of `desc`; per point the row
 * with the least median Hamming distance to all rows of the point (median = element
 * int(0.5*(N-1)) of the sorted distances incl. the 0 self-distance; the first such row wins).
 * best_idx[p] is relative to offsets[p] (-1 for a point without rows).  max_rows >= the largest
 * group (<= 6000).  With ORBX_MEM_DEVICE the caller guarantees that bound. */
int orbm_distinctive(orbm_t* m, const uint8_t* desc, const int32_t* offsets, int n_points, int max_rows, int32_t* best_idx,
                     int32_t* best_median, int mem, void* stream);

/* Frame grid geometry (frame.cc:199-204: mnMinX, mnMinY, mfGridElementWidthInv/HeightInv;
 * frame.h:40-41: FRAME_GRID_COLS = 64, FRAME_GRID_ROWS = 48). */
typedef struct {
  float min_x, min_y, inv_w, inv_h;
  int32_t cols, rows;
} orbm_grid_geom;

/* One projected map point: window centre (u, v), radius r, octave gate
 * (GetFeaturesInArea arguments, frame.cc:679-683). */
typedef struct {
  float u, v, r;
  int32_t min_level, max_level;
} orbm_window_query;

/* bestDist / bestIdx / bestLevel / bestDist2 / bestLevel2 of orb_matcher.cc:81-113
 * (256 / -1 / -1 / 256 / -1 when the window is empty). */
typedef struct {
  int32_t best_dist, best_idx, best_level, best_dist2, best_level2;
} orbm_window_result;

/* Inner loop of ORBmatcher::SearchByProjection (orb_matcher.cc:66-113): for each query, the
 * best and second-best frame keypoint among GetFeaturesInArea(u, v, r, min_level, max_level)
 * (frame.cc:679-746, candidates visited cell-major then in insertion order, strict <).
 * `skip` (n bytes, may be NULL) marks frame keypoints that are already matched
 * (orb_matcher.cc:86-87).  The greedy claim of matches between queries stays with the caller. */
int orbm_window_search(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n,
                       const orbm_grid_geom* geom, const orbm_window_query* queries,
                       const uint8_t* qdesc, int nq, const uint8_t* skip, orbm_window_result* out,
                       int mem, void* stream);

/* The same with the stereo gate the reference applies to keypoints that carry a right-image coordinate
 * (orb_matcher.cc:89-92 in SearchByProjection(Frame, MapPoints): |mTrackProjXR - mvuRight[idx]| > r * scale;
 * orb_matcher.cc:1586-1590 in SearchByProjection(CurrentFrame, LastFrame): |ur - mvuRight[i2]| > radius):
 * keypoint i with kp_u_right[i] > 0 is skipped when |q_u_right[q] - kp_u_right[i]| > q_max_err[q].
 * kp_u_right (n floats, Frame::mvuRight) may be NULL: then this is orbm_window_search. */
int orbm_window_search_stereo(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n,
                              const orbm_grid_geom* geom, const orbm_window_query* queries,
                              const uint8_t* qdesc, int nq, const uint8_t* skip, const float* kp_u_right,
                              const float* q_u_right, const float* q_max_err, orbm_window_result* out,
                              int mem, void* stream);

/* The search of ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>&, th, bRight) (orb_matcher.cc:1130-1187, the same
 * loop in Fuse(pKF, Scw, ...) :1214-1386; LocalMapping::SearchInNeighbors, loop closing) for key frames with NLeft == -1:
 * per projected map point the nearest keypoint of the window (u, v, r = th * scale[nPredictedLevel], levels
 * [nPredictedLevel-1, nPredictedLevel]) whose reprojection error passes the chi-square gate against its OWN level's
 * variance: with a right coordinate (kp_u_right[i] >= 0) (ex^2 + ey^2 + er^2) * inv_level_sigma2[octave] <= 7.8, er =
 * q_u_right[q] - kp_u_right[i]; else (ex^2 + ey^2) * inv_level_sigma2[octave] <= 5.99 (:1159-1178).  No claims: the map
 * points are independent; best_dist <= TH_LOW and the Replace / AddObservation bookkeeping (:1190-1206) stay with the
 * caller.  kp_u_right (KeyFrame::mvuRight) may be NULL (monocular).  Only best_dist / best_idx / best_level of `out`
 * are defined by the reference. */
int orbm_window_search_fuse(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                            const orbm_window_query* queries, const uint8_t* qdesc, int nq, const float* kp_u_right,
                            const float* q_u_right, const float* inv_level_sigma2, int n_levels, orbm_window_result* out,
                            int mem, void* stream);

/* ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th, bFarPoints, thFarPoints)
 * (orb_matcher.cc:42-134; caller tracking.cc:2687, Tracking::SearchLocalPoints) for a frame with Nleft == -1, the
 * WHOLE function including the greedy claim: query q is the window of one map point that passed the reference's
 * entry tests (:50-57: in view, not far, not bad), in vpMapPoints order, with r = RadiusByViewingCos * th *
 * mvScaleFactors[level] and levels [nPredictedLevel-1, nPredictedLevel] (:62-70); `skip` marks frame keypoints that
 * already hold a map point with observations when the call starts (:86-87); kp_u_right / q_u_right / q_max_err as in
 * orbm_window_search_stereo (:89-92).  The map points are taken to have observations (true for the tracker's local map
 * points), so a keypoint claimed by an earlier query is skipped by the later ones, exactly as the reference's loop does:
 * a first kernel finds the eight best keypoints of every window against the initial state, a second walks the queries in
 * order, drops the keypoints claimed in the meantime from those lists and re-scans a window only when fewer than two
 * of a full list survive.  Accept rule (:117-121):
 * best_dist <= th_high (TH_HIGH = 100) and not (best_level == best_level2 and best_dist > nnratio * best_dist2).
 * Outputs: assigned[i] = query whose map point F.mvpMapPoints[i] receives (-1: untouched), *n_matches = return value. */
int orbm_search_by_projection(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                              const orbm_window_query* queries, const uint8_t* qdesc, int nq, const uint8_t* skip,
                              const float* kp_u_right, const float* q_u_right, const float* q_max_err, int th_high,
                              float nnratio, int32_t* assigned, int32_t* n_matches, int mem, void* stream);

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (orb_matcher.cc:1518-1728;
 * Tracking::TrackWithMotionModel, tracking.cc:2192, 2203) for frames with Nleft == -1, everything after the projection
 * (:1539-1566 stays host geometry): query q = one map point of the last frame that projects into the current one, in
 * LastFrame order, with the window (u, v, radius = th * scale[nLastOctave]) and the level range of its motion case
 * (:1568-1576: forward [nLastOctave, -1], backward [0, nLastOctave], else [nLastOctave-1, nLastOctave+1]), its
 * descriptor and q_angle = the last frame's keypoint angle; q_u_right = u - bf * invzc and q_max_err = radius drive the
 * stereo gate (:1586-1590).  Each query takes the nearest keypoint that holds no map point with observations (`skip`
 * at the start, then the claims of earlier queries; strict <, first wins) if its distance is <= th_high (TH_HIGH);
 * with check_orientation the matches outside the three dominant bins of the 30-bin rotation histogram are dropped
 * (:1706-1725).  assigned[i] = query stored in CurrentFrame.mvpMapPoints[i] (-1: untouched), *n_matches = return value.
 * The relocalisation form SearchByProjection(CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) (:1730-1840,
 * tracking.cc:2978, 2992) is the same search after ITS projection: windows (u, v, th * scale[nPredictedLevel],
 * nPredictedLevel - 1 .. nPredictedLevel + 1) of pKF's map points that pass :1751-1771, q_angle = pKF->mvKeysUn[i].angle,
 * skip[i] = CurrentFrame.mvpMapPoints[i] != NULL (:1791), no stereo gate (kp_u_right = NULL), th_high = ORBdist.
 * The loop-closing forms SearchByProjection(KeyFrame* pKF, Sim3 Scw, vpPoints, [vpPointsKFs,] vpMatched, [vpMatchedKF,] th,
 * ratioHamming) (:391-488, :490-596; loopclosing.cc:671, 702, 877) map onto it as well, without the orientation check:
 * windows (u, v, th * scale[nPredictedLevel], nPredictedLevel - 1 .. nPredictedLevel) over pKF's keypoints
 * (KeyFrame::GetFeaturesInArea, keyframe.cc:729-773, visits the same grid in the same order), skip[i] = vpMatched[i] != NULL
 * (:461), th_high = TH_LOW * ratioHamming (:483), check_orientation = 0. */
int orbm_search_by_projection_last(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int n, const orbm_grid_geom* geom,
                                   const orbm_window_query* queries, const uint8_t* qdesc, const float* q_angle, int nq,
                                   const uint8_t* skip, const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                                   int th_high, int check_orientation, int32_t* assigned, int32_t* n_matches, int mem,
                                   void* stream);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches) (orb_matcher.cc:215-389; callers
 * tracking.cc:2053, 2909: TrackReferenceKeyFrame and Relocalization) for n_pairs (key frame, frame) pairs taken from
 * one pool of frames in the [frame][cap] layout orbx_extract_batch and orbv_transform produce: frame f owns
 * kps[f*cap + i], desc[(f*cap + i)*32], i < n_per_frame[f] (NULL: cap), and its FeatureVector fv_nodes / fv_begin
 * [f*cap + j], j < fv_n[f], fv_feats[f*cap + p], p < fv_total[f] (see orbv_transform).  has_point[f*cap + i] != 0:
 * the key frame's feature holds a map point that is not bad (:246-250; NULL: every feature does).  Pair p matches
 * key frame pair_kf[p] against frame pair_f[p]: per shared vocabulary node every key-frame feature with a map point
 * takes, in order, the nearest frame feature of the node that is still free (distance <= TH_LOW = 50 and
 * d1 < nnratio * d2), then (check_orientation) only the matches in the three dominant bins of the 30-bin rotation
 * histogram survive (ComputeThreeMaxima, :1841-1873).  Outputs: match[p*cap + i] = key-frame feature whose map point
 * frame feature i receives (vpMapPointMatches[i]; -1 = NULL), n_matches[p] = the return value.  cap <= 2048.
 * Covers the Nleft == -1 case (monocular and rectified stereo, the EuRoC configuration); the two-camera branches of
 * :268-291 / :329-356 stay with the reference's host code. */
int orbm_search_by_bow(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const int32_t* n_per_frame,
                       const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n, const uint32_t* fv_feats,
                       const int32_t* fv_total, const uint8_t* has_point, const int32_t* pair_kf, const int32_t* pair_f,
                       int n_pairs, float nnratio, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                       void* stream);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (orb_matcher.cc:697-815; caller
 * loopclosing.cc:571, place recognition) over the same pool / pair layout: both features of a match must hold a good
 * map point (has_point applies to both sides), the distance gate is the strict d1 < TH_LOW, every feature of key
 * frame 2 is claimed at most once, and the result is indexed by the feature of key frame 1:
 * match[p*cap + i] = feature of key frame pair_2[p] whose map point vpMatches12[i] is (-1 = NULL). */
int orbm_search_by_bow_kf(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const int32_t* n_per_frame,
                          const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n, const uint32_t* fv_feats,
                          const int32_t* fv_total, const uint8_t* has_point, const int32_t* pair_1, const int32_t* pair_2,
                          int n_pairs, float nnratio, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                          void* stream);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (orb_matcher.cc:817-1040; caller
 * LocalMapping::CreateNewMapPoints, localmapping.cc:382) for key frames with one pinhole camera (cam2_ == NULL, NLeft == -1),
 * over the pool / pair layout of orbm_search_by_bow.  has_point[f*cap + i] != 0: the feature holds a map point -- such
 * features take no part on either side (:883-886, :911; required); u_right[f*cap + i] = KeyFrame::mvuRight (>= 0: stereo
 * observation).  Per pair: pair_f12[p*9 ..] = the fundamental matrix F12 of Pinhole::EpipolarConstrain (pinhole_model.cc:
 * 116-119: K1^-T [t12]x R12 K2^-1, a product of Eigen matrices the host keeps), row-major; pair_ep[p*2 ..] = the epipole
 * pKF2->cam_->Project(T2w * Cw) (:829-830).  scale_factors / level_sigma2 [n_levels] = mvScaleFactors / mvLevelSigma2.
 * A feature of key frame 1 takes the nearest (of equally near: the last) unclaimed feature of key frame 2 in its
 * vocabulary node with distance <= TH_LOW that lies >= 10 * sqrt(scale) px from the epipole (only if neither side is a
 * stereo observation, :932-939) and passes the epipolar line test (:121-134; skipped when `coarse`); then the rotation
 * histogram (check_orientation).  match[p*cap + i] = feature of key frame 2 matched to feature i of key frame 1 (the pairs
 * of vMatchedPairs; -1: none), n_matches[p] = the return value.  cap <= 2048. */
int orbm_search_for_triangulation(orbm_t* m, const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames,
                                  const int32_t* n_per_frame, const uint32_t* fv_nodes, const int32_t* fv_begin, const int32_t* fv_n,
                                  const uint32_t* fv_feats, const int32_t* fv_total, const uint8_t* has_point, const float* u_right,
                                  const int32_t* pair_1, const int32_t* pair_2, int n_pairs, const float* pair_f12,
                                  const float* pair_ep, const float* scale_factors, const float* level_sigma2, int n_levels,
                                  int only_stereo, int coarse, int check_orientation, int32_t* match, int32_t* n_matches, int mem,
                                  void* stream);

/* Deterministic synthetic descriptors: 64-bit word j of row i is
 * splitmix64(seed ^ (4*(first+i)+j)).  `dst` is device memory. */
int orbm_synth_descriptors(int device, uint8_t* dst, int64_t first, int64_t n, uint64_t seed, void* stream);

/* Integer-pipe micro-benchmark on every SM of `device` (the matching roofline denominators):
 *   mode 0: independent popc.b32 per second;
 *   mode 1: 256-bit distances per second of the plain xor + 8 popc + add sequence;
 *   mode 2: 256-bit distances per second of the distance routine the kernels are built with. */
int orbm_popc_peak(int device, int mode, double* per_s);

/* ------------------------------------------------------------------ vocabulary (bag of words)
 *
 * Device-side replacement of the reference's ORBVocabulary = DBoW2::TemplatedVocabulary<FORB>
 * (3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h) for the calls on the data path:
 *   loadFromTextFile                       TemplatedVocabulary.h:1246-1330  (system start-up)
 *   transform(features, BowVector, FeatureVector, levelsup)   :1056-1118
 *     = Frame::ComputeBoW / KeyFrame::ComputeBoW               src/map/frame.cc:761-766
 *   transform(feature, word id, weight, node id, levelsup)     :1139-1179
 * Results are identical to the reference's, doubles included (same operations in the same order). */

typedef struct orbv_vocab orbv_t;

/* Node arrays indexed by node id as loadFromTextFile numbers them (0 = root, then file order; entry 0 of
 * every array is unused): parent id, leaf flag, 32-byte descriptor, weight.  Children are visited in
 * increasing node id; leaves receive word ids in increasing node id.  scoring: 0 L1_NORM, 1 L2_NORM,
 * 2 CHI_SQUARE, 3 KL, 4 BHATTACHARYYA, 5 DOT_PRODUCT; weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY
 * (DBoW2/BowVector.h:31-44).  All arrays are HOST memory. */
int orbv_create(int device, int k, int L, int scoring, int weighting, int32_t n_nodes, const int32_t* parent,
                const uint8_t* is_leaf, const uint8_t* desc, const double* weight, orbv_t** out);
/* The text format of loadFromTextFile / saveToTextFile ("k L scoring weighting", then one line per node:
 * "parent isLeaf d0 .. d31 weight").  An empty last line is ignored (the reference's loader appends a
 * phantom node with an unset descriptor there).  ORBX_E_ARG when the file cannot be read or the header
 * is out of the reference's accepted range (:1267-1272). */
int orbv_load_text(int device, const char* path, orbv_t** out);
void orbv_destroy(orbv_t* v);
const char* orbv_last_error(const orbv_t* v);
int orbv_info(const orbv_t* v, int* k, int* L, int* scoring, int* weighting, int32_t* n_nodes, int32_t* n_words);
int orbv_sync(orbv_t* v);
long long orbv_launch_count(const orbv_t* v);
/* largest `cap` orbv_transform accepts (the per-frame sort runs in shared memory) */
int orbv_max_features(void);

/* Per-feature descent (:1139-1179) for n descriptors: word id, word weight and the id of the node at
 * level L - levelsup on the path (0 = root when that level is <= 0; when the path ends in a leaf above that
 * level, where the reference leaves the value unset, that leaf). */
int orbv_features(orbv_t* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                  uint32_t* node_id, int mem, void* stream);

/* Batched transform(features, BowVector, FeatureVector, levelsup) over n_frames frames in the layout
 * orbx_extract_batch produces: frame f owns desc[(f*cap + i)*32], i < n_per_frame[f] (negative counts = 0;
 * n_per_frame may be NULL: every frame has cap features).  Outputs, all strided by cap per frame:
 *   bow_ids / bow_vals [f*cap + j], j < bow_n[f]   the BowVector in increasing word id
 *   fv_nodes / fv_begin [f*cap + j], j < fv_n[f]   FeatureVector node ids in increasing order and the start
 *                                                  of each node's features inside the frame's fv_feats
 *   fv_feats [f*cap + p], p < fv_total[f]          feature indices grouped by node, increasing in a group
 * Features whose word weight is 0 ("stopped") take no part (:1084).  `mem` applies to every pointer. */
int orbv_transform(orbv_t* v, const uint8_t* desc, int cap, const int32_t* n_per_frame, int n_frames, int levelsup,
                   uint32_t* bow_ids, double* bow_vals, int32_t* bow_n, uint32_t* fv_nodes, int32_t* fv_begin,
                   int32_t* fv_n, uint32_t* fv_feats, int32_t* fv_total, int mem, void* stream);

#ifdef __cplusplus
}
#endif

#endif
