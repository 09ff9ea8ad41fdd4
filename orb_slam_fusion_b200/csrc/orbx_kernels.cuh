// orbx_kernels.cuh -- internal launch interface between the C ABI (orbx_api.cu) and the kernels.
#pragma once

#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#include "orbx_geom.h"

namespace orbx {

// Device-side view of one batch's working set (all pointers are device memory).
struct BatchBuffers {
  uint8_t* pyr;         // [frames][pyr_frame_bytes]   padded level planes
  uint8_t* blur;        // [frames][pyr_frame_bytes]   blurred planes, same layout
  uint32_t* cand_raw_xy; // [frames][cand_frame_cap]   FAST(min, nms) survivors: (y << 16) | x, relative to (16,16)
  uint8_t* cand_raw_sc;  // [frames][cand_frame_cap]   FAST response
  int32_t* cell_strong;  // [frames][total_cells]      1 if a survivor of the cell reaches iniThFAST
  uint32_t* cand_xy;    // [frames][cand_frame_cap]    candidates after the ini/min retry rule
  uint8_t* cand_sc;     // [frames][cand_frame_cap]
  int32_t* node_of;     // [frames][cand_frame_cap]    cell of each raw survivor, then quadtree scratch
  int32_t* n_cand;      // [frames][ORBX_MAX_LEVELS]
  uint32_t* sel_xy;     // [frames][sel_frame_cap]     (y << 16) | x, level coordinates
  uint8_t* sel_sc;      // [frames][sel_frame_cap]
  int32_t* n_sel;       // [frames][ORBX_MAX_LEVELS]
  int32_t* work;        // [frames][sel_frame_cap]     output slot of each selected keypoint (-1: does not fit)
  // resize coefficient tables of the current geometry, per level at tab_off
  const int16_t* xofs;   // source column of destination column
  const int16_t* xalpha; // 2 coefficients per destination column
  const int16_t* yofs;   // source row of destination row (already clamped pair in yofs2)
  const int16_t* ybeta;  // 2 coefficients per destination row
  const CUtensorMap* pyr_maps;  // [nlev] TMA descriptors of the pyramid planes: u8 [frame][padded row][padded byte]
  const CUtensorMap* rs_maps;   // [nlev] entry l: planes of level l-1 with the source box of k_resize_tma's tiles (resize_tile_plan of a batch)
  const CUtensorMap* rs_maps_single;  // the same with the box of the single-frame tile plan (frames < kRsBatchFrames)
  const uint32_t* tile_tab;  // [total_blur_tiles] (level << 24) | (tile row << 12) | tile column of the 128x32 tiles
  // host copies, passed to the kernels as __grid_constant__ parameters: TMA descriptors over the CALLER's frames
  // [frame][row][byte] when level 0 is read in place (FrameGeom::ext0) -- the detector's tile box and the source box of
  // the resize of level 1
  CUtensorMap ext0_fast_map, ext0_rs_map;
};

#ifndef ORBX_FAST_TILE_H
#define ORBX_FAST_TILE_H 32
#endif
constexpr int kFastTileH = ORBX_FAST_TILE_H;  // rows of owned pixels per k_fast_blur tile (32: 256-thread CTAs, 64: 512-thread CTAs)
constexpr int kFastTileBoxW = 160, kFastTileBoxH = kFastTileH + 8;  // bytes x rows of the raw tile k_fast_blur fetches by TMA

// Each launcher enqueues on `st` and returns the number of kernels it launched.
int launch_import(const FrameGeom& g, const BatchBuffers& b, const uint8_t* src, size_t row_stride,
                  size_t frame_stride, int frames, cudaStream_t st);
int launch_zero_counters(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st);  // what k_import does besides copying (level 0 in place)
int launch_pyramid(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st);
// tile plan of the resize of level `lev`: output rows per tile, TMA box (bytes x rows) of the source tile; bw == 0: no TMA
constexpr int kRsBatchFrames = 16;  // launches of at least this many frames use the batch tile plan
void resize_tile_plan(const FrameGeom& g, int lev, int frames, int* th, int* bw, int* bh);
// per-level launchers: the single-frame pipeline runs level l's detector and quadtree as a branch beside the resize chain
int launch_resize_level(const FrameGeom& g, const BatchBuffers& b, int frames, int lev, cudaStream_t st);
int launch_fast_levels(const FrameGeom& g, const BatchBuffers& b, int frames, int lev, int lev_end, cudaStream_t st);    // levels [lev, lev_end)
int launch_octree_levels(const FrameGeom& g, const BatchBuffers& b, int frames, int lev, int lev_end, cudaStream_t st);
int launch_border(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st);
int launch_fast(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st);  // blur + FAST, one kernel
cudaError_t fast_dropped(unsigned int* out, bool reset);  // candidates dropped by k_fast_blur since the last reset (expected: 0)
int launch_octree(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st);
int launch_pattern_init(cudaStream_t st);  // fills the device-resident float pattern table (once per device / geometry)
// writes frame f of the batch to kps[(out_frame0 + f) * cap + slot], n[out_frame0 + f], ...
int launch_describe(const FrameGeom& g, const BatchBuffers& b, int frames, orbx_kp* kps, uint8_t* desc,
                    int cap, int32_t* n, int32_t* n_mono, int out_frame0, cudaStream_t st);
int launch_synth(int kind, uint8_t* dst, int frames, int w, int h, size_t row_stride, size_t frame_stride,
                 uint64_t seed, uint64_t first_frame, int shift_x, uint64_t noise_seed, cudaStream_t st);

cudaError_t resize_configure();                   // opt in to the dynamic shared memory of k_resize_tma
cudaError_t octree_configure(int node_cap);  // opt in to the dynamic shared memory the tree needs
size_t octree_smem_bytes(int node_cap);

// ---- matcher ----
int launch_hamming_pairs(const uint8_t* a, const uint8_t* b, int64_t n, int32_t* out, cudaStream_t st);
// partial top-2 per (block, query) then merge; `partials` must hold knn2_partial_bytes()
size_t knn2_partial_bytes(int nq, int64_t nd);
int launch_knn2(const uint8_t* q, int nq, const uint8_t* db, int64_t nd, int64_t index_base, void* partials,
                int64_t* idx, int32_t* dist, cudaStream_t st);
// sharded database: local top-2 as packed keys (distance << 40 | global row), then the merge of the gathered keys + ratio test
int launch_knn2_keys(const uint8_t* q, int nq, const uint8_t* db, int64_t nd, int64_t index_base, void* partials,
                     unsigned long long* keys, cudaStream_t st);
int launch_top2_keys_ratio(const unsigned long long* keys, int n_parts, int nq, double ratio, int64_t* idx, int32_t* dist,
                           uint8_t* accept, cudaStream_t st);
int launch_top2_merge(const int64_t* idx_parts, const int32_t* dist_parts, int n_parts, int nq, int64_t* idx,
                      int32_t* dist, cudaStream_t st);
int launch_ratio_test(const int64_t* idx, const int32_t* dist, int nq, double ratio, uint8_t* accept,
                      cudaStream_t st);
int launch_stereo_rowband(const orbx_kp* kl, const uint8_t* dl, int nl, const orbx_kp* kr, const uint8_t* dr,
                          int nr, const float* sf, int n_levels, int n_rows, float min_d, float max_d,
                          int32_t* best_idx, int32_t* best_dist, cudaStream_t st);
int launch_window_search(const orbx_kp* kps, const uint8_t* desc, int n, orbm_grid_geom geom,
                         const orbm_window_query* q, const uint8_t* qdesc, int nq, const uint8_t* skip,
                         const float* kp_u_right, const float* q_u_right, const float* q_max_err,
                         orbm_window_result* out, cudaStream_t st, const float* inv_sigma2 = nullptr, int n_levels = 0);  // inv_sigma2: the chi-square gate of ORBmatcher::Fuse
int launch_stereo_refine(const FrameGeom& gl, const uint8_t* pyr_l, const FrameGeom& gr, const uint8_t* pyr_r, const float* sf,
                         const float* isf, const orbx_kp* kl, int nl, const orbx_kp* kr, int nr, const int32_t* best_idx,
                         const int32_t* best_dist, int th_orb_dist, float min_d, float max_d, float bf, float* u_right,
                         float* depth, int32_t* sad, cudaStream_t st);
// the pyramid of frame 0 of an extractor's last single-frame / first-chunk call (defined in orbx_api.cu)
bool orbx_peek_pyramid(const orbx_extractor* h, FrameGeom* g, const uint8_t** pyr, const float** sf, const float** isf,
                       int* device);
bool orbx_peek_single(const orbx_extractor* h, const orbx_kp** kps, const uint8_t** desc, int* n);  // last single frame, device-resident
// order a reader of that pyramid on stream `st` after the extractor's last use, and the extractor's next use after the reader
cudaError_t orbx_pyramid_acquire(const orbx_extractor* h, cudaStream_t st);
cudaError_t orbx_pyramid_release(const orbx_extractor* h, cudaStream_t st);
// max_rows: the largest number of rows any point owns (sizes the shared memory; <= 6000)
size_t projection_scratch_bytes(int nq);  // per-query key lists of launch_search_by_projection
int launch_search_by_projection(const orbx_kp* kps, const uint8_t* desc, int n, orbm_grid_geom geom, const orbm_window_query* q,
                                const uint8_t* qdesc, int nq, const uint8_t* skip, const float* kp_u_right, const float* q_u_right,
                                const float* q_max_err, int th_high, float nnratio, const float* q_angle, int check_orientation,
                                bool last_frame, bool force_sequential, void* scratch, int32_t* assigned, int32_t* n_matches,
                                cudaStream_t st);
cudaError_t projection_configure();  // once per device: dynamic shared memory opt-in of the claim kernels
int launch_search_for_triangulation(const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const uint32_t* fv_nodes, const int32_t* fv_begin,
                                    const int32_t* fv_n, const uint32_t* fv_feats, const int32_t* fv_total, const int32_t* n_per_frame,
                                    const uint8_t* has_point, const float* u_right, const int32_t* pair_1, const int32_t* pair_2,
                                    int n_pairs, const float* pair_f12, const float* pair_ep, const float* scale_factors,
                                    const float* level_sigma2, int n_levels, int only_stereo, int coarse, int check_orientation,
                                    int32_t* match, int32_t* n_matches, cudaStream_t st);
int launch_search_by_bow(const orbx_kp* kps, const uint8_t* desc, int cap, int n_frames, const uint32_t* fv_nodes, const int32_t* fv_begin,
                         const int32_t* fv_n, const uint32_t* fv_feats, const int32_t* fv_total, const int32_t* n_per_frame,
                         const uint8_t* has_point, const int32_t* pair_1, const int32_t* pair_2, int n_pairs, float nnratio,
                         int check_orientation, bool keyframes, int32_t* match, int32_t* n_matches, cudaStream_t st);
int launch_distinctive(const uint8_t* desc, const int32_t* offsets, int n_points, int max_rows, int32_t* best_idx,
                       int32_t* best_median, cudaStream_t st);
int launch_synth_descriptors(uint8_t* dst, int64_t first, int64_t n, uint64_t seed, cudaStream_t st);
// mode 0: popc.b32 per second; 1: plain 8-popc distances per second; 2: ham256 as built, per second
int popc_bench(int mode, double* per_s);


// ---- vocabulary (bag of words) ----
// Device view of a DBoW2 vocabulary tree, re-laid out in "slots": the children of every node occupy consecutive
// slots (breadth-first), so one descent step reads <= k contiguous 32-byte rows.
struct VocabDev {
  const uint4* sdesc;     // [n_slots][2]  node descriptors
  const int2* schild;     // [n_slots]     (first child slot, number of children); 0 children = leaf
  const int32_t* snode;   // [n_slots]     node id in the reference's numbering (file order)
  const int32_t* sword;   // [n_slots]     word id of a leaf (0 for inner nodes, like Node())
  const double* sweight;  // [n_slots]
  int root_beg, root_cnt; // the root's children
  int L;                  // m_L: nominal depth, only used for the FeatureVector level (L - levelsup)
};
int bow_max_features();  // per frame (shared-memory sort)
cudaError_t bow_configure();
int launch_bow_descend(const VocabDev& v, const uint8_t* desc, int cap, const int32_t* n_per_frame, int n_frames, int levelsup,
                       uint32_t* word_id, double* weight, uint32_t* node_id, cudaStream_t st);
int launch_bow_frame(int cap, const int32_t* n_per_frame, int n_frames, int tf_weighting, int must_normalize, int l2_norm,
                     const uint32_t* word_id, const double* weight, const uint32_t* node_id, uint32_t* bow_ids, double* bow_vals,
                     int32_t* bow_n, uint32_t* fv_nodes, int32_t* fv_begin, int32_t* fv_n, uint32_t* fv_feats, int32_t* fv_total,
                     cudaStream_t st);

}  // namespace orbx
