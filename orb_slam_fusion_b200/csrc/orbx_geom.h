// orbx_geom.h -- per-image-geometry constants shared by host code and kernels (plain PODs).
//
// HBM layout of one frame slot (DESIGN.md "Data layout"):
//   pyramid slab : for each level a padded u8 plane, row pitch = multiple of 16 bytes,
//                  interior pixel (0,0) at plane + kPadY*pitch + kPadX, so interior rows start
//                  16-byte aligned and the 19-px REFLECT_101 frame fits in the padding.
//   blurred slab : same layout (only the interior is written).
//   candidates   : per level, worst-case sized SoA lists (xy packed u32, score u8) + scratch.
//   selected     : per level (quota+3) packed entries, in the reference's output order.
#pragma once

#include <stdint.h>

#include "../../include/orbx.h"

namespace orbx {

constexpr int kPadX = 32;  // left padding of every plane (>= 19, multiple of 16)
constexpr int kPadY = 19;  // top/bottom padding rows

struct LevelGeom {
  int w, h;           // interior size of this level
  int pitch;          // bytes per padded row (multiple of 16)
  int plane_off;      // byte offset of the padded plane inside a frame's pyramid slab
  int ncols, nrows;   // FAST grid (orb_extractor.cc:759-765)
  int wcell, hcell;
  int cell_base;      // cells of the lower levels (prefix sum) -> blockIdx.x decoding
  int cand_cap;       // worst-case number of NMS survivors on this level
  int cand_off;       // offset of this level's candidate list inside a frame's candidate arrays
  int quota;          // num_feats_per_lev_[lev]
  int sel_off;        // offset of this level's selected list inside a frame's (sum quota+3) list
  int n_roots;        // DistributeOctTree initial nodes (orb_extractor.cc:548)
  float root_hx;      // h_x (:550)
  float scale;        // scale_factors_[lev]
  int scaled_patch;   // int(31 * scale) (:834)
  int tab_off;        // offset of this level's resize tables (entries) in the table arrays
  int blur_tiles_x;   // 128x32 blur tiles across / tile prefix over lower levels
  int blur_tile_base;
  uint32_t wcell_rcp, hcell_rcp;  // ceil(2^32 / wcell), ceil(2^32 / hcell): exact division of small ints by one IMAD.HI
  int pad_;
};

struct FrameGeom {
  int nlev;
  int w0, h0;
  int total_cells;
  int pyr_frame_bytes;   // bytes of one frame's pyramid slab (multiple of 256)
  int cand_frame_cap;    // sum of cand_cap
  int sel_frame_cap;     // sum of (quota + 3): per-frame keypoint capacity
  int ini_th, min_th;
  int lap0, lap1;
  int node_cap;          // quadtree node-table capacity (per level problem)
  int total_blur_tiles;
  // Level 0 read IN PLACE from the caller's device frames (set per call; NULL: level 0 is the padded plane of the pyramid
  // slab, filled by k_import).  Base, row stride and frame stride are multiples of 16 bytes.
  const uint8_t* ext0;
  unsigned ext0_pitch;             // bytes per row
  unsigned long long ext0_frame;   // bytes per frame
  LevelGeom lv[ORBX_MAX_LEVELS];
};

// interior pixel (x, y) of level l inside a frame's slab
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
    int
    px_off(const LevelGeom& g, int x, int y) {
  return g.plane_off + (y + kPadY) * g.pitch + kPadX + x;  // < pyr_frame_bytes, an int
}

}  // namespace orbx
