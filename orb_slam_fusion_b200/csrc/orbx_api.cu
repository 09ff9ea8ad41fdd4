// orbx_api.cu -- the C ABI of liborbx_b200.so (include/orbx.h): handles, geometry, memory and
// the per-batch kernel sequence.  Host-side restatement of the OrbExtractor constructor tables
// (orb_extractor.cc:407-465) and of cv::resize's coefficient tables (A.2).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

using namespace orbx;

namespace {

#ifndef ORBX_SLOTS
#define ORBX_SLOTS 2
#endif
constexpr int kSlots = ORBX_SLOTS;  // ping-pong working sets so host-memory batches overlap copies and kernels

struct Slot {
  BatchBuffers b{};
  uint8_t* d_img = nullptr;     // staged host frames (grown on demand)
  size_t img_bytes = 0;
  orbx_kp* d_kps = nullptr;     // [max_batch][out_cap]
  uint8_t* d_desc = nullptr;    // [max_batch][out_cap][32]
  int32_t* d_n = nullptr;       // [max_batch] n, then [max_batch] n_mono
  int32_t* h_n = nullptr;       // pinned mirror of d_n
  cudaStream_t stream = nullptr;
  // host-memory batches: the frames of a chunk travel on the handle's copy stream; ev_h2d tells this slot's stream
  // that they have arrived, ev_import tells the copy stream that k_import has consumed the staging buffer
  cudaEvent_t ev_h2d = nullptr, ev_import = nullptr;
  bool import_pending = false;  // ev_import has been recorded for the chunk that last used d_img
  // Ordering between the slot's own stream and foreign streams (ORBX_MEM_DEVICE calls on a caller's stream, the
  // matcher's stream reading the pyramid in orbm_stereo_refine): every use goes through slot_acquire / slot_release.
  mutable cudaEvent_t ev_own = nullptr, ev_foreign = nullptr;
  mutable cudaStream_t foreign = nullptr;
  mutable bool foreign_pending = false;  // the last use ran on `foreign`; ev_foreign was recorded behind it
  mutable bool own_used = false;
  std::vector<void*> allocs;
};

// Before work on slot `s` is enqueued on `st`: wait for the last use of the slot's working set if it ran on another stream.
cudaError_t slot_acquire(const Slot& s, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (s.foreign_pending && s.foreign != st) e = cudaStreamWaitEvent(st, s.ev_foreign, 0);
  if (e == cudaSuccess && st != s.stream && s.own_used) {
    e = cudaEventRecord(s.ev_own, s.stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, s.ev_own, 0);
  }
  return e;
}
// After the work has been enqueued.
cudaError_t slot_release(const Slot& s, cudaStream_t st) {
  if (st == s.stream) { s.foreign_pending = false; s.own_used = true; return cudaSuccess; }
  s.foreign = st;
  s.foreign_pending = true;
  return cudaEventRecord(s.ev_foreign, st);
}

}  // namespace

constexpr int kOutSlack = 64;  // records per frame the device output rows hold beyond out_cap

struct orbx_extractor {
  orbx_params p{};
  int device = 0, max_batch = 1;
  float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS], sigma2[ORBX_MAX_LEVELS], inv_sigma2[ORBX_MAX_LEVELS];
  int quota[ORBX_MAX_LEVELS];
  FrameGeom g{};
  bool geom_valid = false;
  size_t img_pitch = 0;
  int out_cap = 0;
  Slot slot[kSlots];
  cudaStream_t copy_stream = nullptr;  // H2D of host-memory batches (keeps the copy engine busy across chunks)
  long long chunk_counter = 0;         // host-memory chunks alternate between the slots across calls too
  void* d_tables = nullptr;
  uint32_t* d_tile_tab = nullptr;
  int last_frames = 0;       // frames of the last chunk processed on slot 0
  bool border_done = false;  // REFLECT_101 frames of slot 0's pyramid are up to date
  bool level0_external = false;  // the last call on slot 0 read level 0 in place from the caller's frames: slot 0 holds no level-0 plane
  long long launches = 0;
  // single-frame path: the kernel sequence + result copies captured once as a CUDA graph
  cudaGraphExec_t graph = nullptr;
  int graph_lap0 = 0, graph_lap1 = 0;
  size_t graph_rs = 0, graph_fs = 0;
  const void* graph_img = nullptr;   // staging buffer the graph reads
  int graph_launches = 0;            // kernels one replay launches
  uint8_t* h_out = nullptr;          // pinned: out_cap keypoint records, out_cap descriptor rows, then n and n_mono
  uint8_t* d_single = nullptr;       // the same block on the device: the single-frame path returns everything in ONE copy
  // single-frame graph: the levels run as parallel branches (fork after level l exists, join before the slot plan)
  cudaStream_t fork_stream[ORBX_MAX_LEVELS] = {};
  cudaEvent_t ev_level[ORBX_MAX_LEVELS] = {}, ev_branch[ORBX_MAX_LEVELS] = {};
  bool single_pending = false;       // orbx_extract_begin has enqueued a frame that orbx_extract_end has not collected
  int single_n = -1;                 // keypoints of the last collected single frame (they stay in d_single)
  // optional per-stage timing (orbx_set_profiling): one event set per enqueued chunk
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;   // free events
  std::vector<cudaEvent_t> ev_used;   // ORBX_N_STAGES + 1 events per chunk, in order
  double stage_ms[ORBX_N_STAGES] = {0, 0, 0, 0, 0};
  long long stage_chunks = 0;
  char err[256] = "";
};

namespace {

int fail(orbx_t* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(h, call)                                                                              \
  do {                                                                                           \
    const cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return fail(h, ORBX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

int round_half_even(float v) { return (int)lrintf(v); }  // cvRound

void build_tables(orbx_t* h) {
  const int L = h->p.num_levs;
  const double sf = (double)h->p.scale_factor;  // scale_factor_ is a double member (orb_extractor.h:91)
  h->scale[0] = 1.0f;
  h->sigma2[0] = 1.0f;
  for (int i = 1; i < L; i++) {
    h->scale[i] = (float)((double)h->scale[i - 1] * sf);
    h->sigma2[i] = h->scale[i] * h->scale[i];
  }
  for (int i = 0; i < L; i++) {
    h->inv_scale[i] = 1.0f / h->scale[i];
    h->inv_sigma2[i] = 1.0f / h->sigma2[i];
  }
  const float factor = (float)(1.0 / sf);
  float per_lev = ((float)h->p.num_feats * (1 - factor)) / (1 - (float)pow((double)factor, (double)L));
  int sum = 0;
  for (int l = 0; l < L - 1; l++) {
    h->quota[l] = round_half_even(per_lev);
    sum += h->quota[l];
    per_lev *= factor;
  }
  h->quota[L - 1] = h->p.num_feats - sum > 0 ? h->p.num_feats - sum : 0;
}

short sat16(float v) {
  int r = round_half_even(v);
  return (short)(r > 32767 ? 32767 : (r < -32768 ? -32768 : r));
}

// cv::resize INTER_LINEAR 8U tables for one axis
void resize_axis(int src, int dst, int16_t* ofs, int16_t* coef, bool pair_ofs) {
  const double scale = 1.0 / ((double)dst / src);
  for (int d = 0; d < dst; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (pair_ofs) {
      // rows: the coefficients are kept, the two source rows are clamped (VResize over clipped rows)
      const int s0 = s < 0 ? 0 : (s >= src ? src - 1 : s), s1 = s + 1 < 0 ? 0 : (s + 1 >= src ? src - 1 : s + 1);
      ofs[2 * d] = (int16_t)s0;
      ofs[2 * d + 1] = (int16_t)s1;
    } else {
      if (s < 0) { f = 0; s = 0; }
      if (s >= src - 1) { f = 0; s = src - 1; }
      ofs[d] = (int16_t)s;
    }
    coef[2 * d] = sat16((1.f - f) * 2048.f);
    coef[2 * d + 1] = sat16(f * 2048.f);
  }
}

void free_slot(Slot& s) {
  for (void* p : s.allocs) cudaFree(p);
  s.allocs.clear();
  if (s.h_n) cudaFreeHost(s.h_n);
  s.h_n = nullptr;
  s.b = BatchBuffers{};
  if (s.d_img) cudaFree(s.d_img);
  s.img_bytes = 0;
  s.d_img = nullptr; s.d_kps = nullptr; s.d_desc = nullptr; s.d_n = nullptr;
  s.import_pending = false;
}

void free_geometry(orbx_t* h) {
  if (h->graph) cudaGraphExecDestroy(h->graph);
  h->graph = nullptr;
  if (h->h_out) cudaFreeHost(h->h_out);
  h->h_out = nullptr;
  if (h->d_single) cudaFree(h->d_single);
  h->d_single = nullptr;
  for (auto& s : h->slot) free_slot(s);
  if (h->d_tables) cudaFree(h->d_tables);
  h->d_tables = nullptr;
  if (h->d_tile_tab) cudaFree(h->d_tile_tab);
  h->d_tile_tab = nullptr;
  h->geom_valid = false;
}

template <class T>
cudaError_t dmalloc(Slot& s, T** p, size_t n) {
  void* v = nullptr;
  const cudaError_t e = cudaMalloc(&v, (n ? n : 1) * sizeof(T));
  if (e == cudaSuccess) { s.allocs.push_back(v); *p = (T*)v; }
  return e;
}

// TMA descriptor of one level's padded planes: u8 tensor [frames][rows][pitch], box = the tile a CTA of
// k_fast_blur (fast.cu) or k_resize_tma (pyramid.cu) fetches.  The encoder is a driver entry point; it is fetched at run time so the library
// does not link libcuda.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_map(orbx_t* h, CUtensorMap* out, const void* base, size_t width, size_t pitch, int rows, int frames, size_t frame_bytes, int box_w, int box_h) {
  static encode_tiled_fn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn)
      return fail(h, ORBX_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    enc = (encode_tiled_fn)fn;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_bytes};  // bytes, multiples of 16
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, ORBX_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return ORBX_OK;
}

int encode_plane_map(orbx_t* h, CUtensorMap* out, void* base, int pitch, int rows, int frames, size_t frame_bytes, int box_w, int box_h) {
  return encode_map(h, out, base, (size_t)pitch, (size_t)pitch, rows, frames, frame_bytes, box_w, box_h);
}

// Geometry of all levels for a w x h input, device buffers for max_batch frames per slot.
int ensure_geometry(orbx_t* h, int w, int hh) {
  if (h->geom_valid && h->g.w0 == w && h->g.h0 == hh) return ORBX_OK;
  for (auto& s : h->slot)  // asynchronous calls may still be using the buffers of the old geometry
    if (s.stream) CU(h, cudaStreamSynchronize(s.stream));
  free_geometry(h);
  if (w > 32767 || hh > 32767) return fail(h, ORBX_E_UNSUPPORTED, "image larger than 32767 px");
  FrameGeom g{};
  g.nlev = h->p.num_levs;
  g.w0 = w; g.h0 = hh;
  g.ini_th = h->p.ini_th_fast; g.min_th = h->p.min_th_fast;
  int plane = 0, cells = 0, cand = 0, sel = 0, tab = 0, tiles = 0, node_cap = 0;
  for (int l = 0; l < g.nlev; l++) {
    LevelGeom& L = g.lv[l];
    L.w = round_half_even((float)w * h->inv_scale[l]);  // orb_extractor.cc:1096
    L.h = round_half_even((float)hh * h->inv_scale[l]);
    const float width = (float)(L.w - 2 * kFastBorder), height = (float)(L.h - 2 * kFastBorder);
    L.ncols = (int)(width / 35.f);  // :759-765
    L.nrows = (int)(height / 35.f);
    if (L.w < 1 || L.h < 1 || L.ncols < 1 || L.nrows < 1)
      return fail(h, ORBX_E_UNSUPPORTED, "level %d (%dx%d) is smaller than one 35-px FAST cell plus borders", l, L.w, L.h);
    L.wcell = (int)ceilf(width / L.ncols);
    L.hcell = (int)ceilf(height / L.nrows);
    L.wcell_rcp = (uint32_t)((0x100000000ull + (uint64_t)L.wcell - 1) / (uint64_t)L.wcell);
    L.hcell_rcp = (uint32_t)((0x100000000ull + (uint64_t)L.hcell - 1) / (uint64_t)L.hcell);
    L.pitch = (kPadX + L.w + kEdge + 15) / 16 * 16;
    L.plane_off = plane;
    plane += (L.pitch * (L.h + 2 * kPadY) + 255) / 256 * 256;
    L.cell_base = cells;
    cells += L.ncols * L.nrows;
    // NMS survivors are never 8-adjacent: <= ceil(a/2)*ceil(b/2) per a x b cell domain
    const int dom_w = L.w - 2 * kEdge, dom_h = L.h - 2 * kEdge;
    L.cand_cap = ((dom_w + 1) / 2 + L.ncols) * ((dom_h + 1) / 2 + L.nrows);
    L.cand_off = cand;
    cand += (L.cand_cap + 3) / 4 * 4;
    L.quota = h->quota[l];
    L.sel_off = sel;
    L.n_roots = (int)roundf(width / height);  // :548
    if (L.n_roots < 1) return fail(h, ORBX_E_UNSUPPORTED, "level %d is taller than 2:1 (the reference divides by zero)", l);
    L.root_hx = width / L.n_roots;
    int nc = L.quota + 4;
    if (4 * L.n_roots > nc) nc = 4 * L.n_roots;
    nc += 1;
    if (nc > node_cap) node_cap = nc;
    sel += nc;  // a level's list never holds more than its node table
    L.scale = h->scale[l];
    L.scaled_patch = (int)((float)kPatch * h->scale[l]);  // :834
    L.tab_off = tab;
    tab += ((L.w > L.h ? L.w : L.h) + 7) / 8 * 8;  // keeps every level's tables 16-byte aligned
    L.blur_tiles_x = (L.w + 127) / 128;
    L.blur_tile_base = tiles;
    tiles += L.blur_tiles_x * ((L.h + kFastTileH - 1) / kFastTileH);
  }
  g.total_cells = cells;
  g.pyr_frame_bytes = plane;
  g.cand_frame_cap = cand;
  g.sel_frame_cap = sel;
  g.node_cap = node_cap;
  g.total_blur_tiles = tiles;
  if (octree_smem_bytes(node_cap) > 200 * 1024)
    return fail(h, ORBX_E_UNSUPPORTED, "per-level quota %d needs more shared memory than one SM has", node_cap);
  CU(h, octree_configure(node_cap));
  CU(h, resize_configure());
  h->launches += launch_pattern_init(h->slot[0].stream);
  CU(h, cudaStreamSynchronize(h->slot[0].stream));

  // resize tables: per level xofs[w], xalpha[2w], yofs[2h], ybeta[2h], all at tab_off
  std::vector<int16_t> t((size_t)tab * 7, 0);
  int16_t* xofs = t.data();
  int16_t* xalpha = xofs + tab;
  int16_t* yofs = xalpha + 2 * (size_t)tab;
  int16_t* ybeta = yofs + 2 * (size_t)tab;
  for (int l = 1; l < g.nlev; l++) {
    const LevelGeom &D = g.lv[l], &S = g.lv[l - 1];
    resize_axis(S.w, D.w, xofs + D.tab_off, xalpha + 2 * (size_t)D.tab_off, false);
    resize_axis(S.h, D.h, yofs + 2 * (size_t)D.tab_off, ybeta + 2 * (size_t)D.tab_off, true);
  }
  CU(h, cudaMalloc(&h->d_tables, t.size() * sizeof(int16_t)));
  CU(h, cudaMemcpy(h->d_tables, t.data(), t.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
  const int16_t* dt = (const int16_t*)h->d_tables;
  {
    std::vector<uint32_t> tt((size_t)tiles);
    for (int l = 0; l < g.nlev; l++) {
      const LevelGeom& L = g.lv[l];
      const int n = (l + 1 < g.nlev ? g.lv[l + 1].blur_tile_base : tiles) - L.blur_tile_base;
      for (int i = 0; i < n; i++)
        tt[L.blur_tile_base + i] = ((uint32_t)l << 24) | ((uint32_t)(i / L.blur_tiles_x) << 12) | (uint32_t)(i % L.blur_tiles_x);
    }
    CU(h, cudaMalloc((void**)&h->d_tile_tab, tt.size() * sizeof(uint32_t)));
    CU(h, cudaMemcpy(h->d_tile_tab, tt.data(), tt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }

  h->img_pitch = ((size_t)w + 15) / 16 * 16;
  h->out_cap = sel;
  const size_t B = (size_t)h->max_batch;
  for (auto& s : h->slot) {
    CU(h, dmalloc(s, &s.b.pyr, B * plane + 256));   // + slack: 16-byte tile loads may run past the last row
    CU(h, dmalloc(s, &s.b.blur, B * plane + 256));
    {
      CUtensorMap maps[ORBX_MAX_LEVELS];
      for (int l = 0; l < g.nlev; l++) {
        const int rc = encode_plane_map(h, &maps[l], s.b.pyr + g.lv[l].plane_off, g.lv[l].pitch, g.lv[l].h + 2 * kPadY, (int)B, (size_t)plane,
                                        kFastTileBoxW, kFastTileBoxH);
        if (rc != ORBX_OK) return rc;
      }
      for (int plan = 0; plan < 2; plan++) {  // 0: batch tile plan, 1: single-frame tile plan
        CUtensorMap rmaps[ORBX_MAX_LEVELS];
        memset(rmaps, 0, sizeof(rmaps));
        for (int l = 1; l < g.nlev; l++) {
          int th, bw, bh;
          resize_tile_plan(g, l, plan == 0 ? kRsBatchFrames : 1, &th, &bw, &bh);
          if (!bw) continue;
          const int rc = encode_plane_map(h, &rmaps[l], s.b.pyr + g.lv[l - 1].plane_off, g.lv[l - 1].pitch, g.lv[l - 1].h + 2 * kPadY, (int)B,
                                          (size_t)plane, bw, bh);
          if (rc != ORBX_OK) return rc;
        }
        CUtensorMap* drm = nullptr;
        CU(h, dmalloc(s, &drm, (size_t)g.nlev));
        CU(h, cudaMemcpy(drm, rmaps, sizeof(CUtensorMap) * g.nlev, cudaMemcpyHostToDevice));
        (plan == 0 ? s.b.rs_maps : s.b.rs_maps_single) = drm;
      }
      CUtensorMap* dm = nullptr;
      CU(h, dmalloc(s, &dm, (size_t)g.nlev));
      CU(h, cudaMemcpy(dm, maps, sizeof(CUtensorMap) * g.nlev, cudaMemcpyHostToDevice));
      s.b.pyr_maps = dm;
    }
    CU(h, dmalloc(s, &s.b.cand_raw_xy, B * cand));
    CU(h, dmalloc(s, &s.b.cand_raw_sc, B * cand));
    CU(h, dmalloc(s, &s.b.cell_strong, B * cells));
    CU(h, dmalloc(s, &s.b.cand_xy, B * cand));
    CU(h, dmalloc(s, &s.b.cand_sc, B * cand));
    CU(h, dmalloc(s, &s.b.node_of, B * cand));
    CU(h, dmalloc(s, &s.b.n_cand, B * ORBX_MAX_LEVELS));
    CU(h, dmalloc(s, &s.b.sel_xy, B * sel));
    CU(h, dmalloc(s, &s.b.sel_sc, B * sel));
    CU(h, dmalloc(s, &s.b.n_sel, B * ORBX_MAX_LEVELS));
    CU(h, dmalloc(s, &s.b.work, B * sel));
    s.b.xofs = dt;
    s.b.xalpha = dt + tab;
    s.b.yofs = dt + 3 * (size_t)tab;
    s.b.ybeta = dt + 5 * (size_t)tab;
    s.b.tile_tab = h->d_tile_tab;
    // kOutSlack extra records per frame: a caller whose row pitch (`cap`) is a little above out_cap gets its
    // pitch on the device too, so that the results go back over PCIe as ONE linear copy per chunk
    CU(h, dmalloc(s, &s.d_kps, B * (sel + kOutSlack)));
    CU(h, dmalloc(s, &s.d_desc, B * (sel + kOutSlack) * 32));
    CU(h, dmalloc(s, &s.d_n, 2 * B));
    CU(h, cudaMallocHost((void**)&s.h_n, 2 * B * sizeof(int32_t)));
    // the padding of the planes is never written by the pipeline kernels; define it once
    CU(h, cudaMemset(s.b.pyr, 0, B * plane));
    CU(h, cudaMemset(s.b.blur, 0, B * plane));
  }
  h->g = g;
  h->geom_valid = true;
  h->last_frames = 0;
  h->border_done = false;
  return ORBX_OK;
}

// The kernel sequence of OrbExtractor::operator() for `frames` device-resident frames.
void enqueue_pipeline(orbx_t* h, Slot& s, const uint8_t* d_src, size_t row_stride, size_t frame_stride, int frames,
                      int lap0, int lap1, orbx_kp* d_kps, uint8_t* d_desc, int cap, int32_t* d_n, int32_t* d_nmono,
                      int out_frame0, cudaStream_t st, cudaEvent_t after_import = nullptr, bool in_place = false) {
  FrameGeom g = h->g;
  g.lap0 = lap0;
  g.lap1 = lap1;
  if (in_place) {  // level 0 = the caller's frames (s.b.ext0_*_map describe them)
    g.ext0 = d_src;
    g.ext0_pitch = (unsigned)row_stride;
    g.ext0_frame = (unsigned long long)frame_stride;
  }
  int n = 0;
  auto mark = [&]() {
    if (!h->profiling) return;
    cudaEvent_t e;
    if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    h->ev_used.push_back(e);
  };
  // NVTX ranges around the launches of the five stages: free when no tool is attached
  mark();
  nvtxRangePushA("orbx:import");
  n += in_place ? launch_zero_counters(g, s.b, frames, st) : launch_import(g, s.b, d_src, row_stride, frame_stride, frames, st);
  nvtxRangePop();
  if (after_import) cudaEventRecord(after_import, st);  // the staging buffer may be overwritten from here on
  mark();
  nvtxRangePushA("orbx:pyramid");
  n += launch_pyramid(g, s.b, frames, st);
  nvtxRangePop();
  mark();
  nvtxRangePushA("orbx:fast_blur");
  n += launch_fast(g, s.b, frames, st);
  nvtxRangePop();
  mark();
  nvtxRangePushA("orbx:octree");
  n += launch_octree(g, s.b, frames, st);
  nvtxRangePop();
  mark();
  nvtxRangePushA("orbx:describe");
  n += launch_describe(g, s.b, frames, d_kps, d_desc, cap, d_n, d_nmono, out_frame0, st);
  nvtxRangePop();
  mark();
  h->launches += n;
}

// The same kernels for ONE frame as a forked dependency graph (recorded by stream capture): level l's detector tiles and
// quadtree need nothing but level l's plane, so they leave the resize chain as a branch as soon as that plane exists and
// all branches join before the slot plan.  The critical path drops from import + 7 resizes + detector + quadtree of ALL
// levels to import + 7 resizes + the smallest level's detector and quadtree (the big level-0 quadtree runs beside the chain).
// (the branches' streams and events exist before the capture starts: orbx_extract_begin creates them)
void enqueue_single_forked(orbx_t* h, Slot& s, const uint8_t* d_src, size_t row_stride, size_t frame_stride, int lap0, int lap1,
                           orbx_kp* d_kps, uint8_t* d_desc, int cap, int32_t* d_n, int32_t* d_nmono, cudaStream_t st) {
  FrameGeom g = h->g;
  g.lap0 = lap0;
  g.lap1 = lap1;
  // levels < n_fork leave as branches; the rest run on the trunk in one detector and one quadtree launch after the chain
  static const int fork_env = [] { const char* e = getenv("ORBX_SINGLE_FORK_LEVELS"); return e ? atoi(e) : -1; }();  // A/B knob
  // measured on B200 (tools/p50_pinned.py, 752x480, 8 levels): p50 0.1077 ms unforked, 0.1005 with 7 branches, 0.0994 with 5 --
  // the three smallest levels are cheaper as one detector + one quadtree launch behind the chain than as three more branches
  const int n_fork = fork_env >= 0 && fork_env < g.nlev ? fork_env : (g.nlev > 3 ? g.nlev - 3 : 0);
  int n = launch_import(g, s.b, d_src, row_stride, frame_stride, 1, st);
  for (int l = 0; l < g.nlev; l++) {
    if (l > 0) n += launch_resize_level(g, s.b, 1, l, st);
    if (l < n_fork) {
      cudaStream_t br = h->fork_stream[l];
      cudaEventRecord(h->ev_level[l], st);
      cudaStreamWaitEvent(br, h->ev_level[l], 0);
      n += launch_fast_levels(g, s.b, 1, l, l + 1, br);
      n += launch_octree_levels(g, s.b, 1, l, l + 1, br);
      cudaEventRecord(h->ev_branch[l], br);
    }
  }
  n += launch_fast_levels(g, s.b, 1, n_fork, g.nlev, st);
  n += launch_octree_levels(g, s.b, 1, n_fork, g.nlev, st);
  for (int l = 0; l < n_fork; l++) cudaStreamWaitEvent(st, h->ev_branch[l], 0);
  n += launch_describe(g, s.b, 1, d_kps, d_desc, cap, d_n, d_nmono, 0, st);
  h->launches += n;
}

// Host frames -> device staging buffer of slot `s`.  Rows that are (nearly) contiguous go over PCIe
// as ONE linear copy with their padding (a 2-D copy is one DMA descriptor per 752-byte row and runs
// at a fraction of the link rate); the import kernel reads any row stride.
int stage_frames(orbx_t* h, Slot& s, const uint8_t* src, int w, int hh, size_t row_stride, size_t frame_stride, int nf,
                 size_t* d_row_stride, size_t* d_frame_stride, cudaStream_t st = nullptr) {
  if (!st) st = s.stream;
  const bool linear = row_stride <= (size_t)w + 64 && (nf == 1 || frame_stride == row_stride * (size_t)hh);
  const size_t rs = linear ? row_stride : h->img_pitch, fs = rs * (size_t)hh;
  const size_t need = fs * (size_t)nf + 16;
  if (need > s.img_bytes) {
    CU(h, cudaStreamSynchronize(s.stream));
    if (st != s.stream) CU(h, cudaStreamSynchronize(st));
    if (s.d_img) cudaFree(s.d_img);
    s.d_img = nullptr;
    s.img_bytes = 0;
    s.import_pending = false;
    CU(h, cudaMalloc((void**)&s.d_img, need));
    s.img_bytes = need;
  }
  if (linear) {
    CU(h, cudaMemcpyAsync(s.d_img, src, fs * (size_t)(nf - 1) + row_stride * (size_t)(hh - 1) + (size_t)w,
                          cudaMemcpyHostToDevice, st));
  } else {
    for (int f = 0; f < nf; f++)
      CU(h, cudaMemcpy2DAsync(s.d_img + (size_t)f * fs, rs, src + (size_t)f * frame_stride, row_stride, (size_t)w,
                              (size_t)hh, cudaMemcpyHostToDevice, st));
  }
  *d_row_stride = rs;
  *d_frame_stride = fs;
  return ORBX_OK;
}

int check_image(orbx_t* h, const uint8_t* img, int w, int hh, size_t stride) {
  if (!h) return ORBX_E_ARG;
  if (!img || w <= 0 || hh <= 0) return fail(h, ORBX_E_EMPTY, "empty image");
  if (stride < (size_t)w) return fail(h, ORBX_E_ARG, "stride smaller than the row");
  return ORBX_OK;
}

}  // namespace

namespace orbx {
bool orbx_peek_pyramid(const orbx_extractor* h, FrameGeom* g, const uint8_t** pyr, const float** sf, const float** isf,
                       int* device) {
  if (!h || !h->geom_valid || h->last_frames < 1 || h->level0_external) return false;
  *g = h->g;
  *pyr = h->slot[0].b.pyr;
  *sf = h->scale;
  *isf = h->inv_scale;
  *device = h->device;
  return true;
}
// the last single frame's keypoints / descriptors as they lie on the device (d_single), for Frame::ComputeStereoMatches
bool orbx_peek_single(const orbx_extractor* h, const orbx_kp** kps, const uint8_t** desc, int* n) {
  if (!h || !h->d_single || h->single_n < 0 || h->single_pending) return false;
  *kps = reinterpret_cast<const orbx_kp*>(h->d_single);
  *desc = h->d_single + (size_t)h->out_cap * sizeof(orbx_kp);
  *n = h->single_n;
  return true;
}
// a reader on another stream (orbm_stereo_refine on the matcher's stream) brackets its kernels with these
cudaError_t orbx_pyramid_acquire(const orbx_extractor* h, cudaStream_t st) { return slot_acquire(h->slot[0], st); }
cudaError_t orbx_pyramid_release(const orbx_extractor* h, cudaStream_t st) { return slot_release(h->slot[0], st); }
}  // namespace orbx

extern "C" {

int orbx_create(const orbx_params* params, int device, int max_batch, orbx_t** out) {
  if (!params || !out || params->num_levs < 1 || params->num_levs > ORBX_MAX_LEVELS || params->num_feats < 1 ||
      !(params->scale_factor > 1.0f) || max_batch < 1)
    return ORBX_E_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ORBX_E_CUDA;
  orbx_t* h = new (std::nothrow) orbx_extractor();
  if (!h) return ORBX_E_NOMEM;
  h->p = *params;
  h->device = device;
  h->max_batch = max_batch;
  build_tables(h);
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return ORBX_E_CUDA; }
  for (auto& s : h->slot)
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.ev_own, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.ev_foreign, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.ev_import, cudaEventDisableTiming) != cudaSuccess) { orbx_destroy(h); return ORBX_E_CUDA; }
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { orbx_destroy(h); return ORBX_E_CUDA; }
  *out = h;
  return ORBX_OK;
}

void orbx_destroy(orbx_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  for (auto& s : h->slot) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.foreign_pending) cudaEventSynchronize(s.ev_foreign);
  }
  free_geometry(h);
  for (cudaEvent_t e : h->ev_used) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (int l = 0; l < ORBX_MAX_LEVELS; l++) {
    if (h->fork_stream[l]) cudaStreamDestroy(h->fork_stream[l]);
    if (h->ev_level[l]) cudaEventDestroy(h->ev_level[l]);
    if (h->ev_branch[l]) cudaEventDestroy(h->ev_branch[l]);
  }
  for (auto& s : h->slot) {
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
    if (s.ev_import) cudaEventDestroy(s.ev_import);
    if (s.ev_own) cudaEventDestroy(s.ev_own);
    if (s.ev_foreign) cudaEventDestroy(s.ev_foreign);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  delete h;
}

const char* orbx_last_error(const orbx_t* h) { return h ? h->err : "null handle"; }

int orbx_tables(const orbx_t* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota) {
  if (!h) return ORBX_E_ARG;
  const size_t L = (size_t)h->p.num_levs;
  if (scale) memcpy(scale, h->scale, L * sizeof(float));
  if (inv_scale) memcpy(inv_scale, h->inv_scale, L * sizeof(float));
  if (sigma2) memcpy(sigma2, h->sigma2, L * sizeof(float));
  if (inv_sigma2) memcpy(inv_sigma2, h->inv_sigma2, L * sizeof(float));
  if (quota) memcpy(quota, h->quota, L * sizeof(int));
  return ORBX_OK;
}

int orbx_max_keypoints(const orbx_t* h) {
  if (!h) return ORBX_E_ARG;
  if (h->geom_valid) return h->out_cap;
  int s = 0;
  for (int l = 0; l < h->p.num_levs; l++) s += (h->quota[l] + 3 > 64 ? h->quota[l] + 3 : 64) + 2;
  return s;
}

int orbx_sync(orbx_t* h) {
  if (!h) return ORBX_E_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->copy_stream));
  for (auto& s : h->slot) CU(h, cudaStreamSynchronize(s.stream));
  return ORBX_OK;
}

long long orbx_launch_count(const orbx_t* h) { return h ? h->launches : 0; }

int orbx_debug_dropped(orbx_t* h, long long* dropped, int reset) {
  if (!h || !dropped) return ORBX_E_ARG;
  CU(h, cudaSetDevice(h->device));
  for (auto& s : h->slot) CU(h, cudaStreamSynchronize(s.stream));
  unsigned int v = 0;
  CU(h, fast_dropped(&v, reset != 0));
  *dropped = (long long)v;
  return ORBX_OK;
}

int orbx_set_profiling(orbx_t* h, int on) {
  if (!h) return ORBX_E_ARG;
  h->profiling = on != 0;
  return ORBX_OK;
}

int orbx_stage_times(orbx_t* h, double* ms, long long* chunks, int reset) {
  if (!h) return ORBX_E_ARG;
  CU(h, cudaSetDevice(h->device));
  const size_t per = ORBX_N_STAGES + 1;
  for (size_t c = 0; c + per <= h->ev_used.size(); c += per) {
    CU(h, cudaEventSynchronize(h->ev_used[c + per - 1]));
    for (int k = 0; k < ORBX_N_STAGES; k++) {
      float t = 0;
      CU(h, cudaEventElapsedTime(&t, h->ev_used[c + k], h->ev_used[c + k + 1]));
      h->stage_ms[k] += t;
    }
    h->stage_chunks++;
  }
  for (cudaEvent_t e : h->ev_used) h->ev_pool.push_back(e);
  h->ev_used.clear();
  if (ms) for (int k = 0; k < ORBX_N_STAGES; k++) ms[k] = h->stage_ms[k];
  if (chunks) *chunks = h->stage_chunks;
  if (reset) {
    for (int k = 0; k < ORBX_N_STAGES; k++) h->stage_ms[k] = 0;
    h->stage_chunks = 0;
  }
  return ORBX_OK;
}

int orbx_extract_batch(orbx_t* h, const uint8_t* imgs, int n_frames, int w, int hh, size_t row_stride,
                       size_t frame_stride, int mem, int lap0, int lap1, orbx_kp* kps, uint8_t* desc, int cap,
                       int32_t* n, int32_t* n_mono, void* stream) {
  if (!h) return ORBX_E_ARG;
  if (n_frames <= 0) return fail(h, ORBX_E_EMPTY, "no frames");
  int rc = check_image(h, imgs, w, hh, row_stride);
  if (rc) return rc;
  if (!kps || !desc || !n || !n_mono || cap < 1) return fail(h, ORBX_E_ARG, "null output");
  if (mem != ORBX_MEM_HOST && mem != ORBX_MEM_DEVICE && mem != ORBX_MEM_HOST_ASYNC) return fail(h, ORBX_E_ARG, "bad mem kind");
  CU(h, cudaSetDevice(h->device));
  rc = ensure_geometry(h, w, hh);
  if (rc) return rc;
  const int B = h->max_batch;
  h->border_done = false;

  if (mem == ORBX_MEM_DEVICE) {
    Slot& s = h->slot[0];
    cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
    CU(h, slot_acquire(s, st));
    // Frames that lie on 16-byte boundaries are not copied into the pyramid slab: the detector and the first resize fetch
    // their tiles through TMA descriptors over the caller's buffer and the orientation reads it directly (level 0 "in
    // place"; the frames must stay unchanged until the call's work has completed on the stream, as for any asynchronous
    // call).  Level 1 must be one the TMA resize handles (scale factor below ~2), else the frames are imported as before.
    static const bool no_in_place = getenv("ORBX_NO_IN_PLACE") != nullptr;  // A/B knob
    int th1 = 0, bw1 = 0, bh1 = 0;
    if (h->g.nlev > 1) resize_tile_plan(h->g, 1, n_frames < B ? n_frames : B, &th1, &bw1, &bh1);
    const bool in_place = !no_in_place && ((reinterpret_cast<uintptr_t>(imgs) | row_stride | frame_stride) & 15) == 0 &&
                          row_stride <= 0xFFFFFFFFull && frame_stride >= row_stride * (size_t)hh && (h->g.nlev < 2 || bw1 != 0);  // (overlapping / repeated frames: imported)
    for (int f0 = 0; f0 < n_frames; f0 += B) {
      const int nf = n_frames - f0 < B ? n_frames - f0 : B;
      const uint8_t* src = imgs + (size_t)f0 * frame_stride;
      if (in_place) {
        if (h->g.nlev > 1) resize_tile_plan(h->g, 1, nf, &th1, &bw1, &bh1);  // the box of THIS chunk's tile plan
        rc = encode_map(h, &s.b.ext0_fast_map, src, (size_t)w, row_stride, hh, nf, frame_stride, kFastTileBoxW, kFastTileBoxH);
        if (rc == ORBX_OK && h->g.nlev > 1) rc = encode_map(h, &s.b.ext0_rs_map, src, (size_t)w, row_stride, hh, nf, frame_stride, bw1, bh1);
        if (rc != ORBX_OK) { slot_release(s, st); return rc; }
      }
      enqueue_pipeline(h, s, src, row_stride, frame_stride, nf, lap0, lap1, kps, desc, cap, n, n_mono, f0, st, nullptr, in_place);
      h->last_frames = nf;
      h->level0_external = in_place;
    }
    CU(h, slot_release(s, st));
    CU(h, cudaGetLastError());
    return ORBX_OK;
  }

  // host memory: chunks alternate between two working sets, each with its own stream for kernels and result
  // copies; the frames travel on a third stream that only waits for k_import to have consumed the staging buffer
  // of the slot (not for the slot's whole pipeline), so the H2D engine never idles between chunks
  const int dcap = cap <= h->out_cap + kOutSlack ? cap : h->out_cap;  // device row pitch of the outputs: the caller's when the rows can hold it (one linear D2H per chunk)
  for (int f0 = 0; f0 < n_frames; f0 += B) {
    Slot& s = h->slot[h->chunk_counter++ % kSlots];
    const int nf = n_frames - f0 < B ? n_frames - f0 : B;
    const uint8_t* src = imgs + (size_t)f0 * frame_stride;
    size_t drs, dfs;
    if (s.import_pending) CU(h, cudaStreamWaitEvent(h->copy_stream, s.ev_import, 0));
    rc = stage_frames(h, s, src, w, hh, row_stride, frame_stride, nf, &drs, &dfs, h->copy_stream);
    if (rc) return rc;
    CU(h, cudaEventRecord(s.ev_h2d, h->copy_stream));
    CU(h, cudaStreamWaitEvent(s.stream, s.ev_h2d, 0));
    CU(h, slot_acquire(s, s.stream));
    enqueue_pipeline(h, s, s.d_img, drs, dfs, nf, lap0, lap1, s.d_kps, s.d_desc, dcap, s.d_n, s.d_n + B, 0, s.stream, s.ev_import);
    CU(h, slot_release(s, s.stream));
    s.import_pending = true;
    if (dcap == cap) {
      CU(h, cudaMemcpyAsync(kps + (size_t)f0 * cap, s.d_kps, (size_t)nf * cap * sizeof(orbx_kp), cudaMemcpyDeviceToHost, s.stream));
      CU(h, cudaMemcpyAsync(desc + (size_t)f0 * cap * 32, s.d_desc, (size_t)nf * cap * 32, cudaMemcpyDeviceToHost, s.stream));
    } else {
      CU(h, cudaMemcpy2DAsync(kps + (size_t)f0 * cap, (size_t)cap * sizeof(orbx_kp), s.d_kps, (size_t)dcap * sizeof(orbx_kp),
                              (size_t)dcap * sizeof(orbx_kp), nf, cudaMemcpyDeviceToHost, s.stream));
      CU(h, cudaMemcpy2DAsync(desc + (size_t)f0 * cap * 32, (size_t)cap * 32, s.d_desc, (size_t)dcap * 32, (size_t)dcap * 32, nf,
                              cudaMemcpyDeviceToHost, s.stream));
    }
    CU(h, cudaMemcpyAsync(n + f0, s.d_n, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
    CU(h, cudaMemcpyAsync(n_mono + f0, s.d_n + B, nf * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
    if (&s == &h->slot[0]) { h->last_frames = nf; h->level0_external = false; }
  }
  if (mem == ORBX_MEM_HOST)
    for (auto& s : h->slot) CU(h, cudaStreamSynchronize(s.stream));
  CU(h, cudaGetLastError());
  return ORBX_OK;
}

int orbx_extract_begin(orbx_t* h, const uint8_t* img, int w, int hh, size_t stride, int lap0, int lap1) {
  if (!h) return ORBX_E_ARG;
  int rc = check_image(h, img, w, hh, stride);
  if (rc) return rc;
  h->single_pending = false;
  CU(h, cudaSetDevice(h->device));
  rc = ensure_geometry(h, w, hh);
  if (rc) return rc;
  Slot& s = h->slot[0];
  h->border_done = false;
  CU(h, slot_acquire(s, s.stream));
  CU(h, slot_release(s, s.stream));
  size_t drs, dfs;
  rc = stage_frames(h, s, img, w, hh, stride, stride * (size_t)hh, 1, &drs, &dfs);
  if (rc) return rc;
  h->last_frames = 1;
  h->level0_external = false;
  const size_t kp_bytes = (size_t)h->out_cap * sizeof(orbx_kp), desc_bytes = (size_t)h->out_cap * 32;
  const size_t single_bytes = kp_bytes + desc_bytes + 2 * sizeof(int32_t);
  if (!h->h_out) CU(h, cudaMallocHost((void**)&h->h_out, single_bytes));
  if (!h->d_single) CU(h, cudaMalloc((void**)&h->d_single, single_bytes));
  orbx_kp* const dk = reinterpret_cast<orbx_kp*>(h->d_single);
  uint8_t* const dd = h->d_single + kp_bytes;
  int32_t* const dn = reinterpret_cast<int32_t*>(h->d_single + kp_bytes + desc_bytes);  // 28 * cap + 32 * cap is a multiple of 4
  if (h->profiling) {
    // stage events cannot be timed inside a graph: plain launches
    enqueue_pipeline(h, s, s.d_img, drs, dfs, 1, lap0, lap1, dk, dd, h->out_cap, dn, dn + 1, 0, s.stream);
    CU(h, cudaMemcpyAsync(h->h_out, h->d_single, single_bytes, cudaMemcpyDeviceToHost, s.stream));
  } else {
    // 12 kernels + the result copy replayed as one graph launch (single-frame latency is
    // launch- and dependency-bound, SURVEY.md section 7 "hard parts")
    if (h->graph && (h->graph_lap0 != lap0 || h->graph_lap1 != lap1 || h->graph_rs != drs || h->graph_fs != dfs ||
                     h->graph_img != s.d_img)) {
      cudaGraphExecDestroy(h->graph);
      h->graph = nullptr;
    }
    if (!h->graph) {
      cudaGraph_t graph = nullptr;
      // streams and events of the branches are created before the capture starts
      for (int l = 0; l < h->g.nlev; l++) {
        if (!h->fork_stream[l]) CU(h, cudaStreamCreateWithFlags(&h->fork_stream[l], cudaStreamNonBlocking));
        if (!h->ev_level[l]) CU(h, cudaEventCreateWithFlags(&h->ev_level[l], cudaEventDisableTiming));
        if (!h->ev_branch[l]) CU(h, cudaEventCreateWithFlags(&h->ev_branch[l], cudaEventDisableTiming));
      }
      CU(h, cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
      const long long launches_before = h->launches;
      static const bool linear = [] { const char* e = getenv("ORBX_SINGLE_LINEAR"); return e && e[0] == '1'; }();  // A/B: the unforked chain
      if (linear) enqueue_pipeline(h, s, s.d_img, drs, dfs, 1, lap0, lap1, dk, dd, h->out_cap, dn, dn + 1, 0, s.stream);
      else enqueue_single_forked(h, s, s.d_img, drs, dfs, lap0, lap1, dk, dd, h->out_cap, dn, dn + 1, s.stream);
      cudaMemcpyAsync(h->h_out, h->d_single, single_bytes, cudaMemcpyDeviceToHost, s.stream);
      h->graph_launches = (int)(h->launches - launches_before);
      h->launches = launches_before;
      CU(h, cudaStreamEndCapture(s.stream, &graph));
      const cudaError_t ie = cudaGraphInstantiate(&h->graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) { h->graph = nullptr; return fail(h, ORBX_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
      h->graph_lap0 = lap0; h->graph_lap1 = lap1; h->graph_rs = drs; h->graph_fs = dfs; h->graph_img = s.d_img;
    }
    CU(h, cudaGraphLaunch(h->graph, s.stream));
    h->launches += h->graph_launches;
  }
  h->single_pending = true;
  return ORBX_OK;
}

int orbx_extract_end(orbx_t* h, orbx_kp* kps, uint8_t* desc, int cap, int* n, int* n_mono) {
  if (!h) return ORBX_E_ARG;
  if (!n || !n_mono) return fail(h, ORBX_E_ARG, "null output");
  if (!h->single_pending) return fail(h, ORBX_E_ARG, "orbx_extract_end without orbx_extract_begin");
  h->single_pending = false;
  CU(h, cudaSetDevice(h->device));
  Slot& s = h->slot[0];
  const size_t kp_bytes = (size_t)h->out_cap * sizeof(orbx_kp), desc_bytes = (size_t)h->out_cap * 32;
  CU(h, cudaStreamSynchronize(s.stream));
  const int32_t* const hn = reinterpret_cast<const int32_t*>(h->h_out + kp_bytes + desc_bytes);
  const int N = hn[0];
  h->single_n = N;
  if (N < 0) return fail(h, ORBX_E_UNSUPPORTED, "quadtree node table overflow");
  *n = N;
  if (N > cap) return fail(h, ORBX_E_CAP, "%d keypoints, capacity %d", N, cap);
  if (N > 0) {
    if (!kps || !desc) return fail(h, ORBX_E_ARG, "null output");
    memcpy(kps, h->h_out, (size_t)N * sizeof(orbx_kp));
    memcpy(desc, h->h_out + kp_bytes, (size_t)N * 32);
  }
  *n_mono = hn[1];
  return ORBX_OK;
}

int orbx_extract(orbx_t* h, const uint8_t* img, int w, int hh, size_t stride, int lap0, int lap1, orbx_kp* kps,
                 uint8_t* desc, int cap, int* n, int* n_mono) {
  if (!h) return ORBX_E_ARG;
  if (!n || !n_mono) return fail(h, ORBX_E_ARG, "null output");
  const int rc = orbx_extract_begin(h, img, w, hh, stride, lap0, lap1);
  return rc ? rc : orbx_extract_end(h, kps, desc, cap, n, n_mono);
}

int orbx_compute_pyramid(orbx_t* h, const uint8_t* img, int w, int hh, size_t stride) {
  if (!h) return ORBX_E_ARG;
  int rc = check_image(h, img, w, hh, stride);
  if (rc) return rc;
  CU(h, cudaSetDevice(h->device));
  rc = ensure_geometry(h, w, hh);
  if (rc) return rc;
  Slot& s = h->slot[0];
  CU(h, slot_acquire(s, s.stream));
  CU(h, slot_release(s, s.stream));
  size_t drs, dfs;
  rc = stage_frames(h, s, img, w, hh, stride, stride * (size_t)hh, 1, &drs, &dfs);
  if (rc) return rc;
  int nl = launch_import(h->g, s.b, s.d_img, drs, dfs, 1, s.stream);
  nl += launch_pyramid(h->g, s.b, 1, s.stream);
  nl += launch_border(h->g, s.b, 1, s.stream);
  h->launches += nl;
  h->last_frames = 1;
  h->level0_external = false;
  h->border_done = true;
  CU(h, cudaStreamSynchronize(s.stream));
  return ORBX_OK;
}

int orbx_pyramid_level(orbx_t* h, int lev, uint8_t* dst, size_t dst_stride, int* w, int* hh) {
  if (!h) return ORBX_E_ARG;
  if (!h->geom_valid || h->last_frames < 1) return fail(h, ORBX_E_ARG, "no pyramid has been computed");
  if (lev < 0 || lev >= h->g.nlev) return fail(h, ORBX_E_ARG, "bad level");
  const LevelGeom& L = h->g.lv[lev];
  if (w) *w = L.w;
  if (hh) *hh = L.h;
  if (!dst) return ORBX_OK;
  if (lev == 0 && h->level0_external) return fail(h, ORBX_E_ARG, "level 0 of the last call was read in place from the caller's frames");
  if (dst_stride < (size_t)(L.w + 2 * kEdge)) return fail(h, ORBX_E_ARG, "dst_stride too small");
  CU(h, cudaSetDevice(h->device));
  Slot& s = h->slot[0];
  CU(h, slot_acquire(s, s.stream));  // the pyramid may have been produced on a caller's stream (ORBX_MEM_DEVICE)
  CU(h, slot_release(s, s.stream));
  if (!h->border_done) {
    h->launches += launch_border(h->g, s.b, h->last_frames, s.stream);
    h->border_done = true;
  }
  CU(h, cudaMemcpy2DAsync(dst, dst_stride, s.b.pyr + px_off(L, -kEdge, -kEdge), (size_t)L.pitch, (size_t)(L.w + 2 * kEdge),
                          (size_t)(L.h + 2 * kEdge), cudaMemcpyDeviceToHost, s.stream));
  CU(h, cudaStreamSynchronize(s.stream));
  return ORBX_OK;
}

int orbx_stage_download(orbx_t* h, int frame, int stage, int lev, void* dst, size_t dst_bytes, int* count) {
  if (!h || !count) return ORBX_E_ARG;
  if (!h->geom_valid || frame < 0 || frame >= h->last_frames) return fail(h, ORBX_E_ARG, "no such frame in the last call");
  if (lev < 0 || lev >= h->g.nlev) return fail(h, ORBX_E_ARG, "bad level");
  CU(h, cudaSetDevice(h->device));
  Slot& s = h->slot[0];
  const FrameGeom& g = h->g;
  const LevelGeom& L = g.lv[lev];
  CU(h, slot_acquire(s, s.stream));
  CU(h, slot_release(s, s.stream));
  CU(h, cudaStreamSynchronize(s.stream));
  if (stage == ORBX_STAGE_LEVEL && lev == 0 && h->level0_external)
    return fail(h, ORBX_E_ARG, "level 0 of the last call was read in place from the caller's frames");
  if (stage == ORBX_STAGE_LEVEL || stage == ORBX_STAGE_BLUR) {
    const size_t need = (size_t)L.w * L.h;
    *count = (int)need;
    if (!dst) return ORBX_OK;
    if (dst_bytes < need) return fail(h, ORBX_E_CAP, "need %zu bytes", need);
    const uint8_t* base = (stage == ORBX_STAGE_LEVEL ? s.b.pyr : s.b.blur) + (size_t)frame * g.pyr_frame_bytes;
    CU(h, cudaMemcpy2D(dst, (size_t)L.w, base + px_off(L, 0, 0), (size_t)L.pitch, (size_t)L.w, (size_t)L.h,
                       cudaMemcpyDeviceToHost));
    return ORBX_OK;
  }
  if (stage != ORBX_STAGE_CAND && stage != ORBX_STAGE_SELECTED) return fail(h, ORBX_E_ARG, "bad stage");
  const bool cand = stage == ORBX_STAGE_CAND;
  int32_t cnt = 0;
  CU(h, cudaMemcpy(&cnt, (cand ? s.b.n_cand : s.b.n_sel) + frame * ORBX_MAX_LEVELS + lev, sizeof(int32_t),
                   cudaMemcpyDeviceToHost));
  if (cnt < 0) return fail(h, ORBX_E_UNSUPPORTED, "quadtree node table overflow");
  *count = cnt;
  if (!dst || cnt == 0) return ORBX_OK;
  if (dst_bytes < (size_t)cnt * 3 * sizeof(int32_t)) return fail(h, ORBX_E_CAP, "need %zu bytes", (size_t)cnt * 12);
  std::vector<uint32_t> xy((size_t)cnt);
  std::vector<uint8_t> sc((size_t)cnt);
  const size_t off = cand ? (size_t)frame * g.cand_frame_cap + L.cand_off : (size_t)frame * g.sel_frame_cap + L.sel_off;
  CU(h, cudaMemcpy(xy.data(), (cand ? s.b.cand_xy : s.b.sel_xy) + off, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  CU(h, cudaMemcpy(sc.data(), (cand ? s.b.cand_sc : s.b.sel_sc) + off, (size_t)cnt, cudaMemcpyDeviceToHost));
  int32_t* o = (int32_t*)dst;
  for (int i = 0; i < cnt; i++) {
    o[3 * i] = (int32_t)(xy[i] & 0xFFFFu);
    o[3 * i + 1] = (int32_t)(xy[i] >> 16);
    o[3 * i + 2] = sc[i];
  }
  return ORBX_OK;
}

int orbx_synth_frames(int device, int kind, uint8_t* dst, int n_frames, int w, int hh, size_t row_stride,
                      size_t frame_stride, uint64_t seed, uint64_t first_frame, int shift_x, uint64_t noise_seed,
                      void* stream) {
  if (!dst || n_frames < 1 || w < 2 || hh < 1 || row_stride < (size_t)w || (kind != 0 && kind != 1)) return ORBX_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return ORBX_E_CUDA;
  launch_synth(kind, dst, n_frames, w, hh, row_stride, frame_stride, seed, first_frame, shift_x, noise_seed,
               (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
}

}  // extern "C"
