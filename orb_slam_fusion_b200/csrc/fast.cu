// fast.cu -- grid FAST-9 with the iniThFAST -> minThFAST retry, all levels of all frames in one
// launch.
//
// Replaces the per-cell cv::FAST loop of OrbExtractor::ComputeKeyPointsOctTree
// (orb_extractor.cc:744-825).  Arithmetic: SURVEY.md A.3.
//
// One CTA per grid cell.  A cell's detection domain (its ROI minus cv::FAST's 3-px frame) never
// overlaps a neighbour's and non-maximum suppression counts pixels of other cells as 0, so a CTA
// needs its own (wcell+6) x (hcell+6) raw pixels and nothing else:
//   1. raw ROI -> shared memory (u8)
//   2. every domain pixel runs the opposite-pair rejection test at the LOW threshold; the
//      survivors of a warp are compacted with one ballot + one shared atomic per warp
//   3. the compacted list is scored densely (full 16-ring arc test, response = best - 1)
//   4. 3x3 strict NMS on the shared score map; if any NMS survivor reaches the HIGH threshold
//      only those are kept, otherwise all (FAST(20,nms) == FAST(7,nms) filtered by
//      response >= 20, and the retry of :799-801 is decided after NMS)
//   5. survivors are appended to the level's candidate list (one global atomic per CTA).
// The list order is unspecified; the quadtree re-derives the reference order from coordinates.
#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

namespace orbx {

constexpr int kFastMaxCell = 70;                   // wcell, hcell < 70 (cells are >= 35 px, < 2x)
constexpr int kFastRawPitch = kFastMaxCell + 6 + 4;  // 80
constexpr int kFastScPitch = kFastMaxCell + 2;     // score map with a zero frame

__global__ void __launch_bounds__(256) k_fast(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ pyr,
                                              uint32_t* __restrict__ cand_xy, uint8_t* __restrict__ cand_sc,
                                              int32_t* __restrict__ n_cand) {
  __shared__ __align__(16) uint8_t raw[(kFastMaxCell + 6) * kFastRawPitch];
  __shared__ __align__(16) uint8_t score[(kFastMaxCell + 2) * kFastScPitch];
  __shared__ uint16_t list[kFastMaxCell * kFastMaxCell];  // packed (yr << 8) | xr of pixels to score
  __shared__ uint16_t outl[kFastMaxCell * kFastMaxCell / 4 + 64];  // NMS survivors (<= ceil/2 x ceil/2)
  __shared__ int n_list, n_out, n_strong, out_base;

  int lev = 0;
  while (lev + 1 < g.nlev && (int)blockIdx.x >= g.lv[lev + 1].cell_base) lev++;
  const LevelGeom& L = g.lv[lev];
  const int cell = blockIdx.x - L.cell_base;
  const int ci = cell / L.ncols, cj = cell - ci * L.ncols;
  const int f = blockIdx.y;
  // detection domain of this cell in level coordinates
  const int dx0 = kEdge + cj * L.wcell, dy0 = kEdge + ci * L.hcell;
  const int dw = min(L.wcell, L.w - kEdge - dx0), dh = min(L.hcell, L.h - kEdge - dy0);
  if (dw <= 0 || dh <= 0) return;
  const int tid = threadIdx.x;
  const int lo = min(g.ini_th, g.min_th);

  if (tid == 0) { n_list = 0; n_out = 0; n_strong = 0; }
  // zero the score map (with its frame)
  for (int i = tid; i < (kFastMaxCell + 2) * kFastScPitch / 4; i += 256) reinterpret_cast<uint32_t*>(score)[i] = 0;
  // raw ROI: rows dy0-3 .. dy0+dh+2, cols dx0-3 .. dx0+dw+2
  const uint8_t* src = pyr + (size_t)f * g.pyr_frame_bytes + px_off(L, dx0 - 3, dy0 - 3);
  const int rw = dw + 6, rh = dh + 6;
  for (int i = tid; i < rh * kFastRawPitch; i += 256) {
    const int r = i / kFastRawPitch, c = i - r * kFastRawPitch;
    if (c < rw) raw[i] = __ldg(src + (size_t)r * L.pitch + c);
  }
  __syncthreads();

  // ---- 2. rejection test + warp compaction
  const int npx = dw * dh;
  for (int base = 0; base < npx; base += 256) {
    const int p = base + tid;
    bool keep = false;
    int xr = 0, yr = 0;
    if (p < npx) {
      yr = p / dw;
      xr = p - yr * dw;
      const uint8_t* c = &raw[(yr + 3) * kFastRawPitch + xr + 3];
      const int v = c[0], tb = v + lo, td = v - lo;
      // every 9-arc holds one pixel of each opposite pair (k, k+8)
      const int a0 = c[3 * kFastRawPitch], a8 = c[-3 * kFastRawPitch];  // ring 0 / 8
      const int a4 = c[3], a12 = c[-3];                                  // ring 4 / 12
      const bool bright = ((a0 > tb) | (a8 > tb)) & ((a4 > tb) | (a12 > tb));
      const bool dark = ((a0 < td) | (a8 < td)) & ((a4 < td) | (a12 < td));
      keep = bright | dark;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      int wbase = 0;
      const int lane = tid & 31;
      if (lane == 0) wbase = atomicAdd(&n_list, __popc(m));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (keep) list[wbase + __popc(m & ((1u << lane) - 1))] = (uint16_t)((yr << 8) | xr);
    }
  }
  __syncthreads();

  // ---- 3. dense scoring of the compacted list
  const int nl = n_list;
  for (int i = tid; i < nl; i += 256) {
    const int yr = list[i] >> 8, xr = list[i] & 255;
    const uint8_t* c = &raw[(yr + 3) * kFastRawPitch + xr + 3];
    const int dxs[16] = ORBX_RING_DX, dys[16] = ORBX_RING_DY;
    int r[16];
#pragma unroll
    for (int k = 0; k < 16; k++) r[k] = c[dys[k] * kFastRawPitch + dxs[k]];
    const int s = fast9_score(c[0], r, lo);
    if (s > 0) score[(yr + 1) * kFastScPitch + xr + 1] = (uint8_t)s;
  }
  __syncthreads();

  // ---- 4. NMS (only pixels on the list can be corners)
  for (int i = tid; i < nl; i += 256) {
    const int yr = list[i] >> 8, xr = list[i] & 255;
    const uint8_t* sp = &score[(yr + 1) * kFastScPitch + xr + 1];
    const int s = sp[0];
    if (s == 0) continue;
    const bool is_max = s > sp[-1] && s > sp[1] && s > sp[-kFastScPitch - 1] && s > sp[-kFastScPitch] &&
                        s > sp[-kFastScPitch + 1] && s > sp[kFastScPitch - 1] && s > sp[kFastScPitch] &&
                        s > sp[kFastScPitch + 1];
    if (is_max) {
      outl[atomicAdd(&n_out, 1)] = list[i];
      if (s >= g.ini_th) n_strong = 1;
    }
  }
  __syncthreads();

  // ---- 5. retry rule + append
  const int no = n_out;
  if (no == 0) return;
  const bool strong_only = (g.ini_th > lo) && n_strong;
  // at most ceil(70/2)^2 = 1225 survivors -> up to 5 per thread; handled in rounds of 256
  int32_t* counter = n_cand + f * ORBX_MAX_LEVELS + lev;
  const size_t cbase = (size_t)f * g.cand_frame_cap + L.cand_off;
  for (int base = 0; base < no; base += 256) {
    const int i = base + tid;
    bool emit = false;
    uint32_t xy = 0;
    int s = 0;
    if (i < no) {
      const int yr = outl[i] >> 8, xr = outl[i] & 255;
      s = score[(yr + 1) * kFastScPitch + xr + 1];
      emit = !strong_only || s >= g.ini_th;
      // coordinates relative to (16,16) as orb_extractor.cc:816-823
      xy = ((uint32_t)(3 + ci * L.hcell + yr) << 16) | (uint32_t)(3 + cj * L.wcell + xr);
    }
    const int total = __syncthreads_count(emit);
    if (total == 0) continue;
    // rank of this thread among the emitting threads of the round
    __shared__ int warp_cnt[8];
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(m);
    if (tid == 0) out_base = atomicAdd(counter, total);
    __syncthreads();
    if (emit) {
      int before = __popc(m & ((1u << lane) - 1));
      for (int w = 0; w < wid; w++) before += warp_cnt[w];
      const int pos = out_base + before;
      if (pos < L.cand_cap) {
        cand_xy[cbase + pos] = xy;
        cand_sc[cbase + pos] = (uint8_t)s;
      }
    }
    __syncthreads();
  }
}

int launch_fast(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  cudaMemsetAsync(b.n_cand, 0, sizeof(int32_t) * ORBX_MAX_LEVELS * (size_t)frames, st);
  dim3 grid(g.total_cells, frames);
  k_fast<<<grid, 256, 0, st>>>(g, b.pyr, b.cand_xy, b.cand_sc, b.n_cand);
  return 1;
}

}  // namespace orbx
