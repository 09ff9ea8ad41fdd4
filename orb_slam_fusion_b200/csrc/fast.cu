// fast.cu -- 7x7 Gaussian blur and grid FAST-9 (with the iniThFAST -> minThFAST retry) fused in one
// tile kernel, all levels of all frames in one launch.
//
// Replaces the per-cell cv::FAST loop of OrbExtractor::ComputeKeyPointsOctTree
// (orb_extractor.cc:744-825) and the clone() + cv::GaussianBlur of :1054-1055.
// Arithmetic: SURVEY.md A.3 / A.6.  Both stencils need the same 128x32 tile with a 4-px halo, so
// the tile is read from HBM once.
//
// The reference calls cv::FAST on 700 overlapping ~42x44 ROIs per frame.  The detection domains of
// the cells (ROI minus cv::FAST's 3-px frame) tile [19, w-19) x [19, h-19) exactly, so the level is
// processed as one streaming stencil over aligned 128x32 tiles, and the cell structure is applied
// where it matters:
//   * non-maximum suppression treats a neighbour in another cell as 0 (a pixel on a cell edge
//     simply skips the neighbours across it);
//   * the retry rule of :799-801 ("if the cell is empty at iniThFAST, detect again at minThFAST",
//     decided after NMS) is FAST(min,nms) filtered by response >= ini when any survivor of the
//     cell reaches ini: the kernel emits every FAST(min,nms) survivor together with its cell and
//     sets a per-cell "strong" flag; the quadtree kernel applies the filter when it reads the list.
// Per tile: raw pixels (+4 halo) -> shared memory as aligned words; opposite-pair rejection test on
// four pixels per thread with byte SIMD (VABSDIFF4); survivors compacted per warp; dense 16-ring
// scoring; NMS on the shared score map; one global atomic per CTA to append.
// The list order is unspecified; the quadtree re-derives the reference order from coordinates.
#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

namespace orbx {

constexpr int kFtW = 128, kFtH = kFastTileH;   // tile of owned pixels
constexpr int kFtThreads = 8 * kFtH, kFtWarps = kFtThreads / 32;  // one warp per four owned rows
constexpr int kFtRawW = 34;                   // words per row the stencils use: bytes X0-4 .. X0+131
constexpr int kFtPitch = 4 * kFtRawW;         // 136: byte pitch of the score plane (column cb = x - X0 + 4)
constexpr int kFtRawPitch = 160;              // raw rows are ten 16-byte chunks from X0-16 (128-bit LDGSTS): column = cb + 12
constexpr int kFtRawPW = kFtRawPitch / 4, kFtRawOrg = 3;  // words per raw row; raw word of stencil word 0
constexpr int kFtRawH = kFtH + 8;             // rows Y0-4 .. Y0+35
static_assert(kFtRawPitch == kFastTileBoxW && kFtRawH == kFastTileBoxH, "TMA box of the host-side tensor maps");
constexpr int kFtScH = kFtH + 2;              // score rows Y0-1 .. Y0+32
#ifndef ORBX_FAST_STRIP
#define ORBX_FAST_STRIP 5
#endif
constexpr int kFtStrip = ORBX_FAST_STRIP;      // score rows per thread in the rejection pass: 34 columns x 7 groups x 5 rows
static_assert(kFtRawW * ((kFtScH + kFtStrip - 1) / kFtStrip) <= kFtThreads, "one thread per (word column, strip)");
constexpr int kFtMaxOut = ((kFtW + 8) / 2) * ((kFtH + 3) / 2) + 64;  // NMS survivors of one tile: at most every other pixel of every other row

// byte-wise |a - b| > t for four pixels at once; t <= 126.  VABSDIFF4 is a native instruction,
// the compare is the high-bit trick (no byte carries: (d & 0x7f) + (0x7f - t) <= 0xfe).
__device__ __forceinline__ uint32_t exceeds4(uint32_t a, uint32_t b, uint32_t k) {
  const uint32_t d = __vabsdiffu4(a, b);
  return (((d & 0x7f7f7f7fu) + k) | d) & 0x80808080u;
}

#define G4(a, b, c, d) ((uint32_t)(a) | ((uint32_t)(b) << 8) | ((uint32_t)(c) << 16) | ((uint32_t)(d) << 24))

__device__ __forceinline__ int reflect1(int p, int len) {  // reflect-101, one fold (|overshoot| < len)
  return p < 0 ? -p : (p >= len ? 2 * (len - 1) - p : p);
}

// Candidates the kernel had to drop because a list was full.  Both capacities are upper bounds by construction (at most
// every other pixel of every other row survives a 3x3 NMS), so this stays 0; the fuzz tests assert it (orbx_debug_dropped).
__device__ unsigned int g_fast_dropped = 0;

// Shared-memory accesses through a 32-bit shared-window address held in a register (see the note in the scoring loop).
template <int OFF = 0>
__device__ __forceinline__ int lds_u8(uint32_t a) {
  int v;
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// the 16 ring pixels of a candidate from the raw tile: byte loads at compile-time offsets from one 32-bit shared address
constexpr int kRingDx[16] = ORBX_RING_DX, kRingDy[16] = ORBX_RING_DY;
template <int K>
__device__ __forceinline__ void ring_load(uint32_t c, int (&r)[16]) {
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(r[K]) : "r"(c), "n"(kRingDy[K] * kFtRawPitch + kRingDx[K]));
  if constexpr (K < 15) ring_load<K + 1>(c, r);
}

#ifndef ORBX_FAST_MINB
#define ORBX_FAST_MINB 7
#endif
// q = v / d for 0 <= v < 2^15, d < 2^15 with rcp = ceil(2^32 / d) (host: LevelGeom::wcell_rcp / hcell_rcp)
__device__ __forceinline__ int div_rcp(int v, uint32_t rcp) { return (int)__umulhi((uint32_t)v, rcp); }

__global__ void __launch_bounds__(kFtThreads, ORBX_FAST_MINB * 256 / kFtThreads) k_fast_blur(const __grid_constant__ FrameGeom g, const CUtensorMap* __restrict__ pyr_maps, const __grid_constant__ CUtensorMap ext0_map,
                                              uint8_t* __restrict__ blur, uint32_t* __restrict__ cand_xy, uint8_t* __restrict__ cand_sc,
                                              int32_t* __restrict__ cand_cell, int32_t* __restrict__ n_cand,
                                              int32_t* __restrict__ cell_strong, const uint32_t* __restrict__ tile_tab) {
  // The blur's u16 intermediate and the detector's score map / lists are live in different phases
  // and share one buffer.
  constexpr int kScoreBytes = kFtScH * kFtPitch, kListBytes = 2 * kFtScH * kFtPitch, kOutBytes = 2 * kFtMaxOut;
  constexpr int kTmpBytes = 2 * (kFtH + 6) * kFtW;
  constexpr int kUnionBytes = kScoreBytes + kListBytes + kOutBytes > kTmpBytes ? kScoreBytes + kListBytes + kOutBytes : kTmpBytes;
  // ONE shared object: every array is the same base plus a constant, so the shared-window base (three uniform-datapath
  // instructions wherever a separate __shared__ array is first touched in a region) is formed once
  struct __align__(128) Smem {
    uint32_t raw_w[kFtRawH * kFtRawPW];
    uint8_t u_mem[(kUnionBytes + 15) / 16 * 16];
    unsigned long long tile_bar;  // mbarrier the TMA tile load completes on
    int n_list, n_out, out_base, warp_corners[kFtWarps];
    uint8_t xmask_l[kFtPitch], xmask_r[kFtPitch], ymask_u[kFtScH], ymask_d[kFtScH];  // 0 where the neighbour lies in another cell, else 255
  };
  __shared__ Smem sm;
  uint32_t(&raw_w)[kFtRawH * kFtRawPW] = sm.raw_w;
  unsigned long long& tile_bar = sm.tile_bar;
  uint8_t* const u_mem = sm.u_mem;
  uint8_t* score = u_mem;
  uint16_t* list = reinterpret_cast<uint16_t*>(u_mem + kScoreBytes);   // (score row << 8) | byte column of pixels to score
  uint16_t* outl = reinterpret_cast<uint16_t*>(u_mem + kScoreBytes + kListBytes);
  uint32_t* tmp2 = reinterpret_cast<uint32_t*>(u_mem);                 // blur: (kFtH + 6) / 2 row pairs x kFtW, (row 2p) | (row 2p+1) << 16
  int& n_list = sm.n_list;
  int& n_out = sm.n_out;
  int& out_base = sm.out_base;
  int(&warp_corners)[kFtWarps] = sm.warp_corners;
  uint8_t(&xmask_l)[kFtPitch] = sm.xmask_l;
  uint8_t(&xmask_r)[kFtPitch] = sm.xmask_r;
  uint8_t(&ymask_u)[kFtScH] = sm.ymask_u;
  uint8_t(&ymask_d)[kFtScH] = sm.ymask_d;

  // tile -> (level, tile column, tile row), precomputed on the host
  const uint32_t tinfo = __ldg(tile_tab + blockIdx.x);
  const int lev = (int)(tinfo >> 24), tx = (int)(tinfo & 0xFFFu), ty = (int)((tinfo >> 12) & 0xFFFu);
  const LevelGeom& L = g.lv[lev];
  const int X0 = tx * kFtW, Y0 = ty * kFtH;
  const int f = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31;
  const int lo = min(g.ini_th, g.min_th);
  const size_t fo = (size_t)f * g.pyr_frame_bytes;

  // ---- 1. raw tile: rows y = Y0-4 .. Y0+35, 160 bytes each from x = X0-16, fetched by ONE TMA tensor copy
  // (cp.async.bulk.tensor over the level's padded plane [frame][row][byte]; the part of the box beyond the
  // plane is zero-filled by the hardware).  What lies in the plane's padding is never used: the
  // BORDER_REFLECT_101 halo the blur needs at the image edges (3 px) is patched in afterwards from the tile
  // itself, only in edge tiles, and the detector stays 19 px inside.
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tile_bar);
  if (tid == 0) {
    // level 0 read in place: the tile comes straight from the caller's frames; there is no padding around them, so the box
    // coordinates are the image's own (negative ones and those beyond the image are zero-filled like the plane's padding)
    const bool ext = lev == 0 && g.ext0 != nullptr;
    const CUtensorMap* map = ext ? &ext0_map : pyr_maps + lev;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the initialised barrier is visible to the async proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(kFtRawH * kFtRawPitch) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(raw_w)),
                 "l"(map), "r"(X0 - 16 + (ext ? 0 : kPadX)), "r"(Y0 - 4 + (ext ? 0 : kPadY)), "r"(f), "r"(bar)
                 : "memory");
  }
  // (while the copy is in flight)
  if (tid < kFtPitch) {  // cell edges of the tile's columns / rows (NMS does not look across them)
    const int v = X0 - 4 + tid - kEdge + 64 * L.wcell, m = v - div_rcp(v, L.wcell_rcp) * L.wcell;
    xmask_l[tid] = m == 0 ? 0 : 255;            // first column of a cell: no left neighbour
    xmask_r[tid] = m == L.wcell - 1 ? 0 : 255;  // last column: no right neighbour
  } else if (tid < kFtPitch + kFtScH) {
    const int v = Y0 - 1 + (tid - kFtPitch) - kEdge + 64 * L.hcell, m = v - div_rcp(v, L.hcell_rcp) * L.hcell;
    ymask_u[tid - kFtPitch] = m == 0 ? 0 : 255;
    ymask_d[tid - kFtPitch] = m == L.hcell - 1 ? 0 : 255;
  }
  __syncthreads();  // every thread sees the initialised barrier
  {
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar) : "memory");
    } while (!done);
  }
  {
    uint8_t* rawb = reinterpret_cast<uint8_t*>(raw_w);
    const bool left = X0 == 0, right = X0 + kFtW + 3 > L.w, top = Y0 == 0, bottom = Y0 + kFtH + 3 > L.h;
    if (left || right) {  // columns: x = -k <- k and x = w-1+k <- w-1-k (k = 1..3), rows inside the image
      const int r = tid >> 2, k = (tid & 3) + 1, y = Y0 - 4 + r;
      if (r < kFtRawH && k <= 3 && y >= 0 && y < L.h) {
        if (left) rawb[r * kFtRawPitch + 16 - k] = rawb[r * kFtRawPitch + 16 + k];
        if (right && L.w - 1 + k <= X0 + kFtW + 2) rawb[r * kFtRawPitch + (L.w - 1 + k - X0 + 16)] = rawb[r * kFtRawPitch + (L.w - 1 - k - X0 + 16)];
      }
    }
    if (top || bottom) {
      if (left || right) __syncthreads();  // rows copy the patched columns
      // rows: y = -k <- k and y = h-1+k <- h-1-k (k = 1..3), all 34 words
      for (int i = tid; i < 3 * kFtRawPW; i += kFtThreads) {
        const int k = i / kFtRawPW + 1, c = i - (k - 1) * kFtRawPW;
        if (top) raw_w[(4 - k) * kFtRawPW + c] = raw_w[(4 + k) * kFtRawPW + c];
        if (bottom && L.h - 1 + k <= Y0 + kFtH + 2) raw_w[(L.h - 1 + k - Y0 + 4) * kFtRawPW + c] = raw_w[(L.h - 1 - k - Y0 + 4) * kFtRawPW + c];
      }
    }
    if (left || right || top || bottom) __syncthreads();
  }

  // ---- 1b. 7x7 Gaussian blur of the tile's owned pixels (clone() + GaussianBlur, orb_extractor.cc:1054-1055;
  // Q8.8 kernel [18,34,48,56,48,34,18]): horizontal pass with DP4A on funnel-shifted byte
  // windows into u16, stored as vertical PAIRS (row 2p | row 2p+1 << 16) so that the vertical pass is
  // DP2A with two taps per instruction; 32-bit stores.
  {
    const int rows_out = min(kFtH, L.h - Y0), cols_out = min(kFtW, L.w - X0);
    const int q = tid & 31;
    if (4 * q < cols_out) {  // horizontal: thread = (quad q, row pairs warp, warp+8, warp+16); blur row b = raw row b + 1
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int pr = (tid >> 5) + kFtWarps * i;
        if (2 * pr < rows_out + 6) {
          uint32_t hv[2][4];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const uint32_t* w = &raw_w[(2 * pr + e + 1) * kFtRawPW + kFtRawOrg + q];
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            // output X0+4q+j is centred on raw byte 4q+4+j (byte 0 = w0's first): taps t = 0..6 on bytes 4q+1+j+t.
            // The taps are shifted inside the coefficient words instead of shifting the pixels.
            hv[e][0] = __dp4a(w0, G4(0, 18, 34, 48), __dp4a(w1, G4(56, 48, 34, 18), 0u));
            hv[e][1] = __dp4a(w0, G4(0, 0, 18, 34), __dp4a(w1, G4(48, 56, 48, 34), __dp4a(w2, G4(18, 0, 0, 0), 0u)));
            hv[e][2] = __dp4a(w0, G4(0, 0, 0, 18), __dp4a(w1, G4(34, 48, 56, 48), __dp4a(w2, G4(34, 18, 0, 0), 0u)));
            hv[e][3] = __dp4a(w1, G4(18, 34, 48, 56), __dp4a(w2, G4(48, 34, 18, 0), 0u));  // each <= 255*256
          }
          *reinterpret_cast<uint4*>(&tmp2[pr * kFtW + 4 * q]) =
              make_uint4(__byte_perm(hv[0][0], hv[1][0], 0x5410), __byte_perm(hv[0][1], hv[1][1], 0x5410), __byte_perm(hv[0][2], hv[1][2], 0x5410),
                         __byte_perm(hv[0][3], hv[1][3], 0x5410));
        }
      }
    }
    __syncthreads();
    const int yy0 = (tid >> 5) * 4;
    if (4 * q < cols_out && yy0 < rows_out) {  // vertical: thread = (quad q, 4 output rows), 5 row pairs read once
      // output row yy0+o reads rows yy0+o .. yy0+o+6: for even o the pairs p = o/2 .. o/2+3 with taps
      // (18,34)(48,56)(48,34)(18,0), for odd o the same pairs with taps (0,18)(34,48)(56,48)(34,18)
      constexpr uint32_t ke[4] = {18u | (34u << 8), 48u | (56u << 8), 48u | (34u << 8), 18u};
      constexpr uint32_t ko[4] = {18u << 8, 34u | (48u << 8), 56u | (48u << 8), 34u | (18u << 8)};
      uint32_t acc[4][4];
#pragma unroll
      for (int o = 0; o < 4; o++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[o][k] = 32768u;
#pragma unroll
      for (int pj = 0; pj < 5; pj++) {
        const uint4 v = *reinterpret_cast<const uint4*>(&tmp2[(yy0 / 2 + pj) * kFtW + 4 * q]);
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int o = 0; o < 4; o++) {
          const int t = pj - o / 2;
          if (t >= 0 && t < 4) {
            const uint32_t kk = (o & 1) ? ko[t] : ke[t];
#pragma unroll
            for (int k = 0; k < 4; k++) acc[o][k] = __dp2a_lo(vv[k], kk, acc[o][k]);
          }
        }
      }
      const int x = X0 + 4 * q;
      uint8_t* d = blur + fo + px_off(L, x, Y0 + yy0);
#pragma unroll
      for (int o = 0; o < 4; o++) {
        if (yy0 + o < rows_out) {
          // byte 2 of each accumulator (< 2^24) is the rounded result
          const uint32_t packed = __byte_perm(__byte_perm(acc[o][0], acc[o][1], 0x0062), __byte_perm(acc[o][2], acc[o][3], 0x0062), 0x5410);
          uint8_t* dp = d + o * L.pitch;
          *reinterpret_cast<uint32_t*>(dp) = packed;  // <= 3 bytes of a straddling quad land in unused row padding
        }
      }
    }
  }
  // tiles without pixels of the detection domain [19, w-19) x [19, h-19) are done
  if (X0 >= L.w - kEdge || Y0 >= L.h - kEdge || X0 + kFtW <= kEdge || Y0 + kFtH <= kEdge) return;
  __syncthreads();  // tmp is dead: its memory becomes the score map and the lists

  if (tid == 0) { n_list = 0; n_out = 0; }
  static_assert(kScoreBytes % 16 == 0, "score map is cleared with 128-bit stores");
  for (int i = tid; i < kScoreBytes / 16; i += kFtThreads) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  // ---- 2. rejection test, four pixels per thread, + warp compaction.
  // Every 9-arc holds one pixel of each opposite pair (k, k+8), so a corner needs a ring pixel that
  // differs from the centre by more than t in the pair (0,8) AND in the pair (4,12).
  // Item (rr, wc): score row rr (y = Y0-1+rr, raw row rr+3), raw word wc (x = X0-4+4wc+j).
  // Thread (wc, rg) walks raw word column wc (x = X0-4+4wc+j) down the kFtStrip score rows of row
  // group rg (y = Y0-1+rr, raw row rr+3): column masks, neighbour indices and addresses are set up
  // once per thread, and the warp scan + shared atomic of the compaction run once per strip.
  const uint32_t kthr = lo > 126 ? 0x80808080u : (uint32_t)(0x7f - lo) * 0x01010101u;  // thresholds beyond the byte trick: everything passes
  const int xlo = max(kEdge, X0 - 1), xhi = min(L.w - kEdge, X0 + kFtW + 1);  // scored columns [xlo, xhi)
  {
    const int wc = tid % kFtRawW, rg = tid / kFtRawW;
    const int x = X0 - 4 + 4 * wc;
    // 0x80 per byte lane inside the scored column range [xlo, xhi): drop the first a and keep the first b lanes
    const int a = min(max(xlo - x, 0), 4), b = min(max(xhi - x, 0), 4);
    const uint32_t lane_mask = __funnelshift_lc(0u, 0x80808080u, 8 * a) & __funnelshift_rc(0x80808080u, 0u, 8 * (4 - b));
    const int wl = wc > 0 ? -1 : 0, wr = wc < kFtRawW - 1 ? 1 : 0;
    const int rr0 = rg * kFtStrip;
    // rows of the strip inside the score map and the detection domain: i_lo <= i < i_hi, as one bit per row.
    // Rows outside are computed all the same (their loads stay inside the CTA's shared memory) and masked
    // out: cheaper than a branch per row.
    const int y0s = Y0 - 1 + rr0, i_lo = max(kEdge - y0s, 0), i_hi = min(min(kFtScH - rr0, L.h - kEdge - y0s), kFtStrip);
    const uint32_t rowbits = i_hi > i_lo ? (1u << i_hi) - (1u << i_lo) : 0u;
    const uint32_t* row = &raw_w[(rr0 + 3) * kFtRawPW + kFtRawOrg + wc];
    uint32_t keep[kFtStrip];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < kFtStrip; i++) {
      const uint32_t m = (rowbits >> i) & 1u ? lane_mask : 0u;
      const uint32_t* rp = row + i * kFtRawPW;
      const uint32_t c = rp[0];
      const uint32_t up = rp[-3 * kFtRawPW], dn = rp[3 * kFtRawPW];
      const uint32_t lf = __byte_perm(rp[wl], c, 0x4321);   // pixels x-3
      const uint32_t rt = __byte_perm(c, rp[wr], 0x6543);   // pixels x+3
      // |a - b| > t per byte: bit 7 of (((d & 0x7f) + (0x7f - t)) | d).  The two differences of an opposite pair are OR-ed
      // BEFORE the threshold: (d0 | d1) >= max(d0, d1) bytewise, so the test can only let MORE pixels through than
      // "d0 > t or d1 > t" (for t = 2^k - 1, the EuRoC minThFAST = 7 included, exactly the same ones) -- it is a pre-filter,
      // the scoring pass decides -- and costs two thresholds per quad instead of four.
      const uint32_t m01 = __vabsdiffu4(dn, c) | __vabsdiffu4(up, c), m23 = __vabsdiffu4(rt, c) | __vabsdiffu4(lf, c);
      const uint32_t s01 = (m01 & 0x7f7f7f7fu) + kthr, s23 = (m23 & 0x7f7f7f7fu) + kthr;
      const uint32_t k = (s01 | m01) & (s23 | m23) & m;
      keep[i] = k;
      cnt += __popc(k);
    }
    if (__any_sync(0xffffffffu, cnt != 0)) {
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int wbase = 0;
      if (lane == 31) wbase = atomicAdd(&n_list, incl);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      if (cnt) {
        uint16_t* lp = &list[wbase + incl - cnt];
#pragma unroll
        for (int i = 0; i < kFtStrip; i++) {
          const uint32_t b = keep[i] >> 7;  // bit 0 / 8 / 16 / 24 = lane kept
          if (b) {
            const uint16_t val = (uint16_t)(((rr0 + i) << 8) | (4 * wc));
            if (b & 0x00000001u) *lp++ = val;
            if (b & 0x00000100u) *lp++ = val + 1;
            if (b & 0x00010000u) *lp++ = val + 2;
            if (b & 0x01000000u) *lp++ = val + 3;
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- 3. dense scoring of the compacted list (response = best - 1, 0 if not a corner at `lo`).
  // The list is cut in chunks of 32 items; warp w scores the chunks w, w+8, ... and packs the owned
  // pixels that turn out to be corners back into ITS OWN chunks (already read), densely in that order:
  // virtual slot v of warp w lives at chunk (v / 32) * 8 + w, lane v % 32.  The NMS pass below then runs
  // on full warps of corners only.
  const int nl = n_list;
  const int wrp = tid >> 5;
  int n_corner = 0;  // corners this warp has packed (warp-uniform)
  // 32-bit shared-window addresses held in registers, explicit ld / st.shared: left to generic pointers the compiler
  // re-forms the window base (five uniform-datapath instructions) at every use inside these loops
  const uint32_t sm_raw = (uint32_t)__cvta_generic_to_shared(raw_w) + 4 * kFtRawOrg + 3 * kFtRawPitch;
  const uint32_t sm_score = (uint32_t)__cvta_generic_to_shared(score), sm_list = (uint32_t)__cvta_generic_to_shared(list);
  for (int c0 = wrp * 32; c0 < nl; c0 += kFtThreads) {
    const int i = c0 + lane;
    bool corner = false;
    uint32_t item = 0;
    if (i < nl) {
      item = lds_u16(sm_list + 2u * (uint32_t)i);
      const int rr = (int)(item >> 8), cb = (int)(item & 255u);
      const uint32_t c = sm_raw + (uint32_t)(rr * kFtRawPitch + cb);
      int r[16], ctr;
      ring_load<0>(c, r);
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(ctr) : "r"(c));
      const int sc = fast9_score(ctr, r, lo);
      if (sc > 0) {
        sts_u8(sm_score + (uint32_t)(rr * kFtPitch + cb), (uint32_t)sc);
        corner = rr >= 1 && rr <= kFtH && cb >= 4 && cb < 4 + kFtW;  // not a halo pixel
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, corner);  // every lane has read its item
    if (corner) {
      const int v = n_corner + __popc(bal & ((1u << lane) - 1u));
      sts_u16(sm_list + 2u * (uint32_t)((v >> 5) * kFtThreads + (wrp << 5) + (v & 31)), item);
    }
    n_corner += __popc(bal);
  }
  if (lane == 0) warp_corners[wrp] = n_corner;
  __syncthreads();

  // ---- 4. NMS of the owned corners; neighbours across a cell edge count as 0 (byte masks per column / row).
  // The warps packed their corners separately; the CTA walks the concatenation of the eight lists, so the pass takes
  // ceil(corners / 256) rounds of full warps instead of ceil(corners of a warp / 32) rounds of every warp.
  {
    int pre[kFtWarps + 1];
    pre[0] = 0;
#pragma unroll
    for (int k = 0; k < kFtWarps; k++) pre[k + 1] = pre[k] + warp_corners[k];
    const uint32_t sm_xl = (uint32_t)__cvta_generic_to_shared(xmask_l), sm_xr = (uint32_t)__cvta_generic_to_shared(xmask_r);
    const uint32_t sm_yu = (uint32_t)__cvta_generic_to_shared(ymask_u), sm_yd = (uint32_t)__cvta_generic_to_shared(ymask_d);
    const uint32_t sm_outl = (uint32_t)__cvta_generic_to_shared(outl);
    for (int gi = tid; gi < pre[kFtWarps]; gi += kFtThreads) {
      int w = 0, off = 0;
#pragma unroll
      for (int k = 1; k < kFtWarps; k++)
        if (gi >= pre[k]) { w = k; off = pre[k]; }
      const int v = gi - off;
      const uint32_t item = lds_u16(sm_list + 2u * (uint32_t)((v >> 5) * kFtThreads + (w << 5) + (v & 31)));
      const uint32_t rr = item >> 8, cb = item & 255u;
      const uint32_t sp = sm_score + rr * kFtPitch + cb;
      const int s = lds_u8(sp);
      const int ml = lds_u8(sm_xl + cb), mr = lds_u8(sm_xr + cb), mu = lds_u8(sm_yu + rr), md = lds_u8(sm_yd + rr);
      int m = imax3(lds_u8<-1>(sp) & ml, lds_u8<1>(sp) & mr, lds_u8<-kFtPitch>(sp) & mu);
      m = imax3(m, lds_u8<kFtPitch>(sp) & md, lds_u8<-kFtPitch - 1>(sp) & (ml & mu));
      m = imax3(m, lds_u8<-kFtPitch + 1>(sp) & (mr & mu), lds_u8<kFtPitch - 1>(sp) & (ml & md));
      m = imax(m, lds_u8<kFtPitch + 1>(sp) & (mr & md));
      if (s > m) {
        const int o = atomicAdd(&n_out, 1);
        if (o < kFtMaxOut) sts_u16(sm_outl + 2u * (uint32_t)o, item);
        else atomicAdd(&g_fast_dropped, 1u);  // never expected: kFtMaxOut is the NMS bound of a tile
      }
    }
  }
  __syncthreads();

  // ---- 5. append to the level's candidate list
  const int no = min(n_out, kFtMaxOut);
  if (no == 0) return;
  if (tid == 0) out_base = atomicAdd(n_cand + f * ORBX_MAX_LEVELS + lev, no);
  __syncthreads();
  const size_t cbase = (size_t)f * g.cand_frame_cap + L.cand_off;
  for (int i = tid; i < no; i += kFtThreads) {
    const uint32_t item = lds_u16((uint32_t)__cvta_generic_to_shared(outl) + 2u * (uint32_t)i);
    const int rr = (int)(item >> 8), cb = (int)(item & 255u);
    const int s = lds_u8(sm_score + (uint32_t)(rr * kFtPitch + cb));
    const int x = X0 - 4 + cb, y = Y0 - 1 + rr;
    const int cell = div_rcp(y - kEdge, L.hcell_rcp) * L.ncols + div_rcp(x - kEdge, L.wcell_rcp);
    const int pos = out_base + i;
    if (pos >= L.cand_cap) atomicAdd(&g_fast_dropped, 1u);  // never expected: cand_cap is the NMS bound of the level
    if (pos < L.cand_cap) {
      // coordinates relative to (16,16) as orb_extractor.cc:816-823
      cand_xy[cbase + pos] = ((uint32_t)(y - kFastBorder) << 16) | (uint32_t)(x - kFastBorder);
      cand_sc[cbase + pos] = (uint8_t)s;
      cand_cell[cbase + pos] = cell;
    }
    if (s >= g.ini_th) cell_strong[(size_t)f * g.total_cells + L.cell_base + cell] = 1;
  }
}

cudaError_t fast_dropped(unsigned int* out, bool reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_fast_dropped, sizeof(unsigned int));
  if (e == cudaSuccess && reset) {
    const unsigned int zero = 0;
    e = cudaMemcpyToSymbol(g_fast_dropped, &zero, sizeof(unsigned int));
  }
  return e;
}

int launch_fast(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  // n_cand and cell_strong were zeroed by k_import, the first kernel of every pipeline
  dim3 grid(g.total_blur_tiles, 1, frames);
  k_fast_blur<<<grid, kFtThreads, 0, st>>>(g, b.pyr_maps, b.ext0_fast_map, b.blur, b.cand_raw_xy, b.cand_raw_sc, b.node_of, b.n_cand, b.cell_strong, b.tile_tab);
  return 1;
}

// levels [lev, lev_end) only (the single-frame pipeline forks per level: level l's tiles need nothing but level l's plane)
int launch_fast_levels(const FrameGeom& g, const BatchBuffers& b, int frames, int lev, int lev_end, cudaStream_t st) {
  const int first = g.lv[lev].blur_tile_base;
  const int count = (lev_end < g.nlev ? g.lv[lev_end].blur_tile_base : g.total_blur_tiles) - first;
  dim3 grid(count, 1, frames);
  k_fast_blur<<<grid, kFtThreads, 0, st>>>(g, b.pyr_maps, b.ext0_fast_map, b.blur, b.cand_raw_xy, b.cand_raw_sc, b.node_of, b.n_cand, b.cell_strong,
                                          b.tile_tab + first);
  return 1;
}

}  // namespace orbx
