// pyramid.cu -- image pyramid: level-0 import, bilinear 8U resize chain, REFLECT_101 border.
//
// Replaces OrbExtractor::ComputePyramid (orb_extractor.cc:1093-1117: cv::resize INTER_LINEAR +
// copyMakeBorder).  Arithmetic: SURVEY.md A.2, bit-exact with OpenCV 4.13's 8-bit fixed-point path.
// Byte streams with 16-byte aligned interior rows and 32-bit packed stores; no tensor-core work.
// (The 7x7 Gaussian blur lives in fast.cu: it shares the detector's shared-memory tile.)
#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

namespace orbx {

// ------------------------------------------------------------------ import (level 0)
// Copies the caller's frames into the padded level-0 planes.  16 bytes per thread.
__global__ void __launch_bounds__(256) k_import(const __grid_constant__ FrameGeom g, uint8_t* __restrict__ pyr,
                                                const uint8_t* __restrict__ src, size_t row_stride,
                                                size_t frame_stride, int aligned16, int chunks_per_row,
                                                int32_t* __restrict__ n_cand, int32_t* __restrict__ cell_strong) {
  // one thread per 16-byte chunk of a frame (rows x chunks flattened, so narrow images still fill the CTAs)
  const int f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // The first kernel of the pipeline also zeroes the counters k_fast_blur accumulates into (two memset nodes less
  // in the single-frame graph); the grid has at least w*h/16 threads per frame, far more than cells.
  for (int c = i; c < g.total_cells; c += gridDim.x * blockDim.x) cell_strong[(size_t)f * g.total_cells + c] = 0;
  if (i < ORBX_MAX_LEVELS) n_cand[f * ORBX_MAX_LEVELS + i] = 0;
  const LevelGeom& L = g.lv[0];
  const int y = i / chunks_per_row;
  if (y >= L.h) return;
  const int x = (i - y * chunks_per_row) * 16;
  const uint8_t* s = src + (size_t)f * frame_stride + (size_t)y * row_stride + x;
  uint8_t* d = pyr + (size_t)f * g.pyr_frame_bytes + px_off(L, x, y);
  if (x + 16 <= L.w) {
    uint4 v;
    if (aligned16) {
      v = __ldg(reinterpret_cast<const uint4*>(s));
    } else {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; k++)
        w[k] = (uint32_t)s[4 * k] | ((uint32_t)s[4 * k + 1] << 8) | ((uint32_t)s[4 * k + 2] << 16) |
               ((uint32_t)s[4 * k + 3] << 24);
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(d) = v;
  } else {
    for (int k = 0; x + k < L.w; k++) d[k] = s[k];
  }
}

// Level 0 in place: nothing to copy, only the counters k_fast_blur accumulates into.
__global__ void __launch_bounds__(256) k_zero_counters(const __grid_constant__ FrameGeom g, int32_t* __restrict__ n_cand,
                                                       int32_t* __restrict__ cell_strong) {
  const int f = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.total_cells; c += gridDim.x * blockDim.x) cell_strong[(size_t)f * g.total_cells + c] = 0;
  if (blockIdx.x == 0 && threadIdx.x < ORBX_MAX_LEVELS) n_cand[f * ORBX_MAX_LEVELS + threadIdx.x] = 0;
}
int launch_zero_counters(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  k_zero_counters<<<dim3((g.total_cells + 255) / 256, frames), 256, 0, st>>>(g, b.n_cand, b.cell_strong);
  return 1;
}

int launch_import(const FrameGeom& g, const BatchBuffers& b, const uint8_t* src, size_t row_stride,
                  size_t frame_stride, int frames, cudaStream_t st) {
  const int aligned = ((reinterpret_cast<uintptr_t>(src) | row_stride | frame_stride) & 15) == 0;
  const int cpr = (g.lv[0].w + 15) / 16;
  dim3 grid((cpr * g.lv[0].h + 255) / 256, frames);
  k_import<<<grid, 256, 0, st>>>(g, b.pyr, src, row_stride, frame_stride, aligned, cpr, b.n_cand, b.cell_strong);
  return 1;
}

// ------------------------------------------------------------------ bilinear resize
// Level `lev` from level `lev-1` (orb_extractor.cc:1106), cv::resize INTER_LINEAR 8U arithmetic
// (SURVEY.md A.2): H[x] = S[sx]*a0 + S[sx+1]*a1, dst = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
// CTA = 128 output columns x `th` output rows.  The horizontal pass runs ONCE per source row the tile
// needs (adjacent output rows share source rows) and keeps H>>4 (15 bits) as u16 in shared memory;
// the vertical pass combines two shared rows per output row.  4 px per thread, 32-bit stores.
#ifndef ORBX_RS_TH
#define ORBX_RS_TH 32
#endif
constexpr int kRsTW = 128, kRsMaxTH = ORBX_RS_TH, kRsRows = 2 * kRsMaxTH + 4;
// The TMA kernel's own tile heights (dynamic shared memory).  Batches: 48 output rows per CTA -- the per-CTA set-up (column
// tables, TMA, row table) is a quarter of the kernel's instructions at 32 rows, and 48 is where its amortisation stops
// paying against occupancy (per 512 frames: 32 rows 0.479 ms, 40: 0.451, 48: 0.448, 56: 0.470, 64: 0.474).  A single frame
// has fewer tiles than the GPU has SMs, so it keeps 32-row tiles: shorter CTAs, shorter chain.
#ifndef ORBX_RS_TMA_TH
#define ORBX_RS_TMA_TH 48
#endif
#ifndef ORBX_RS_TMA_TH_SINGLE
#define ORBX_RS_TMA_TH_SINGLE 32
#endif
constexpr int kRsTmaTH = ORBX_RS_TMA_TH, kRsTmaTHSingle = ORBX_RS_TMA_TH_SINGLE;
static_assert(kRsTmaTHSingle <= kRsTmaTH && kRsTmaTH % 8 == 0, "the single-frame tile fits the batch tile's buffers");
constexpr int kRsSrcChunks = 19, kRsSrcPitch = 16 * kRsSrcChunks;  // staged source columns per tile (scale <= ~2.2)

#ifndef ORBX_RS_MINB
#define ORBX_RS_MINB 1
#endif
__global__ void __launch_bounds__(256, ORBX_RS_MINB) k_resize(const __grid_constant__ FrameGeom g, uint8_t* __restrict__ pyr,
                                                const int16_t* __restrict__ xofs, const int16_t* __restrict__ xalpha,
                                                const int16_t* __restrict__ yofs, const int16_t* __restrict__ ybeta,
                                                int lev, int th) {
  __shared__ __align__(16) uint16_t hq[kRsRows * kRsTW];
  __shared__ __align__(16) uint8_t src_sm[kRsRows * kRsSrcPitch];  // source rows, columns sx_lo ..
  const LevelGeom& D = g.lv[lev];
  const LevelGeom& S = g.lv[lev - 1];
  const int x0 = blockIdx.x * kRsTW, y0 = blockIdx.y * th;
  const int y1 = min(y0 + th, D.h);  // output rows [y0, y1)
  const int t = D.tab_off, tid = threadIdx.x;
  uint8_t* frame = pyr + (size_t)blockIdx.z * g.pyr_frame_bytes;
  // source rows [row_lo, row_hi] used by this tile (yofs holds the clamped pair of every output row)
  const int row_lo = yofs[2 * (t + y0)], row_hi = yofs[2 * (t + y1 - 1) + 1];
  const int n_rows = min(row_hi - row_lo + 1, kRsRows);
  const int src_base = px_off(S, 0, row_lo);

  // ---- source rows of the tile -> shared memory with 16-byte LDGSTS (cp.async): one asynchronous,
  // coalesced round trip instead of dependent single-byte gathers from global memory
  const int sx_lo = xofs[t + x0] & ~15;                                  // first source column, 16-byte aligned
  const int sx_hi = xofs[t + min(x0 + kRsTW, D.w) - 1] + 2;             // one past the last byte read
  const int n_chunks = (sx_hi - sx_lo + 15) >> 4;
  const bool staged = n_chunks <= kRsSrcChunks;                           // scale factors beyond ~2.2 gather from global
  if (staged) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(src_sm);
    const uint8_t* gsrc = frame + src_base + sx_lo;
    for (int i = tid; i < n_rows * n_chunks; i += 256) {
      const int r = i / n_chunks, c = i - r * n_chunks;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sbase + (unsigned)(r * kRsSrcPitch + 16 * c)),
                   "l"(gsrc + r * S.pitch + 16 * c));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }

  // ---- horizontal pass: thread = (quad q of output columns, up to 5 source rows); the column tables
  // are read once per thread (while the LDGSTS copies are in flight)
  {
    const int rpg = (n_rows + 7) >> 3;  // source rows per warp: the 8 warps share the rows evenly
    const int q = tid & 31, r0 = (tid >> 5) * rpg;
    const int dx0 = x0 + 4 * q;
    const bool active = dx0 < D.w && r0 < n_rows;
    int sxs[4] = {0, 0, 0, 0}, a0s[4] = {0, 0, 0, 0}, a1s[4] = {0, 0, 0, 0};
    if (active) {
      if (dx0 + 4 <= D.w) {
        // tables are 8-byte (xofs) / 16-byte (xalpha) aligned at multiples of 4 columns: tab_off % 8 == 0
        const short4 sx = *reinterpret_cast<const short4*>(xofs + t + dx0);
        const int4 al = *reinterpret_cast<const int4*>(xalpha + 2 * (t + dx0));  // (a0 | a1 << 16) per column
        sxs[0] = sx.x; sxs[1] = sx.y; sxs[2] = sx.z; sxs[3] = sx.w;
        const int als[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { a0s[k] = (int)(int16_t)(als[k] & 0xFFFF); a1s[k] = als[k] >> 16; }
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (dx0 + k < D.w) {
            sxs[k] = xofs[t + dx0 + k];
            a0s[k] = xalpha[2 * (t + dx0 + k)];
            a1s[k] = xalpha[2 * (t + dx0 + k) + 1];
          }
        }
      }
    }
    if (staged) {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      __syncthreads();
    }
    if (active) {
      // a1 == 0 where sx is the last column, so sx+1 may read the (allocated) padding
      if (staged) {
        // constant row pitch: the 8 byte addresses are formed once, rows are immediate offsets
        const uint8_t* sp = src_sm + r0 * kRsSrcPitch - sx_lo;
        const uint8_t* p0 = sp + sxs[0];
        const uint8_t* p1 = sp + sxs[1];
        const uint8_t* p2 = sp + sxs[2];
        const uint8_t* p3 = sp + sxs[3];
#pragma unroll
        for (int i = 0; i < (kRsRows + 7) / 8; i++) {
          if (i < rpg && r0 + i < n_rows) {
            const uint32_t o0 = (uint32_t)((p0[i * kRsSrcPitch] * a0s[0] + p0[i * kRsSrcPitch + 1] * a1s[0]) >> 4);
            const uint32_t o1 = (uint32_t)((p1[i * kRsSrcPitch] * a0s[1] + p1[i * kRsSrcPitch + 1] * a1s[1]) >> 4);
            const uint32_t o2 = (uint32_t)((p2[i * kRsSrcPitch] * a0s[2] + p2[i * kRsSrcPitch + 1] * a1s[2]) >> 4);
            const uint32_t o3 = (uint32_t)((p3[i * kRsSrcPitch] * a0s[3] + p3[i * kRsSrcPitch + 1] * a1s[3]) >> 4);
            *reinterpret_cast<uint2*>(&hq[(r0 + i) * kRsTW + 4 * q]) = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));
          }
        }
      } else {
        const uint8_t* sp = frame + src_base + r0 * S.pitch;
        for (int i = 0; i < rpg && r0 + i < n_rows; i++) {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; k++) o[k] = (uint32_t)((sp[sxs[k]] * a0s[k] + sp[sxs[k] + 1] * a1s[k]) >> 4);
          *reinterpret_cast<uint2*>(&hq[(r0 + i) * kRsTW + 4 * q]) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
          sp += S.pitch;
        }
      }
    }
  }
  __syncthreads();

  // ---- vertical pass: thread = (quad q, kRsMaxTH / 8 output rows)
  {
    constexpr int kVRows = kRsMaxTH / 8;
    const int q = tid & 31, yy0 = (tid >> 5) * kVRows;
    const int dx0 = x0 + 4 * q;
    if (dx0 < D.w) {
      uint8_t* d = frame + px_off(D, dx0, y0 + yy0);
#pragma unroll
      for (int i = 0; i < kVRows; i++) {
        const int dy = y0 + yy0 + i;
        if (dy < y1) {
          const uint32_t yo = *reinterpret_cast<const uint32_t*>(yofs + 2 * (t + dy));   // (row0 | row1 << 16)
          const uint32_t yb = *reinterpret_cast<const uint32_t*>(ybeta + 2 * (t + dy));  // (b0 | b1 << 16)
          const int r0 = (int)(yo & 0xFFFF) - row_lo, r1 = (int)(yo >> 16) - row_lo;
          const int b0 = (int)(int16_t)(yb & 0xFFFF), b1 = (int)yb >> 16;
          const uint2 u0 = *reinterpret_cast<const uint2*>(&hq[r0 * kRsTW + 4 * q]);
          const uint2 u1 = *reinterpret_cast<const uint2*>(&hq[r1 * kRsTW + 4 * q]);
          const int h0[4] = {(int)(u0.x & 0xFFFF), (int)(u0.x >> 16), (int)(u0.y & 0xFFFF), (int)(u0.y >> 16)};
          const int h1[4] = {(int)(u1.x & 0xFFFF), (int)(u1.x >> 16), (int)(u1.y & 0xFFFF), (int)(u1.y >> 16)};
          uint32_t packed = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            // b0 + b1 <= 2049 and h <= 255 * 2049 / 16, so the sum is in [0, 1022]: the saturate_cast of
            // cv::resize can never clip and is not spelled out
            const int v = (((b0 * h0[k]) >> 16) + ((b1 * h1[k]) >> 16) + 2) >> 2;
            packed |= (uint32_t)v << (8 * k);
          }
          uint8_t* dp = d + i * D.pitch;
          // a quad that straddles the right edge spills <= 3 bytes into the row padding, which nothing
          // reads before k_border rewrites it
          *reinterpret_cast<uint32_t*>(dp) = packed;
        }
      }
    }
  }
}

// TMA variant of the same tile (the common case: the source columns of 128 outputs fit BW <= 256 bytes).
// The source rows [row_lo, row_lo + bh) x [sx_lo, sx_lo + BW) of the tile arrive by ONE tensor copy
// (cp.async.bulk.tensor over level lev-1's padded planes; the box size is part of the level's tensor map,
// see resize_tile_plan); what the box covers beyond the plane is zero-filled and only ever multiplied
// by a zero coefficient.  Dynamic shared memory: bh x BW source bytes, then bh x 128 u16 of H>>4.
// The innermost TMA coordinate must put the box on a 16-byte boundary of the row (an unaligned start
// faults with "illegal instruction" on sm_100a), so the box starts at the source column rounded down to 16.
constexpr int kRsSrcAlign = 15, kRsBwSmall = 176;
__device__ __forceinline__ uint32_t mulhi_plain(uint32_t a, uint32_t b) {
  uint32_t d;
  asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));  // volatile: not to be fused into mad.hi
  return d;
}
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b) { return a + b + 2u; }
#ifndef ORBX_RS_TMA_MINB
#define ORBX_RS_TMA_MINB 5
#endif
template <int BW, int TH>  // TH: the launch's maximum tile height (batch or single-frame plan), rows per warp = TH / 8
__global__ void __launch_bounds__(256, ORBX_RS_TMA_MINB) k_resize_tma(const __grid_constant__ FrameGeom g, uint8_t* __restrict__ pyr,
                                                    const CUtensorMap* __restrict__ rs_maps, const __grid_constant__ CUtensorMap ext0_map,
                                                    const int16_t* __restrict__ xofs,
                                                    const int16_t* __restrict__ xalpha, const int16_t* __restrict__ yofs,
                                                    const int16_t* __restrict__ ybeta, int lev, int th, int bh) {
  extern __shared__ uint8_t rs_smem[];
  // the TMA destination must be 128-byte aligned; the dynamic region is only 16-byte aligned when the
  // kernel also has static shared memory, so the base is rounded up here (the launch adds 128 bytes)
  uint8_t* src_sm = rs_smem + ((128u - ((unsigned)__cvta_generic_to_shared(rs_smem) & 127u)) & 127u);
  // The horizontal sums are kept as H & ~15 in 32-bit words and the row coefficients as b << 12: (b * (H >> 4)) >> 16 ==
  // (b * (H & ~15)) >> 20 == the high word of (b << 12) * (H & ~15), so the vertical pass is one IMAD.HI per product and
  // the horizontal pass pays one mask per value (no shift)
  uint32_t* hq = reinterpret_cast<uint32_t*>(src_sm + ((bh * BW + 127) & ~127));
  unsigned long long& tile_bar = *reinterpret_cast<unsigned long long*>(hq + bh * kRsTW);
  // the row table lives in the same (dynamic) shared object as everything else: one shared-window base for the kernel
  uint4* row_tab = reinterpret_cast<uint4*>(hq + bh * kRsTW + 4);
  const LevelGeom& D = g.lv[lev];
  const int x0 = blockIdx.x * kRsTW, y0 = blockIdx.y * th;
  const int y1 = min(y0 + th, D.h);  // output rows [y0, y1)
  const int t = D.tab_off, tid = threadIdx.x;
  uint8_t* frame = pyr + (size_t)blockIdx.z * g.pyr_frame_bytes;
  const int row_lo = yofs[2 * (t + y0)], row_hi = yofs[2 * (t + y1 - 1) + 1];
  const int n_rows = min(row_hi - row_lo + 1, bh);
  const int sx_lo = xofs[t + x0] & ~kRsSrcAlign;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&tile_bar);
  if (tid == 0) {
    // level 1 of a call that reads level 0 in place takes its source rows from the caller's frames (no padding around them)
    const bool ext = lev == 1 && g.ext0 != nullptr;
    const CUtensorMap* map = ext ? &ext0_map : rs_maps + lev;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bh * BW) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(src_sm)),
                 "l"(map), "r"(sx_lo + (ext ? 0 : kPadX)), "r"(row_lo + (ext ? 0 : kPadY)), "r"((int)blockIdx.z), "r"(bar)
                 : "memory");
  }

  // ---- horizontal pass: thread = (quad q of output columns, the warp's share of the source rows); the
  // column tables are read once per thread while the tile is in flight.  The <= 8 source bytes a quad
  // needs (scale < 2) are cut out of three aligned words per row as one 64-bit window; each output is
  // then PRMT (its two neighbouring bytes) + DP2A with the packed coefficient pair (a0 | a1 << 16).
  {
    const int rpg = (n_rows + 7) >> 3;
    const int q = tid & 31, r0 = (tid >> 5) * rpg;
    const int dx0 = x0 + 4 * q;
    const bool active = dx0 < D.w && r0 < n_rows;
    int sx0 = sx_lo;
    uint32_t al[4] = {0, 0, 0, 0}, sel[4] = {0x10, 0x10, 0x10, 0x10};
    if (active) {
      int sxs[4];
      if (dx0 + 4 <= D.w) {
        const short4 sx = *reinterpret_cast<const short4*>(xofs + t + dx0);
        const uint4 a4 = *reinterpret_cast<const uint4*>(xalpha + 2 * (t + dx0));  // (a0 | a1 << 16) per column, both in [0, 2048]
        sxs[0] = sx.x; sxs[1] = sx.y; sxs[2] = sx.z; sxs[3] = sx.w;
        al[0] = a4.x; al[1] = a4.y; al[2] = a4.z; al[3] = a4.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const bool in = dx0 + k < D.w;
          sxs[k] = xofs[t + (in ? dx0 + k : dx0)];
          al[k] = in ? *reinterpret_cast<const uint32_t*>(xalpha + 2 * (t + dx0 + k)) : 0u;
        }
      }
      sx0 = sxs[0];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t dl = (uint32_t)(sxs[k] - sx0);  // 0..6
        sel[k] = dl | ((dl + 1u) << 4);
      }
    }
    if (tid < th) {  // the tile's row table for the vertical pass: byte offsets of the two H rows, b0 << 12, b1 << 12
      const int dy = min(y0 + tid, D.h - 1);
      const uint32_t yo = *reinterpret_cast<const uint32_t*>(yofs + 2 * (t + dy));
      const uint32_t yb = *reinterpret_cast<const uint32_t*>(ybeta + 2 * (t + dy));
      row_tab[tid] = make_uint4(((yo & 0xFFFFu) - (uint32_t)row_lo) * (uint32_t)(kRsTW * 4), ((yo >> 16) - (uint32_t)row_lo) * (uint32_t)(kRsTW * 4),
                                (yb & 0xFFFFu) << 12, (yb >> 16) << 12);
    }
    __syncthreads();  // every thread sees the initialised barrier
    {
      uint32_t done;
      do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar) : "memory");
      } while (!done);
    }
    if (active) {
      const int off = sx0 - sx_lo, sh = 8 * (off & 3);
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(src_sm + r0 * BW) + (off >> 2);
#pragma unroll
      for (int i = 0; i < (2 * TH + 4 + 7) / 8; i++) {
        if (i < rpg && r0 + i < n_rows) {
          const uint32_t w0 = wp[i * (BW / 4)], w1 = wp[i * (BW / 4) + 1], w2 = wp[i * (BW / 4) + 2];
          const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);  // bytes sx0 .. sx0+7
          const uint32_t o0 = __dp2a_lo(al[0], __byte_perm(lo, hi, sel[0]), 0u) & ~15u;
          const uint32_t o1 = __dp2a_lo(al[1], __byte_perm(lo, hi, sel[1]), 0u) & ~15u;
          const uint32_t o2 = __dp2a_lo(al[2], __byte_perm(lo, hi, sel[2]), 0u) & ~15u;
          const uint32_t o3 = __dp2a_lo(al[3], __byte_perm(lo, hi, sel[3]), 0u) & ~15u;
          *reinterpret_cast<uint4*>(&hq[(r0 + i) * kRsTW + 4 * q]) = make_uint4(o0, o1, o2, o3);
        }
      }
    }
  }
  __syncthreads();

  // ---- vertical pass: thread = (quad q, kRsTmaTH / 8 output rows).  (b * (H >> 4)) >> 16 is the high word of
  // (b << 12) * (H & ~15): one IMAD.HI per product, both factors as stored.
  {
    constexpr int kVRows = TH / 8;
    const int q = tid & 31, yy0 = (tid >> 5) * kVRows;
    const int dx0 = x0 + 4 * q;
    if (dx0 < D.w) {
      uint8_t* d = frame + px_off(D, dx0, y0 + yy0);
#pragma unroll
      for (int i = 0; i < kVRows; i++) {
        if (y0 + yy0 + i < y1) {
          const uint4 rt = row_tab[yy0 + i];
          const uint32_t b0 = rt.z, b1 = rt.w;  // << 12
          const uint8_t* hb = reinterpret_cast<const uint8_t*>(hq) + 16 * q;
          const uint4 u0 = *reinterpret_cast<const uint4*>(hb + rt.x);
          const uint4 u1 = *reinterpret_cast<const uint4*>(hb + rt.y);
          // b0 + b1 <= 2049 and h <= 255 * 2049 / 16, so the sum is in [0, 1022]: the saturate_cast of
          // cv::resize can never clip and is not spelled out
          // two plain multiply-highs and ONE three-input add per pixel (left to itself the compiler chains the second
          // product as a 64-bit multiply-add, which costs a zeroed register pair and a separate +2 per pixel)
          const uint32_t v0 = add3(mulhi_plain(b1, u1.x), mulhi_plain(b0, u0.x)) >> 2;
          const uint32_t v1 = add3(mulhi_plain(b1, u1.y), mulhi_plain(b0, u0.y)) >> 2;
          const uint32_t v2 = add3(mulhi_plain(b1, u1.z), mulhi_plain(b0, u0.z)) >> 2;
          const uint32_t v3 = add3(mulhi_plain(b1, u1.w), mulhi_plain(b0, u0.w)) >> 2;
          // a quad that straddles the right edge spills <= 3 bytes into the row padding, which nothing
          // reads before k_border rewrites it
          *reinterpret_cast<uint32_t*>(d + i * D.pitch) = __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
        }
      }
    }
  }
}

void resize_tile_plan(const FrameGeom& g, int lev, int frames, int* th, int* bw, int* bh) {
  // rows per tile so that the source rows of a tile fit the shared buffer (any scale factor)
  const double ry = (double)g.lv[lev - 1].h / g.lv[lev].h, rx = (double)g.lv[lev - 1].w / g.lv[lev].w;
  // source span of 128 output columns: first tap of the first .. second tap of the last
  const int need_w = (int)ceil(127.0 * rx) + 3 + kRsSrcAlign;
  *bw = need_w <= kRsBwSmall ? kRsBwSmall : (need_w <= 256 ? 256 : 0);  // 0: the LDGSTS / gather kernel
  const int max_th = *bw ? (frames >= kRsBatchFrames ? kRsTmaTH : kRsTmaTHSingle) : kRsMaxTH, max_rows = *bw ? 2 * max_th + 4 : kRsRows;
  int t = (int)((max_rows - 3) / ry);
  t = t < 1 ? 1 : (t > max_th ? max_th : t);
  *th = t;
  const int need_h = (int)ceil((t - 1) * ry) + 3;  // source rows of t output rows
  *bh = need_h < max_rows ? need_h : max_rows;
}

cudaError_t resize_configure() {  // scale factors near 2 need more than the default 48 KB of dynamic shared memory
  cudaError_t e = cudaFuncSetAttribute(k_resize_tma<kRsBwSmall, kRsTmaTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resize_tma<256, kRsTmaTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resize_tma<kRsBwSmall, kRsTmaTHSingle>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_resize_tma<256, kRsTmaTHSingle>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  return e;
}

int launch_resize_level(const FrameGeom& g, const BatchBuffers& b, int frames, int lev, cudaStream_t st) {
  int th, bw, bh;
  resize_tile_plan(g, lev, frames, &th, &bw, &bh);
  const CUtensorMap* maps = frames >= kRsBatchFrames ? b.rs_maps : b.rs_maps_single;
  dim3 grid((g.lv[lev].w + kRsTW - 1) / kRsTW, (g.lv[lev].h + th - 1) / th, frames);
  const size_t smem = 128 + (size_t)((bh * bw + 127) & ~127) + (size_t)bh * kRsTW * 4 + 16 + sizeof(uint4) * kRsTmaTH;  // alignment slack, tile, H rows (u32), mbarrier, row table
  const bool batch = frames >= kRsBatchFrames;
  if (bw == kRsBwSmall && batch)
    k_resize_tma<kRsBwSmall, kRsTmaTH><<<grid, 256, smem, st>>>(g, b.pyr, maps, b.ext0_rs_map, b.xofs, b.xalpha, b.yofs, b.ybeta, lev, th, bh);
  else if (bw == kRsBwSmall)
    k_resize_tma<kRsBwSmall, kRsTmaTHSingle><<<grid, 256, smem, st>>>(g, b.pyr, maps, b.ext0_rs_map, b.xofs, b.xalpha, b.yofs, b.ybeta, lev, th, bh);
  else if (bw == 256 && batch)
    k_resize_tma<256, kRsTmaTH><<<grid, 256, smem, st>>>(g, b.pyr, maps, b.ext0_rs_map, b.xofs, b.xalpha, b.yofs, b.ybeta, lev, th, bh);
  else if (bw == 256)
    k_resize_tma<256, kRsTmaTHSingle><<<grid, 256, smem, st>>>(g, b.pyr, maps, b.ext0_rs_map, b.xofs, b.xalpha, b.yofs, b.ybeta, lev, th, bh);
  else
    k_resize<<<grid, 256, 0, st>>>(g, b.pyr, b.xofs, b.xalpha, b.yofs, b.ybeta, lev, th);
  return 1;
}

int launch_pyramid(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  int n = 0;
  for (int lev = 1; lev < g.nlev; lev++) n += launch_resize_level(g, b, frames, lev, st);
  return n;
}

// ------------------------------------------------------------------ 19-px REFLECT_101 frame
// Only needed when the caller wants img_pyramid_ back (orb_extractor.cc:1109-1114); no kernel of
// the extractor reads the frame.  One thread per border pixel, all levels in one launch.
__global__ void __launch_bounds__(256) k_border(const __grid_constant__ FrameGeom g, uint8_t* __restrict__ pyr) {
  const int lev = blockIdx.y;
  const LevelGeom& L = g.lv[lev];
  const int W = L.w + 2 * kEdge;
  const int n_tb = 2 * kEdge * W;       // top + bottom strips
  const int n_lr = 2 * kEdge * L.h;     // left + right strips
  uint8_t* frame = pyr + (size_t)blockIdx.z * g.pyr_frame_bytes;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_tb + n_lr; i += gridDim.x * blockDim.x) {
    int x, y;
    if (i < n_tb) {
      const int r = i / W;
      x = i - r * W - kEdge;
      y = r < kEdge ? r - kEdge : L.h + (r - kEdge);
    } else {
      const int j = i - n_tb;
      y = j / (2 * kEdge);
      const int c = j - y * (2 * kEdge);
      x = c < kEdge ? c - kEdge : L.w + (c - kEdge);
    }
    frame[px_off(L, x, y)] = frame[px_off(L, reflect101(x, L.w), reflect101(y, L.h))];
  }
}

int launch_border(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  dim3 grid(32, g.nlev, frames);
  k_border<<<grid, 256, 0, st>>>(g, b.pyr);
  return 1;
}

}  // namespace orbx
