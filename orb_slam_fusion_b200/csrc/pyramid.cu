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
                                                size_t frame_stride, int aligned16, int chunks_per_row) {
  // one thread per 16-byte chunk of a frame (rows x chunks flattened, so narrow images still fill the CTAs)
  const int f = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
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

int launch_import(const FrameGeom& g, const BatchBuffers& b, const uint8_t* src, size_t row_stride,
                  size_t frame_stride, int frames, cudaStream_t st) {
  const int aligned = ((reinterpret_cast<uintptr_t>(src) | row_stride | frame_stride) & 15) == 0;
  const int cpr = (g.lv[0].w + 15) / 16;
  dim3 grid((cpr * g.lv[0].h + 255) / 256, frames);
  k_import<<<grid, 256, 0, st>>>(g, b.pyr, src, row_stride, frame_stride, aligned, cpr);
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

int launch_pyramid(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  int n = 0;
  for (int lev = 1; lev < g.nlev; lev++) {
    // rows per tile so that the source rows of a tile fit the shared buffer (any scale factor)
    const double ratio = (double)g.lv[lev - 1].h / g.lv[lev].h;
    int th = (int)((kRsRows - 3) / ratio);
    th = th < 1 ? 1 : (th > kRsMaxTH ? kRsMaxTH : th);
    dim3 grid((g.lv[lev].w + kRsTW - 1) / kRsTW, (g.lv[lev].h + th - 1) / th, frames);
    k_resize<<<grid, 256, 0, st>>>(g, b.pyr, b.xofs, b.xalpha, b.yofs, b.ybeta, lev, th);
    n++;
  }
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
