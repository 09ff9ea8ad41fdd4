// pyramid.cu -- image pyramid (bilinear 8U resize, REFLECT_101 border) and the 7x7 Gaussian blur.
//
// Replaces OrbExtractor::ComputePyramid (orb_extractor.cc:1093-1117: cv::resize INTER_LINEAR +
// copyMakeBorder) and the clone()+cv::GaussianBlur of orb_extractor.cc:1054-1055.
// Arithmetic: SURVEY.md A.2 / A.6, bit-exact with OpenCV 4.13's 8-bit fixed-point paths.
// Both stages are HBM/L2-bound byte streams: one read and one write per pixel, 16-byte aligned
// interior rows, 32-bit packed stores, no tensor-core work.
#include "orbx_kernels.cuh"
#include "orbx_math.cuh"

namespace orbx {

// ------------------------------------------------------------------ import (level 0)
// Copies the caller's frames into the padded level-0 planes.  16 bytes per thread.
__global__ void __launch_bounds__(256) k_import(const __grid_constant__ FrameGeom g, uint8_t* __restrict__ pyr,
                                                const uint8_t* __restrict__ src, size_t row_stride,
                                                size_t frame_stride, int aligned16) {
  const int f = blockIdx.z;
  const int y = blockIdx.y;
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const LevelGeom& L = g.lv[0];
  if (x >= L.w) return;
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
  dim3 grid((g.lv[0].w + 16 * 256 - 1) / (16 * 256), g.lv[0].h, frames);
  k_import<<<grid, 256, 0, st>>>(g, b.pyr, src, row_stride, frame_stride, aligned);
  return 1;
}

// ------------------------------------------------------------------ bilinear resize
// Level `lev` from level `lev-1` (orb_extractor.cc:1106).  Block 32x8 threads, 4 px per thread.
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ FrameGeom g, uint8_t* __restrict__ pyr,
                                                const int16_t* __restrict__ xofs, const int16_t* __restrict__ xalpha,
                                                const int16_t* __restrict__ yofs, const int16_t* __restrict__ ybeta,
                                                int lev) {
  const LevelGeom& D = g.lv[lev];
  const LevelGeom& S = g.lv[lev - 1];
  const int dy = blockIdx.y * 8 + threadIdx.y;
  const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  if (dy >= D.h || dx0 >= D.w) return;
  uint8_t* frame = pyr + (size_t)blockIdx.z * g.pyr_frame_bytes;
  const int t = D.tab_off;
  const int sy0 = yofs[2 * (t + dy)], sy1 = yofs[2 * (t + dy) + 1];
  const int b0 = ybeta[2 * (t + dy)], b1 = ybeta[2 * (t + dy) + 1];
  const uint8_t* r0 = frame + px_off(S, 0, sy0);
  const uint8_t* r1 = frame + px_off(S, 0, sy1);
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int dx = dx0 + k;
    if (dx < D.w) {
      const int sx = xofs[t + dx];
      const int sx1 = min(sx + 1, S.w - 1);
      const int a0 = xalpha[2 * (t + dx)], a1 = xalpha[2 * (t + dx) + 1];
      const int h0 = r0[sx] * a0 + r0[sx1] * a1;
      const int h1 = r1[sx] * a0 + r1[sx1] * a1;
      packed |= (uint32_t)resize_vcombine(h0, h1, b0, b1) << (8 * k);
    }
  }
  uint8_t* d = frame + px_off(D, dx0, dy);
  if (dx0 + 4 <= D.w) {
    *reinterpret_cast<uint32_t*>(d) = packed;
  } else {
    for (int k = 0; dx0 + k < D.w; k++) d[k] = (uint8_t)(packed >> (8 * k));
  }
}

int launch_pyramid(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  int n = 0;
  for (int lev = 1; lev < g.nlev; lev++) {
    dim3 grid((g.lv[lev].w + 127) / 128, (g.lv[lev].h + 7) / 8, frames);
    k_resize<<<grid, dim3(32, 8), 0, st>>>(g, b.pyr, b.xofs, b.xalpha, b.yofs, b.ybeta, lev);
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

// ------------------------------------------------------------------ 7x7 Gaussian blur
// All levels in one launch; CTA = one 128x32 output tile.  Raw tile (38 rows x 136 B, reflect-101
// at the true image edge) -> horizontal Q8.8 pass into u16 -> vertical pass -> packed u32 stores.
constexpr int kBlurTW = 128, kBlurTH = 32, kBlurRawPitch = 136;

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ FrameGeom g, const uint8_t* __restrict__ pyr,
                                              uint8_t* __restrict__ blur) {
  __shared__ __align__(16) uint8_t raw[(kBlurTH + 6) * kBlurRawPitch];
  __shared__ __align__(16) uint16_t tmp[(kBlurTH + 6) * kBlurTW];
  int lev = 0;
  while (lev + 1 < g.nlev && (int)blockIdx.x >= g.lv[lev + 1].blur_tile_base) lev++;
  const LevelGeom& L = g.lv[lev];
  const int tile = blockIdx.x - L.blur_tile_base;
  const int ty = tile / L.blur_tiles_x, tx = tile - ty * L.blur_tiles_x;
  const int x0 = tx * kBlurTW, y0 = ty * kBlurTH;
  const size_t fo = (size_t)blockIdx.z * g.pyr_frame_bytes;
  const uint8_t* src = pyr + fo;
  const int tid = threadIdx.x;

  // raw[r][c]: image pixel (x0 + c - 4, y0 + r - 3); c = 1..134 used
  for (int i = tid; i < (kBlurTH + 6) * 32; i += 256) {  // the 128 aligned middle columns, 4 at a time
    const int r = i >> 5, q = i & 31;
    const int y = reflect101(y0 + r - 3, L.h);
    const int x = x0 + 4 * q;
    uint32_t v;
    if (x + 4 <= L.w) {
      v = *reinterpret_cast<const uint32_t*>(src + px_off(L, x, y));
    } else {
      v = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) v |= (uint32_t)src[px_off(L, reflect101(x + k, L.w), y)] << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(&raw[r * kBlurRawPitch + 4 + 4 * q]) = v;
  }
  for (int i = tid; i < (kBlurTH + 6) * 6; i += 256) {  // 3 + 3 halo columns
    const int r = i / 6, k = i - r * 6;
    const int c = k < 3 ? 1 + k : 132 + (k - 3);
    const int y = reflect101(y0 + r - 3, L.h);
    raw[r * kBlurRawPitch + c] = src[px_off(L, reflect101(x0 + c - 4, L.w), y)];
  }
  __syncthreads();

  for (int i = tid; i < (kBlurTH + 6) * 32; i += 256) {  // horizontal pass, 4 px per step
    const int r = i >> 5, q = i & 31;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw[r * kBlurRawPitch + 4 * q]);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    int p[12];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      p[k] = (w0 >> (8 * k)) & 255;
      p[4 + k] = (w1 >> (8 * k)) & 255;
      p[8 + k] = (w2 >> (8 * k)) & 255;
    }
    uint16_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++)  // output x = 4q+k is centred on raw column 4q+k+4
      o[k] = (uint16_t)gauss7_h(p[k + 1], p[k + 2], p[k + 3], p[k + 4], p[k + 5], p[k + 6], p[k + 7]);
    *reinterpret_cast<uint2*>(&tmp[r * kBlurTW + 4 * q]) =
        make_uint2((uint32_t)o[0] | ((uint32_t)o[1] << 16), (uint32_t)o[2] | ((uint32_t)o[3] << 16));
  }
  __syncthreads();

  uint8_t* dst = blur + fo;
  for (int i = tid; i < kBlurTH * 32; i += 256) {  // vertical pass
    const int yy = i >> 5, q = i & 31;
    const int y = y0 + yy, x = x0 + 4 * q;
    if (y >= L.h || x >= L.w) continue;
    int t[7][4];
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const uint2 v = *reinterpret_cast<const uint2*>(&tmp[(yy + j) * kBlurTW + 4 * q]);
      t[j][0] = v.x & 0xFFFF; t[j][1] = v.x >> 16; t[j][2] = v.y & 0xFFFF; t[j][3] = v.y >> 16;
    }
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      packed |= (uint32_t)gauss7_v(t[0][k], t[1][k], t[2][k], t[3][k], t[4][k], t[5][k], t[6][k]) << (8 * k);
    uint8_t* d = dst + px_off(L, x, y);
    if (x + 4 <= L.w) *reinterpret_cast<uint32_t*>(d) = packed;
    else for (int k = 0; x + k < L.w; k++) d[k] = (uint8_t)(packed >> (8 * k));
  }
}

int launch_blur(const FrameGeom& g, const BatchBuffers& b, int frames, cudaStream_t st) {
  dim3 grid(g.total_blur_tiles, 1, frames);
  k_blur<<<grid, 256, 0, st>>>(g, b.pyr, b.blur);
  return 1;
}

}  // namespace orbx
