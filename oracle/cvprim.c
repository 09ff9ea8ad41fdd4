/*
 * oracle/cvprim.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See cvprim.h.  Compile with -ffp-contract=off: cv2's scalar results are
 * reproduced only without FMA contraction.
 */
#include "cvprim.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

int cvp_round_f(float v) { return (int)lrintf(v); }
int cvp_round_d(double v) { return (int)lrint(v); }
int cvp_floor_f(float v) {
  int i = (int)v;
  return i - (i > v);
}
int cvp_ceil_f(float v) {
  int i = (int)v;
  return i + (i < v);
}

static short sat_s16_from_float(float v) {
  int r = cvp_round_f(v);
  if (r > 32767) r = 32767;
  if (r < -32768) r = -32768;
  return (short)r;
}

static int clip_idx(int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; }

/* INTER_LINEAR, 8UC1: 11-bit fixed-point coefficients, int32 horizontal pass,
 * two-step shifted vertical pass.  Follows the arithmetic of OpenCV's
 * resizeGeneric_ / HResizeLinear / VResizeLinear<uchar,int,short,...>. */
void cvp_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                          uint8_t* dst, int dw, int dh, size_t dstride) {
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
  short* ialpha = (short*)malloc(sizeof(short) * 2 * (size_t)dw);
  int* row0 = (int*)malloc(sizeof(int) * (size_t)dw);
  int* row1 = (int*)malloc(sizeof(int) * (size_t)dw);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cvp_floor_f(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[2 * dx] = sat_s16_from_float((1.f - fx) * 2048.f);
    ialpha[2 * dx + 1] = sat_s16_from_float(fx * 2048.f);
  }
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cvp_floor_f(fy);
    fy -= sy;
    const short b0 = sat_s16_from_float((1.f - fy) * 2048.f);
    const short b1 = sat_s16_from_float(fy * 2048.f);
    const uint8_t* s0 = src + (size_t)clip_idx(sy, 0, sh) * sstride;
    const uint8_t* s1 = src + (size_t)clip_idx(sy + 1, 0, sh) * sstride;
    for (int dx = 0; dx < dw; dx++) {
      const int sx = xofs[dx];
      const int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
      const int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
      row0[dx] = s0[sx] * a0 + s0[sx1] * a1;
      row1[dx] = s1[sx] * a0 + s1[sx1] * a1;
    }
    uint8_t* d = dst + (size_t)dy * dstride;
    for (int dx = 0; dx < dw; dx++) {
      int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
      d[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
  free(xofs); free(ialpha); free(row0); free(row1);
}

static int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

void cvp_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstride,
                              uint8_t* dst, size_t dstride, int b) {
  for (int y = -b; y < h + b; y++) {
    const uint8_t* s = src + (size_t)reflect101(y, h) * sstride;
    uint8_t* d = dst + (size_t)(y + b) * dstride;
    for (int x = -b; x < w + b; x++) d[x + b] = s[reflect101(x, w)];
  }
}

/* Bresenham ring of radius 3 in OpenCV's order (x, y). */
static const int kRing[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},  {3, 0},  {3, -1},
                                 {2, -2}, {1, -3},  {0, -3},  {-1, -3}, {-2, -2}, {-3, -1},
                                 {-3, 0}, {-3, 1},  {-2, 2},  {-1, 3}};

int cvp_fast9_best(const uint8_t* p, size_t stride) {
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 16; k++)
    d[k] = v - p[(ptrdiff_t)kRing[k][1] * (ptrdiff_t)stride + kRing[k][0]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -256;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) {
      if (d[k + j] < mn) mn = d[k + j];
      if (d[k + j] > mx) mx = d[k + j];
    }
    if (mn > best) best = mn;
    if (-mx > best) best = -mx;
  }
  return best;
}

int cvp_fast9_nms_u8(const uint8_t* img, int w, int h, size_t stride, int threshold,
                     int* out_xyr, int cap) {
  if (w < 7 || h < 7) return 0;
  /* score map over the whole ROI; 0 outside the detection domain and for non-corners */
  int* score = (int*)calloc((size_t)w * (size_t)h, sizeof(int));
  ptrdiff_t off[16];
  for (int k = 0; k < 16; k++) off[k] = (ptrdiff_t)kRing[k][1] * (ptrdiff_t)stride + kRing[k][0];
  static const int order[8] = {0, 4, 2, 6, 1, 5, 3, 7};
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* p = img + (size_t)y * stride + x;
      /* every 9-arc holds one pixel of each opposite pair (k, k+8): cheap rejection before
       * the full score, as cv::FAST's own pre-test does */
      const int tb = p[0] + threshold, td = p[0] - threshold;
      int bright = 1, dark = 1;
      for (int q = 0; q < 8 && (bright | dark); q++) {
        const int a = p[off[order[q]]], b = p[off[order[q] + 8]];
        bright &= (a > tb) | (b > tb);
        dark &= (a < td) | (b < td);
      }
      if (!(bright | dark)) continue;
      const int best = cvp_fast9_best(p, stride);
      if (best > threshold) score[(size_t)y * w + x] = best - 1;
    }
  int n = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const int s = score[(size_t)y * w + x];
      if (!s) continue;
      const int* r = score + (size_t)y * w + x;
      if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] &&
          s > r[w - 1] && s > r[w] && s > r[w + 1]) {
        if (n < cap) {
          out_xyr[3 * n] = x;
          out_xyr[3 * n + 1] = y;
          out_xyr[3 * n + 2] = s;
        }
        n++;
      }
    }
  free(score);
  return n;
}

void cvp_gauss7x7_u8(const uint8_t* src, int w, int h, size_t sstride,
                     uint8_t* dst, size_t dstride) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18}; /* Q8.8 of getGaussianKernel(7, 2) */
  uint16_t* tmp = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
  uint8_t* pad = (uint8_t*)malloc((size_t)w + 6);
  for (int y = 0; y < h; y++) {
    const uint8_t* s = src + (size_t)y * sstride;
    for (int i = 0; i < 3; i++) {
      pad[i] = s[reflect101(i - 3, w)];
      pad[w + 3 + i] = s[reflect101(w + i, w)];
    }
    memcpy(pad + 3, s, (size_t)w);
    uint16_t* t = tmp + (size_t)y * w;
    for (int x = 0; x < w; x++) /* <= 255*256, fits 16 bits */
      t[x] = (uint16_t)(k[0] * (pad[x] + pad[x + 6]) + k[1] * (pad[x + 1] + pad[x + 5]) +
                        k[2] * (pad[x + 2] + pad[x + 4]) + k[3] * pad[x + 3]);
  }
  for (int y = 0; y < h; y++) {
    const uint16_t* r[7];
    for (int j = 0; j < 7; j++) r[j] = tmp + (size_t)reflect101(y + j - 3, h) * w;
    uint8_t* d = dst + (size_t)y * dstride;
    for (int x = 0; x < w; x++) {
      const uint32_t acc = (uint32_t)k[0] * ((uint32_t)r[0][x] + r[6][x]) + (uint32_t)k[1] * ((uint32_t)r[1][x] + r[5][x]) +
                           (uint32_t)k[2] * ((uint32_t)r[2][x] + r[4][x]) + (uint32_t)k[3] * r[3][x];
      d[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
  free(pad);
  free(tmp);
}

float cvp_fast_atan2(float y, float x) {
  static const float scale = (float)(180 / 3.1415926535897932384626433832795);
  const float p1 = 0.9997878412794807f * scale;
  const float p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale;
  const float p7 = -0.04432655554792128f * scale;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}
