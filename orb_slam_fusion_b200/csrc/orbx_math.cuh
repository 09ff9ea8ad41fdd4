// orbx_math.cuh -- pixel-level arithmetic of the ORB front end, shared by every kernel.
//
// Everything here is written once as host+device inline functions so that the exact integer /
// float32 sequences can also be unit-tested on a machine without a GPU
// (tests/host_emul.cc compiles this header with g++).  On the device every float operation uses
// the explicit round-to-nearest intrinsics, so nvcc can never contract a mul+add into an FMA:
// cv::fastAtan2 and the rBRIEF steering are only reproducible without contraction
// (SURVEY.md A.5 / A.7).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define ORBX_HD __host__ __device__ __forceinline__
#else
#define ORBX_HD inline
#include <math.h>
#endif

namespace orbx {

constexpr int kEdge = 19;       // kEdgeThreshold (orb_extractor.cc:74)
constexpr int kHalfPatch = 15;  // kHalfPatchSize (:73)
constexpr int kPatch = 31;      // kPatchSize (:72)
constexpr int kFastBorder = 16; // kEdgeThreshold - 3: origin of the FAST grid (:751)

// ---- float32 without contraction -------------------------------------------------------
ORBX_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
ORBX_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
ORBX_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
ORBX_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
// cvRound(float): round half to even.  On the device the conversion instruction (F2I) issues on the
// quarter-rate XU pipe through the MIO queue; for |v| < 2^22 adding 1.5 * 2^23 rounds to the nearest
// integer (ties to even, the FADD's own rounding) and leaves it in the low mantissa bits.
ORBX_HD int f_round(float v) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
#else
  return (int)lrintf(v);
#endif
}

// ---- cv::fastAtan2 (degrees), SURVEY.md A.5 -----------------------------------------------
ORBX_HD float fast_atan2_deg(float y, float x) {
  const float scale = (float)(180 / 3.1415926535897932384626433832795);
  const float p1 = f_mul(0.9997878412794807f, scale);
  const float p3 = f_mul(-0.3258083974640975f, scale);
  const float p5 = f_mul(0.1555786518463281f, scale);
  const float p7 = f_mul(-0.04432655554792128f, scale);
  const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
  const float ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
  float a;
  if (ax >= ay) {
    const float c = f_div(ay, f_add(ax, eps));
    const float c2 = f_mul(c, c);
    a = f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = f_div(ax, f_add(ay, eps));
    const float c2 = f_mul(c, c);
    a = f_sub(90.f, f_mul(f_add(f_mul(f_add(f_mul(f_add(f_mul(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = f_sub(180.f, a);
  if (y < 0) a = f_sub(360.f, a);
  return a;
}

// ---- FAST-9/16 --------------------------------------------------------------------------
// Ring offsets (dx, dy) in cv::FAST order.
#define ORBX_RING_DX \
  { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 }
#define ORBX_RING_DY \
  { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 }

// true iff the 16-bit cyclic mask holds a run of >= 9 ones
ORBX_HD bool has_run9(uint32_t mask16) {
  const uint32_t m = mask16 | (mask16 << 16);
  uint32_t x = m & (m >> 1);
  x &= x >> 2;
  x &= x >> 4;
  x &= m >> 8;
  return (x & 0xFFFFu) != 0;
}

ORBX_HD int imin(int a, int b) { return a < b ? a : b; }
ORBX_HD int imax(int a, int b) { return a > b ? a : b; }

ORBX_HD int imin3(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __vimin3_s32(a, b, c);  // one VIMNMX3
#else
  return imin(a, imin(b, c));
#endif
}
ORBX_HD int imax3(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __vimax3_s32(a, b, c);
#else
  return imax(a, imax(b, c));
#endif
}

// FAST-9 "best" value of a pixel with centre c and ring r[16]:
//   best = max over the 16 cyclic 9-arcs of max(min_arc(c - r), -max_arc(c - r)).
// The pixel is a corner at threshold t iff best > t; cv::FAST reports response = best - 1.
// Returns 0 when the pixel is not a corner at threshold t (best <= t).
// Sliding 9-window minima / maxima by composition of 3-windows: 16 + 16 three-input ops each.
ORBX_HD int fast9_score(int c, const int (&r)[16], int t) {
#if defined(__CUDA_ARCH__)
  // Both halves of the test in one pass of 16-bit SIMD (VIMNMX3.S16x2): word k holds
  // (r[k] - c) in the high half and (c - r[k] + 256) in the low half -- one IMAD, the bias keeps the low
  // half positive so nothing borrows.  min over an arc of the high half is -max_arc(c - r).
  const uint32_t bias = (uint32_t)(c + 256) - ((uint32_t)c << 16);
  uint32_t D[16], A[16], B[16];
#pragma unroll
  for (int k = 0; k < 16; k++) D[k] = (uint32_t)r[k] * 0xFFFFu + bias;
#pragma unroll
  for (int k = 0; k < 16; k++) A[k] = __vimin3_s16x2(D[k], D[(k + 1) & 15], D[(k + 2) & 15]);
#pragma unroll
  for (int k = 0; k < 16; k++) B[k] = __vimin3_s16x2(A[k], A[(k + 3) & 15], A[(k + 6) & 15]);
  uint32_t m = __vimax3_s16x2(B[0], B[1], B[2]);
#pragma unroll
  for (int k = 3; k < 15; k += 2) m = __vimax3_s16x2(m, B[k], B[k + 1]);
  m = __vmaxs2(m, B[15]);
  const int best2 = imax((int)(m & 0xFFFFu) - 256, ((int)m) >> 16);
  return best2 > t ? best2 - 1 : 0;
#else
  int d[16], a[16], b[16];
#pragma unroll
  for (int k = 0; k < 16; k++) d[k] = c - r[k];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = imin3(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);   // min of 3
#pragma unroll
  for (int k = 0; k < 16; k++) b[k] = imin3(a[k], a[(k + 3) & 15], a[(k + 6) & 15]);   // min of 9
  int best = imax3(b[0], b[1], b[2]);
#pragma unroll
  for (int k = 3; k < 15; k += 2) best = imax3(best, b[k], b[k + 1]);
  best = imax(best, b[15]);
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = imax3(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);   // max of 3
#pragma unroll
  for (int k = 0; k < 16; k++) b[k] = imax3(a[k], a[(k + 3) & 15], a[(k + 6) & 15]);   // max of 9
  int worst = imin3(b[0], b[1], b[2]);
#pragma unroll
  for (int k = 3; k < 15; k += 2) worst = imin3(worst, b[k], b[k + 1]);
  worst = imin(worst, b[15]);
  best = imax(best, -worst);
  return best > t ? best - 1 : 0;
#endif
}

// ---- bilinear resize, 8U fixed point -----------------------------------
// vertical combine of two horizontally interpolated int32 rows
ORBX_HD int resize_vcombine(int h0, int h1, int b0, int b1) {
  int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// ---- 7x7 Gaussian, Q8.8 fixed point --------------------------------------
ORBX_HD int gauss7_h(int p0, int p1, int p2, int p3, int p4, int p5, int p6) {
  return 18 * (p0 + p6) + 34 * (p1 + p5) + 48 * (p2 + p4) + 56 * p3;  // <= 255*256
}
ORBX_HD int gauss7_v(int t0, int t1, int t2, int t3, int t4, int t5, int t6) {
  const uint32_t acc = 18u * (uint32_t)(t0 + t6) + 34u * (uint32_t)(t1 + t5) + 48u * (uint32_t)(t2 + t4) +
                       56u * (uint32_t)t3;
  return (int)((acc + 32768u) >> 16);
}
ORBX_HD int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

// ---- rBRIEF steering (orb_extractor.cc:105-113) ------------------------------------------
// (a, b) = (cos, sin) of the keypoint angle; returns the sampling offset of pattern point (px,py)
ORBX_HD void rbrief_offset_f(float a, float b, float fx, float fy, int& row, int& col) {
  row = f_round(f_add(f_mul(fx, b), f_mul(fy, a)));
  col = f_round(f_sub(f_mul(fx, a), f_mul(fy, b)));
}
ORBX_HD void rbrief_offset(float a, float b, int px, int py, int& row, int& col) {
  rbrief_offset_f(a, b, (float)px, (float)py, row, col);
}

// ---- splitmix64 and the synthetic frame generators ------------------------
ORBX_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace orbx
