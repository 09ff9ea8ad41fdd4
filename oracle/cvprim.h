/*
 * oracle/cvprim.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Bit-exact C restatements of the OpenCV primitives that the reference's ORB
 * front end calls (OpenCV is an un-vendored dependency of the reference,
 * CMakeLists.txt:38 "find_package(OpenCV 3.2 QUIET)").  Behaviour is pinned to
 * cv2 4.13.0 (the only OpenCV in this image) by tests/test_oracle_cv2.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may link or call this.  The product path (orb_slam_fusion_b200/csrc) never
 * does.
 */
#ifndef ORB_ORACLE_CVPRIM_H
#define ORB_ORACLE_CVPRIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cvRound / cvFloor / cvCeil on float (round-half-to-even like lrintf). */
int cvp_round_f(float v);
int cvp_round_d(double v);
int cvp_floor_f(float v);
int cvp_ceil_f(float v);

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1
 * (call site: orb_extractor.cc:1106). */
void cvp_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                          uint8_t* dst, int dw, int dh, size_t dstride);

/* cv::copyMakeBorder(..., BORDER_REFLECT_101) with equal border b on all sides
 * (call sites: orb_extractor.cc:1109-1114).  dst is (w+2b) x (h+2b). */
void cvp_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstride,
                              uint8_t* dst, size_t dstride, int b);

/* cv::FAST(roi, kps, threshold, nonmaxSuppression=true), TYPE_9_16
 * (call sites: orb_extractor.cc:783,800).  Emits keypoints row-major as
 * (x, y, response) triples into out_xyr (ints), returns the count
 * (never more than cap are written; the true count is still returned). */
int cvp_fast9_nms_u8(const uint8_t* img, int w, int h, size_t stride, int threshold,
                     int* out_xyr, int cap);

/* FAST-9 corner score at one pixel: max over the 16 cyclic 9-arcs of
 * max(min_arc d, -max_arc d) with d_k = I(p) - I(p + o_k); a pixel is a corner
 * at threshold t iff this value > t, and cv::FAST reports response = value-1. */
int cvp_fast9_best(const uint8_t* p, size_t stride);

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on a
 * continuous CV_8UC1 image: the 8-bit fixed-point path (call site
 * orb_extractor.cc:1054-1055).  src and dst may not alias. */
void cvp_gauss7x7_u8(const uint8_t* src, int w, int h, size_t sstride,
                     uint8_t* dst, size_t dstride);

/* cv::fastAtan2(y, x) in degrees (call site orb_extractor.cc:99). */
float cvp_fast_atan2(float y, float x);

#ifdef __cplusplus
}
#endif
#endif
