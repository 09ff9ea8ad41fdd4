// oracle/minicv -- TEST INFRASTRUCTURE, not product code.  See core/core.hpp.
#ifndef MINICV_IMGPROC_HPP
#define MINICV_IMGPROC_HPP
#include "../core/core.hpp"

namespace cv {

// cv::resize, INTER_LINEAR on CV_8UC1 only (orb_extractor.cc:1106).
inline void resize(InputArray src_, OutputArray dst_, Size dsize, double = 0, double = 0,
                   int interpolation = INTER_LINEAR) {
  assert(interpolation == INTER_LINEAR);
  Mat src = src_.getMat();
  dst_.create(dsize.height, dsize.width, src.type());  // keeps an equally sized ROI in place
  Mat dst = dst_.getMat();
  cvp_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

// cv::copyMakeBorder with BORDER_REFLECT_101 (+BORDER_ISOLATED) and equal borders
// (orb_extractor.cc:1109-1114).  When src is the interior ROI of dst the interior is left
// untouched and only the frame is written, as OpenCV does.
inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right,
                           int borderType) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  assert(top == bottom && left == right && top == left);
  Mat src = src_.getMat();
  dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
  Mat dst = dst_.getMat();
  const bool in_place = (src.data == dst.data + (size_t)top * dst.step + left);
  if (in_place) {
    Mat tmp = src.clone();
    cvp_border_reflect101_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step, top);
  } else {
    cvp_border_reflect101_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top);
  }
}

// cv::GaussianBlur, Size(7,7), sigma 2, BORDER_REFLECT_101, CV_8UC1 (orb_extractor.cc:1055):
// the 8-bit fixed-point path; src may alias dst.
inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy = 0,
                         int borderType = BORDER_DEFAULT) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && (sy == 2 || sy == 0));
  assert(borderType == BORDER_REFLECT_101);
  Mat src = src_.getMat().clone();
  dst_.create(src.rows, src.cols, src.type());
  Mat dst = dst_.getMat();
  cvp_gauss7x7_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}

}  // namespace cv
#endif
