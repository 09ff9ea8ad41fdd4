// oracle/minicv -- TEST INFRASTRUCTURE, not product code.
//
// A minimal stand-in for the OpenCV C++ API surface that the reference's
// src/cam/orb_feature/orb_extractor.cc touches, so that file can be compiled
// UNMODIFIED in an image that has no OpenCV development headers.
// Every image primitive forwards to oracle/cvprim.c, which tests pin bit-for-bit
// to cv2 4.13.0.  The same header lets the product's C++ facade
// (include/cam/orb_feature/*.h) be type-checked and exercised without OpenCV.
#ifndef MINICV_CORE_HPP
#define MINICV_CORE_HPP

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../../../cvprim.h"

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32S 4
#define CV_32F 5

typedef unsigned char uchar;

inline int cvRound(double v) { return cvp_round_d(v); }
inline int cvRound(float v) { return cvp_round_f(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(float v) { return cvp_floor_f(v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(float v) { return cvp_ceil_f(v); }

namespace cv {

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { NORM_L1 = 2, NORM_HAMMING = 6 };

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  template <typename U>
  Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
           int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class Mat {
 public:
  int rows, cols;
  size_t step;
  uchar* data;

  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  // header over user memory (no ownership), like cv::Mat(rows, cols, type, void*, step)
  Mat(int r, int c, int type, void* ptr, size_t stp = 0)
      : rows(r), cols(c), step(stp ? stp : (size_t)c * esz(type)), data((uchar*)ptr), type_(type) {}

  static size_t esz(int type) { return (type == CV_32S || type == CV_32F) ? 4 : 1; }
  size_t elemSize() const { return esz(type_); }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    type_ = type;
    rows = r;
    cols = c;
    step = (size_t)c * esz(type);
    const size_t n = step * (size_t)r;
    buf_ = std::shared_ptr<uchar>(n ? new uchar[n] : nullptr, std::default_delete<uchar[]>());
    data = buf_.get();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }
  size_t step1() const { return step / elemSize(); }
  size_t total() const { return (size_t)rows * cols; }

  Mat operator()(const Rect& r) const {
    Mat m(*this);
    m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
    m.rows = r.height;
    m.cols = r.width;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat row(int y) const { return rowRange(y, y + 1); }

  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
    return m;
  }
  // cv::Mat::copyTo(OutputArray): reuses the destination memory when size and type match
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int y = 0; y < rows; y++) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
  }
  void copyTo(Mat&& dst) const { copyTo(dst); }

  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    if (m.data) std::memset(m.data, 0, m.step * (size_t)r);
    return m;
  }

  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

 private:
  int type_;
  std::shared_ptr<uchar> buf_;
};

// InputArray / OutputArray proxies: only the Mat flavour is needed.
class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  Mat* ref() const { return m_; }
 protected:
  Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& m) : _InputArray(m) {}
  _OutputArray(Mat&& m) : _InputArray(m) {}
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void create(Size sz, int type) const { m_->create(sz, type); }
  void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;

inline float fastAtan2(float y, float x) { return cvp_fast_atan2(y, x); }

// cv::norm(a, b, NORM_L1) for 8-bit single-channel matrices of equal size (frame.cc:935): sum |a - b| as a double
inline double norm(const Mat& a, const Mat& b, int normType) {
  (void)normType;
  long long s = 0;
  for (int y = 0; y < a.rows; y++) {
    const uchar *pa = a.ptr(y), *pb = b.ptr(y);
    for (int x = 0; x < a.cols; x++) s += pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x];
  }
  return (double)s;
}

}  // namespace cv

#include "persistence_stub.hpp"

#endif
