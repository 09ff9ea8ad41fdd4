// oracle/minicv -- TEST INFRASTRUCTURE, not product code.  See core/core.hpp.
#ifndef MINICV_OPENCV_HPP
#define MINICV_OPENCV_HPP
#include "core/core.hpp"
#include "features2d/features2d.hpp"
#include "highgui/highgui.hpp"
#include "imgproc/imgproc.hpp"
#endif
