// oracle/minicv -- TEST INFRASTRUCTURE.  Nothing from highgui is used on the hot path.
#ifndef MINICV_HIGHGUI_HPP
#define MINICV_HIGHGUI_HPP
#include "../core/core.hpp"
#endif
