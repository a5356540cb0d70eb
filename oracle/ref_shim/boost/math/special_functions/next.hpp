// boost::math::nextafter stand-in (math_functions.cpp:233-237).
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class T> T nextafter(T a, T b) { return std::nextafter(a, b); }
} }
