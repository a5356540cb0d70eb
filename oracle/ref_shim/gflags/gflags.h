// gflags stand-in: include/caffe/common.hpp:5,24-26 and GlobalInit (common.cpp:43-50) only need these names.
#pragma once
#define GFLAGS_GFLAGS_H_
namespace gflags {
inline void ParseCommandLineFlags(int*, char***, bool) {}
}  // namespace gflags
