// boost::shared_ptr stand-in (include/caffe/common.hpp:4 does `using boost::shared_ptr`).
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
}  // namespace boost
